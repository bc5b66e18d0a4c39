"""`audioset_convnext_inf.pytorch.pytorch_utils` hot-path subset: `forward` (pytorch_utils.py:63-137, the batch
evaluation loop, served by the double-buffered HostPipeline) and `move_data_to_device` (pytorch_utils.py:9-15)."""
import torch

from audioset_convnext_inf_b200.evalloop import forward  # noqa: F401


def move_data_to_device(x, device):
    """pytorch_utils.py:9-15: float / int arrays become tensors on `device`; anything else is returned unchanged."""
    if "float" in str(x.dtype):
        x = torch.Tensor(x)
    elif "int" in str(x.dtype):
        x = torch.LongTensor(x)
    else:
        return x
    return x.to(device)


__all__ = ["forward", "move_data_to_device"]
