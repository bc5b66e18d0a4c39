"""`audioset_convnext_inf.pytorch.convnext` of the reference (convnext.py:44-87, 130-511, 514-541, 641-708), served by
the B200 package: same class / factory names, ctor kwargs, state-dict keys and return types."""
from audioset_convnext_inf_b200.convnext import (Block, ConvNeXt, LayerNorm, LogmelFilterBank,  # noqa: F401
                                                 Spectrogram, convnext_tiny, load_checkpoint)

__all__ = ["Block", "ConvNeXt", "LayerNorm", "convnext_tiny"]
