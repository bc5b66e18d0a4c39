"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference inference hot path (waveform -> log-mel -> ConvNeXt-Tiny
-> heads) used as the parity oracle.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package; the product
package (`audioset-convnext-inf_b200/`) never does.

Parity pin status: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the reference
itself run in the build container: `oracle/make_golden.py` imports the unmodified
reference (`oracle/ref_import.py`), runs it on seeded inputs/weights and commits the
outputs under `tests/golden/`; `tests/test_oracle.py` checks the oracle against those.
"""
