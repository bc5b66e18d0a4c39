"""TEST INFRASTRUCTURE ONLY -- not product code.

Minimal stand-in for the third-party package `torchlibrosa` (pinned 0.0.9 by the
reference, environment.yml:71 / pyproject.toml:38) so that the reference's
`src/audioset_convnext_inf/pytorch/convnext.py` (imports at convnext.py:11-12) can be
imported UNMODIFIED in a container that has neither torchlibrosa nor librosa.

It restates the published torchlibrosa 0.0.9 / librosa 0.8.1 algorithm for exactly the
three classes the reference constructs (convnext.py:179-210); see SURVEY.md Appendix A.
Only `oracle/` and `tests/` may import this.
"""
