"""ciri-long_b200 -- B200-native batched striped Smith-Waterman behind CIRI-long's ``ssw_wrap`` API.

Layout (only what the hot path needs):
  csrc/          hand-written sm_100a CUDA kernels + the C ABI (libssw_cuda.so, include/ssw_cuda.h)
  ssw_wrap.py    drop-in mirror of libs/striped_smith_waterman/ssw_wrap.py (Aligner / PyAlignRes /
                 CAlignRes) plus the batched entry points
  workloads.py   synthetic pair generators for the BASELINE.json configurations

The directory name is not a Python identifier; ``import ciri_long_b200`` (the loader module at the
repository root) resolves to this package.
"""
__version__ = "0.1.0"
