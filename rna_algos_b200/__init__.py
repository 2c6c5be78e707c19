"""rna_algos_b200 — B200-native (sm_100a) McCaskill inside/outside, base-pairing probabilities,
centroid estimator and Durbin pair-HMM forward-backward behind the public API of heartsh/rna-algos.

The compute path is the in-tree CUDA library `librna_algos_b200.so` (C ABI: include/rna_algos_b200.h).
There is no CPU fallback: importing `rna_algos_b200.api` without the built library raises ImportError,
and creating a Handle without a CUDA device raises RnaError(RNA_ERR_NO_DEVICE).
"""
from . import tables  # noqa: F401  (pure data; usable without the library)

__all__ = ["tables", "api"]
__version__ = "0.1.0"
