"""elections.dtree_b200 — B200 (sm_100a) implementation of elections.dtree's posterior-simulation
hot path (sample_posterior / sample_predictive on a Dirichlet-tree + IRV social choice).

    csrc/      CUDA kernels, host tree and the C ABI (include/dtree_b200.h) -> lib/libdtree_b200.so
    api.py     host-side mirror of the reference's R-facing API over that ABI
    sharding.py  one-process-per-GPU sharding of elections + the win-count all-reduce
    workloads.py the pinned synthetic inputs of the benchmark configurations
"""
from . import workloads  # noqa: F401
from . import _native, sharding  # noqa: F401
from .api import (dirichlet_tree, dirtree, gseed, irv_indices, reset, sample_posterior,  # noqa: F401
                  sample_predictive, set_seed, social_choice, update)
