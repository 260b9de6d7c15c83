"""ogbench_b200 -- the OGBench offline replay sampler (GCDataset / HGCDataset.sample) on B200.

Drop-in for `impls/utils/datasets.py` of hliuson/ogbench: same classes and call signatures, datasets resident in
HBM, hand-written sm_100a CUDA kernels behind a C-ABI (include/ogb_sampler.h).  No CPU fallback.
"""

from .datasets import ATCDataset, Dataset, GCDataset, HGCDataset, ReplayBuffer, get_size  # noqa: F401
from .device_array import DeviceArray  # noqa: F401
from .prefetch import Prefetcher  # noqa: F401

__all__ = ['Dataset', 'GCDataset', 'HGCDataset', 'ATCDataset', 'ReplayBuffer', 'DeviceArray', 'Prefetcher', 'get_size']
