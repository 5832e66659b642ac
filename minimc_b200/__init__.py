"""minimc_b200: B200-native particle-history transport loop with the API surface of agtumulak/minimc.

The product is `libminimc_b200.so` (sm_100a CUDA kernels + C++ host behind the C ABI of
include/minimc_b200.h); this package is the ctypes binding, the deck generators and the
torch.distributed plumbing used by bench.py and the tests.
"""
__version__ = "0.1.0"
