"""ma3_b200: B200-native (sm_100a) sampling path for Make-An-Audio-3.

Host side is PyTorch (device memory, streams, torch.distributed); all arithmetic on the path runs in the
hand-written CUDA kernels of csrc/ behind the C ABI in include/ma3_b200.h.
"""
__version__ = "0.1.0"
