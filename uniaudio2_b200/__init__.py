"""uniaudio2_b200 - B200-native (sm_100a) inference hot path for UniAudio 2.0.

Host side is Python/PyTorch (device memory, streams, torch.distributed plumbing); the compute runs in
hand-written CUDA kernels behind the C ABI of include/ua2_b200.h (libua2_b200.so, bound with ctypes).
"""
__version__ = "0.1.0"
