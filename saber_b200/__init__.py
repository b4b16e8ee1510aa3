"""saber_b200 — B200-native (sm_100a) implementation of SABER's SAM2 slice-wise tomogram hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every arithmetic op on the
path is a hand-written CUDA kernel reached through the C ABI in ``include/saber_b200.h``.
"""
__version__ = "0.1.0"
