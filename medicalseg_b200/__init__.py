"""medicalseg_b200 — B200-native (sm_100a) hot path for PaddleCV-SIG/MedicalSeg's VNet.

Host side in Python over PyTorch tensors (device memory, streams, torch.distributed); every device op is a
hand-written CUDA kernel behind the C ABI in include/medseg_b200.h (libmedseg_b200.so).  No CPU fallback.
"""
__version__ = "0.1.0"
