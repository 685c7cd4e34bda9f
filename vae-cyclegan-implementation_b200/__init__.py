"""B200-native VAE-CycleGAN training step: sm_100a CUDA kernels behind the reference's Python API.

Layout: csrc/ (CUDA kernels + C ABI, built in-tree into libvcg_b200.so), lib.py (ctypes binding),
ops.py (tensor-level kernel wrappers), plan.py (network executors), functions.py
(torch.autograd.Function wrappers), Networks.py / Losses.py (the reference's module API),
optim.py (multi-tensor Adam), dist.py (data-parallel gradient exchange), train.py (CLI)."""
__version__ = "0.1.0"
