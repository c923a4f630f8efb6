"""lkgd_b200 - B200-native (sm_100a) implementation of the LKGD / Stable-Video-Diffusion denoise hot path.

Host side: Python / PyTorch (device memory, streams, torch.distributed).  Compute: hand-written CUDA kernels
behind the C ABI of ``include/lkgd_b200.h`` (``lkgd_b200/lib/liblkgd_b200.so``).  There is no CPU fallback."""

__version__ = "0.1.0"
