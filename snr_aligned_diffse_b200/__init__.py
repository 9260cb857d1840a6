"""B200-native (sm_100a) implementation of the SNR-aligned diffusion speech-enhancement hot path.

Host side: Python mirror of the reference's `sgmse` API (sub-package `sgmse`), calling
hand-written CUDA kernels in `csrc/` through the C-ABI shared library declared in
`include/snrse_b200.h`.  There is no CPU fallback: importing `._lib` without the built library, or
calling a compute entry point without a B200, raises.
"""
__version__ = "0.1.0"
