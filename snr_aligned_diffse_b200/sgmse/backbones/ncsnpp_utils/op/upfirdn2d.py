"""`upfirdn2d` (mirror of sgmse-bbed/sgmse/backbones/ncsnpp_utils/op/upfirdn2d.py:145-156).

The reference JIT-compiles `upfirdn2d.cpp` + `upfirdn2d_kernel.cu` into a pybind op and falls back to pure torch on
CPU tensors.  Here the call goes to `snrse_upfirdn2d` in the sm_100a library (include/snrse_b200.h); CPU inputs are
moved to the GPU and the result is returned on the input's device -- there is no CPU implementation.
"""
import torch

from ..... import ops


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """input [N, C, H, W]; kernel [kh, kw]; same factor and padding on both axes, as in the reference wrapper."""
    return ops.upfirdn2d(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


def upfirdn2d_native(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Same signature as the reference's pure-torch statement (op/upfirdn2d.py:159-200); runs the CUDA operator."""
    return ops.upfirdn2d(input, kernel, (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1))
