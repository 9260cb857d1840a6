"""FIR resampling entry points with the reference's names and semantics
(sgmse-bbed/sgmse/backbones/ncsnpp_utils/up_or_down_sampling.py:181-257), on top of the CUDA `upfirdn2d` operator.

`upsample_2d(x, k, factor, gain)` / `downsample_2d(x, k, factor, gain)` take `[N, C, H, W]`; `k` is a separable 1-D
filter or a square 2-D one (default: a box of `factor` taps); it is normalised to unit sum, so a constant image comes
out scaled by `gain`.  Both are one `upfirdn2d` call; they differ in which side resamples, in the filter gain and in how
the `len(k) - factor` taps of overhang are split into left / right padding.
"""
import numpy as np
import torch

from .op import upfirdn2d


def _setup_kernel(k):
    """Unit-sum square filter as float32 numpy (outer product of a 1-D filter with itself)."""
    taps = np.array(k, dtype=np.float32)
    taps = np.outer(taps, taps) if taps.ndim == 1 else taps
    if taps.ndim != 2 or taps.shape[0] != taps.shape[1]:
        raise AssertionError("FIR filter must be 1-D or square 2-D")
    return taps / taps.sum()


def _resample(x, k, factor, scale, up):
    if not (isinstance(factor, int) and factor >= 1):
        raise AssertionError("factor must be a positive integer")
    taps = _setup_kernel([1] * factor if k is None else k) * scale
    overhang = taps.shape[0] - factor
    before, after = (overhang + 1) // 2, overhang // 2
    kernel = torch.tensor(taps, device=x.device)
    if up:      # zero insertion shifts the image by factor - 1 samples: compensated on the leading side
        return upfirdn2d(x, kernel, up=factor, pad=(before + factor - 1, after))
    return upfirdn2d(x, kernel, down=factor, pad=(before, after))


def upsample_2d(x, k=None, factor=2, gain=1):
    """[N, C, H, W] -> [N, C, H*factor, W*factor]; the filter carries gain * factor^2 (zero insertion dilutes by factor^2)."""
    return _resample(x, k, factor, gain * factor ** 2, up=True)


def downsample_2d(x, k=None, factor=2, gain=1):
    """[N, C, H, W] -> [N, C, H//factor, W//factor]."""
    return _resample(x, k, factor, gain, up=False)
