#!/usr/bin/env python
"""Golden vectors for the general `upfirdn2d` operator from the UNMODIFIED reference
(`upfirdn2d_native`, sgmse-bbed/sgmse/backbones/ncsnpp_utils/op/upfirdn2d.py:159-200), and the oracle pinned
against them.  The reference module JIT-compiles its CUDA extension at import; only the pure-torch function is
needed here, so the module source is executed with `torch.utils.cpp_extension.load` stubbed out (the reference
file itself is read from /root/reference, not copied).

    python oracle/make_golden_upfirdn2d.py     # build container only; writes tests/golden/upfirdn2d.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ncsnpp as o_ncsnpp  # noqa: E402

REF_FILE = "/root/reference/sgmse-bbed/sgmse/backbones/ncsnpp_utils/op/upfirdn2d.py"

# (N, C, H, W, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
CASES = [
    (2, 3, 6, 10, 4, 4, 2, 2, 1, 1, 2, 1, 2, 1),      # upsample_2d with [1,3,3,1]
    (2, 3, 6, 10, 4, 4, 1, 1, 2, 2, 1, 1, 1, 1),      # downsample_2d with [1,3,3,1]
    (1, 2, 5, 7, 3, 5, 3, 2, 2, 3, 2, 3, 1, 4),       # non-square kernel, different factors per axis
    (1, 1, 8, 8, 2, 2, 1, 1, 1, 1, -1, 0, 0, -2),     # negative padding crops
    (2, 1, 4, 9, 1, 1, 2, 3, 1, 1, 0, 0, 0, 0),       # 1-tap kernel: pure zero insertion
    (1, 2, 7, 5, 6, 6, 2, 2, 2, 2, 3, 2, 2, 3),       # up and down together
    (3, 1, 9, 4, 4, 3, 1, 2, 3, 1, 5, 0, 0, 2),       # asymmetric padding
]


def load_reference_native():
    import torch.utils.cpp_extension as ce
    saved = ce.load
    ce.load = lambda *a, **k: types.SimpleNamespace()     # no JIT build: only upfirdn2d_native is called
    try:
        ns = {"__file__": REF_FILE, "__name__": "ref_upfirdn2d"}
        exec(compile(open(REF_FILE).read(), REF_FILE, "exec"), ns)
    finally:
        ce.load = saved
    return ns["upfirdn2d_native"]


def main():
    native = load_reference_native()
    out = {}
    worst = 0.0
    for i, c in enumerate(CASES):
        N, C, H, W, kh, kw = c[:6]
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn(N, C, H, W, generator=g)
        k = torch.randn(kh, kw, generator=g)
        ref = native(x, k, *c[6:])
        mine = o_ncsnpp.upfirdn2d_general(x, k, *c[6:])
        assert mine.shape == ref.shape, (c, mine.shape, ref.shape)
        worst = max(worst, float((mine - ref).abs().max()))
        out[f"x{i}"], out[f"k{i}"], out[f"y{i}"] = x.numpy(), k.numpy(), ref.numpy()
    assert worst < 5e-6, worst
    out["cases"] = np.asarray(CASES, dtype=np.int64)
    np.savez(os.path.join(ROOT, "tests", "golden", "upfirdn2d.npz"), **out)
    print("upfirdn2d golden written; oracle vs reference max abs", worst)


if __name__ == "__main__":
    main()
