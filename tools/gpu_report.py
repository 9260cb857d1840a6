#!/usr/bin/env python
"""Diagnostic report for the tcgen05 implicit-GEMM kernel (run on the GPU box; writes gpurun_out/).

Probes with structured operands (identity / delta weights) so that a wrong swizzle, descriptor or
tile mapping shows up as a recognisable permutation rather than just "mismatch".
"""
import json
import math
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

from snr_aligned_diffse_b200 import ops  # noqa: E402

DEV = "cuda"
rep = {"device": torch.cuda.get_device_name(0)}


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def stats(got, ref):
    err = (got - ref).abs()
    return dict(max_err=float(err.max()), mean_err=float(err.mean()), ref_absmax=float(ref.abs().max()),
                frac_bad=float((err > 2 ** -6 * ref.abs() + 1e-2).float().mean()))


def probe(name, B, H, W, Ci, Co, taps, kind):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Ci, H, W, generator=g).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    if kind == "identity":      # out[n] = x[n] at the centre tap
        w = torch.zeros(Co, Ci, k, k)
        for n in range(min(Co, Ci)):
            w[n, n, k // 2, k // 2] = 1.0
    elif kind == "shift":       # out[n](h,w) = x[n](h-1, w+1): tap (r=0, s=2)
        w = torch.zeros(Co, Ci, k, k)
        for n in range(min(Co, Ci)):
            w[n, n, 0, 2] = 1.0
    else:
        w = torch.randn(Co, Ci, k, k, generator=g) / math.sqrt(Ci * taps)
    w = w.to(torch.bfloat16)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), padding=k // 2)
    wt = (w.permute(0, 2, 3, 1).reshape(Co, -1) if taps == 9 else w.reshape(Co, Ci)).contiguous()
    out = ops.conv_nhwc(nhwc(x).to(DEV), wt.to(DEV), taps, impl=0)
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    st = stats(got, ref)
    if st["frac_bad"] > 0:
        bad = ((got - ref).abs() > 2 ** -6 * ref.abs() + 1e-2)
        idx = bad.nonzero()[:8].tolist()
        st["first_bad"] = [(i, float(got[tuple(i)]), float(ref[tuple(i)])) for i in idx]
        st["bad_per_channel_first16"] = bad.float().mean(dim=(0, 2, 3))[:16].tolist()
        st["bad_per_row_first16"] = bad.float().mean(dim=(0, 1, 3))[:16].tolist()
        st["bad_per_col_first16"] = bad.float().mean(dim=(0, 1, 2))[:16].tolist()
    rep[name] = st
    print(name, json.dumps(st)[:400], flush=True)


try:
    probe("id_1x1_64_one_tile", 1, 8, 16, 64, 64, 1, "identity")
    probe("id_1x1_128", 1, 8, 16, 128, 128, 1, "identity")
    probe("rand_1x1_128", 1, 8, 16, 128, 128, 1, "rand")
    probe("id_3x3_128", 1, 8, 16, 128, 128, 9, "identity")
    probe("shift_3x3_128", 1, 8, 16, 128, 128, 9, "shift")
    probe("rand_3x3_128", 1, 8, 16, 128, 128, 9, "rand")
    probe("rand_3x3_256", 1, 16, 32, 256, 256, 9, "rand")
    probe("rand_3x3_128_multi_tile", 2, 32, 64, 128, 128, 9, "rand")
    probe("rand_3x3_odd", 1, 12, 6, 128, 128, 9, "rand")
    probe("rand_1x1_768", 1, 16, 16, 256, 768, 1, "rand")
except Exception:
    rep["exception"] = traceback.format_exc()
    print(rep["exception"], flush=True)
finally:
    with open(os.path.join(OUT, "tcgen05_probe.json"), "w") as f:
        json.dump(rep, f, indent=1)
