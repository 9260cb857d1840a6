#!/usr/bin/env python
"""SNR estimator (SNRNet) forward at the bench shape (16 x 4 s: 16 x 2 x 256 x 512 planar features): ms per forward from a
captured graph of 10 forwards, and the per-kernel split from torch's profiler.  python tools/snr_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snr_aligned_diffse_b200 import ops  # noqa: E402
from snr_aligned_diffse_b200.synth import synth_state_dict  # noqa: E402

net = ops.SNRNetEngine()
net.load_state_dict(synth_state_dict(net.param_shapes(), seed=1), "cuda")
feat = torch.randn(16, 2, 256, 512, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3):
    out = net.forward(feat)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        out = net.forward(feat)
for _ in range(5):
    g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
row = dict(shape=[16, 2, 256, 512], ms_per_forward=round(e0.elapsed_time(e1) / 200, 4), finite=bool(torch.isfinite(out).all()))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        net.forward(feat)
    torch.cuda.synchronize()
row["kernels_us"] = {e.key.split("::")[-1][:28]: round(e.device_time_total / 5, 1) for e in prof.key_averages() if "snr_" in e.key}
print(json.dumps(row))
