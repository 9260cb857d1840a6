#!/usr/bin/env python
"""STFT + exponent transform / inverse transform + iSTFT microbench (BASELINE.json config 5):
60 s utterances at B=1 (latency-bound: 19 MB moved) and B=64, plus the 16 x 4 s shape of the bench step.

Algorithmic bytes: forward reads 4*L and writes 8*256*Tpad bytes per utterance; the inverse reads
8*256*Tpad and writes 4*L.  Timed with CUDA events on the launching stream, inputs larger than L2 at B=64
(245 MB waves, 990 MB spectrograms); at the small shapes a 256 MB buffer is rewritten between iterations.
Also checks the CUDA result against torch.stft / torch.istft of the same input on the GPU (cuFFT) and prints the
cuFFT-based torch pipeline's time for the same work ("the kernel set to beat").
    python tools/stft_bench.py > gpurun_out/stft_bench.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snr_aligned_diffse_b200 import ops  # noqa: E402

DEV = "cuda"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6448.4


def timeit(fn, iters, flush):
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for i in range(iters):
        if flush is not None:
            flush.add_(1)
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return ts[len(ts) // 2], ts[0]


def torch_fwd(w, win):
    S = torch.stft(w, n_fft=510, hop_length=128, window=win, center=True, return_complex=True)
    S = 0.15 * S.abs() ** 0.5 * torch.exp(1j * S.angle())
    return torch.nn.functional.pad(S, (0, (-S.shape[-1]) % 64))


def torch_inv(S, win, L):
    S = S / 0.15
    S = S.abs() ** 2.0 * torch.exp(1j * S.angle())
    return torch.istft(S, n_fft=510, hop_length=128, window=win, center=True, length=L)


def run(B, seconds, flush):
    L = int(seconds * 16000)
    g = torch.Generator(device=DEV).manual_seed(B)
    w = torch.randn(B, L, device=DEV, generator=g) * 0.1
    win = torch.hann_window(510, periodic=True, device=DEV)
    Y = ops.stft(w)
    tpad = Y.shape[-1]
    back = ops.istft(Y, L)
    # parity against torch on the GPU (first utterances only at the large batch: the torch pipeline needs many temporaries)
    nb = min(B, 4)
    ref = torch_fwd(w[:nb], win)
    nf = 1 + L // 128
    err_f = float((Y[:nb] - ref).abs().max() / ref.abs().max())
    refb = torch_inv(Y[:nb], win, L)
    err_b = float((back[:nb] - refb).abs().max() / refb.abs().max())
    fwd_bytes = B * (4 * L + 8 * 256 * tpad)
    inv_bytes = fwd_bytes
    iters = 20 if B * seconds < 2000 else 10
    t_f, t_f_best = timeit(lambda: ops.stft(w), iters, flush)
    t_b, t_b_best = timeit(lambda: ops.istft(Y, L), iters, flush)
    t_tf, _ = timeit(lambda: torch_fwd(w[:nb], win), 5, flush)
    t_tb, _ = timeit(lambda: torch_inv(Y[:nb], win, L), 5, flush)
    peak = peaks()
    return {"workload": f"{B} x {seconds:g} s @ 16 kHz (Tpad={tpad}, {nf} frames)", "B": B, "L": L, "tpad": tpad,
            "stft_ms": round(t_f, 4), "stft_ms_best": round(t_f_best, 4), "stft_algo_GBps": round(fwd_bytes / t_f / 1e6, 1),
            "stft_frac_of_hbm": round(fwd_bytes / t_f / 1e6 / peak, 4),
            "istft_ms": round(t_b, 4), "istft_ms_best": round(t_b_best, 4),
            "istft_algo_GBps": round(inv_bytes / t_b / 1e6, 1), "istft_frac_of_hbm": round(inv_bytes / t_b / 1e6 / peak, 4),
            "algorithmic_bytes_each_way": fwd_bytes, "hbm_peak_GBps": peak,
            "frames_per_s_fwd": round(B * tpad / t_f * 1e3), "l2": "flush between iterations" if flush is not None else "inputs > L2",
            "torch_cufft_fwd_ms_scaled_to_B": round(t_tf * B / nb, 3), "torch_cufft_inv_ms_scaled_to_B": round(t_tb * B / nb, 3),
            "max_err_vs_torch_fwd": err_f, "max_err_vs_torch_inv": err_b}


def main():
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=DEV)   # 256 MB > 126 MB L2
    cases = ((16, 4.0, flush), (1, 60.0, flush), (64, 60.0, None))
    if "--big-only" in sys.argv:      # for ncu captures of the 64 x 60 s launches
        cases = cases[2:]
    for B, sec, fl in cases:
        r = run(B, sec, fl)
        assert r["max_err_vs_torch_fwd"] < 2e-5 and r["max_err_vs_torch_inv"] < 2e-5, r
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
