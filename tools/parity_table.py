#!/usr/bin/env python
"""gpurun_out/parity_rows.jsonl (written by tests/test_gpu_configs.py and tests/test_gpu_reference.py on the GPU box)
-> markdown table.  Usage: python tools/parity_table.py gpurun_out/parity_rows.jsonl > profiles/r02_parity.md"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print("# r02 — parity of the CUDA path at every BASELINE.json config shape (tests/test_gpu_configs.py, "
      "tests/test_gpu_reference.py; 1xB200)\n")
print("bf16 activations / tensor-core operands, fp32 accumulation, against the fp32 CPU oracle (pinned on the unmodified "
      "reference, `tests/golden/oracle_vs_reference.json`) or against the unmodified reference itself run on the same GPU box "
      "in strict fp32 (`baseline/ref_runner.py --task parity`).  Inputs: the benchmark's `synth_waves(B, L, seed=1000)`, SNR "
      "estimator in the loop unless stated, explicit noise draw fed to both sides.  Snapped timestep index / t: exact in every "
      "row; noise/clean ratio 1e-4, norm factor 1e-6 relative.\n")
print("| case | item | t_30 index | spectrogram rel-L2 (bound 2e-2) | waveform SI-SDR dB (bound >= 30) | max-abs / peak (bound 4 %, 5 % above 4 s) |")
print("|---|---:|---:|---:|---:|---:|")
for r in rows:
    if "taps" in r or "si_sdr_db_min" in r:
        continue
    rl = "-" if r.get("rel_l2") is None else f"{r['rel_l2']:.3e}"
    print(f"| {r['case']} | {r['item']} | {r['t_index']} | {rl} | {r['si_sdr_db']:.1f} | {100 * r['maxabs_of_peak']:.2f} % |")
for r in rows:
    if "si_sdr_db_min" in r:
        print(f"| {r['case']} | all | = reference | - | min {r['si_sdr_db_min']:.1f} / mean {r['si_sdr_db_mean']:.1f} | max {100 * r['maxabs_of_peak_max']:.2f} % |")
for r in rows:
    if "taps" in r:
        print(f"\n**{r['case']}** (NCSN++ forward at Tpad 7552, attention blocks at n = 7552 / 472 tokens; per-module rel-L2 vs the "
              f"oracle, bound 2e-2): output {r['out_rel_l2']:.3e}; "
              + ", ".join(f"module {k}: {v:.2e}" for k, v in r["taps"].items()) + f".  Oracle time {r['oracle_seconds']} s on the box's host cores.")
