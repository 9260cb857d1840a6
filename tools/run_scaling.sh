#!/bin/bash
# BASELINE.json configs 2-5 at N GPUs of one box: bench.py (16 x 4 s per GPU), the 824-utterance sweep at the three
# fixed_snr settings, 60 s long-form, and the 60-NFE predictor-corrector loop.  One JSON line per run in
# gpurun_out/scale_N.jsonl.   usage: tools/run_scaling.sh N
N=${1:-1}
OUT=gpurun_out/scale_${N}.jsonl
mkdir -p gpurun_out
: > $OUT
if [ "$N" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
else
  RUN="python"
fi
run() { timeout 600 $RUN "$@" 2>> gpurun_out/scale_${N}.err | grep '^{' >> $OUT; }
run bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline
for fs in 0.17783 0.31623 0.56234; do run tools/sweep_bench.py --workload vbd --fixed-snr $fs; done
run tools/sweep_bench.py --workload long --count 2
run tools/sweep_bench.py --workload pc
cat $OUT | cut -c1-600
