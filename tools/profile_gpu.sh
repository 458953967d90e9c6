#!/bin/bash
# Profiling recipe (run under gpurun on one B200): plain run, ncu launch list, ncu --set full of the
# two hot kernels.  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --rows 3125000 --no-cpu --no-e2e --queries 2048"
$CMD > gpurun_out/${TAG}_prof_plain.json 2> gpurun_out/${TAG}_prof_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:hash_tc -s 4 -c 2 \
    -f -o gpurun_out/${TAG}_hash_tc $CMD > gpurun_out/${TAG}_ncu_hash.log 2>&1
echo "hash capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 2 -c 2 \
    -f -o gpurun_out/${TAG}_rerank $CMD > gpurun_out/${TAG}_ncu_rerank.log 2>&1
echo "rerank capture exit $?"
ls -la gpurun_out | tail -12
