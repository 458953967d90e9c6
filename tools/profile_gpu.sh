#!/bin/bash
# Profiling recipe (run under gpurun on one B200): plain run, ncu launch list, ncu --set full of the
# two hot kernels.  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --rows 3125000 --no-cpu --no-e2e --no-configs --queries 2048"
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
# round 2: the device band index (join kernel) through the API leg of the bench
CMD2="python bench.py --steps 1 --warmup 3 --rows 781250 --no-cpu --no-configs --no-rerank"
ncu --set full --clock-control none --import-source on -k regex:index_join_kernel -s 2 -c 1 \
    -f -o gpurun_out/${TAG}_index_join $CMD2 > gpurun_out/${TAG}_ncu_index.log 2>&1
echo "index join capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:index_hash_query -s 20 -c 2 \
    -f -o gpurun_out/${TAG}_index_query_small $CMD2 > gpurun_out/${TAG}_ncu_index_small.log 2>&1
echo "index latency-path capture exit $?"
CMD3="python bench.py --workload hash128 --rows 25000000 --steps 1 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:hash_tc -s 6 -c 1 \
    -f -o gpurun_out/${TAG}_hash_tc_dim128 $CMD3 > gpurun_out/${TAG}_ncu_hash128.log 2>&1
echo "hash128 capture exit $?"
ls -la gpurun_out | tail -12
