#!/bin/bash
# ncu --set full of the tcgen05 hash kernel for a list of name:workload:kernel:flags[:rows] specs.  Usage: tools/gpu_ncu.sh TAG spec...
set -u
mkdir -p gpurun_out
TAG=$1; shift
for spec in "$@"; do
  IFS=: read name wl kern flags rows <<< "$spec"
  LSHX_TC_FLAGS=${flags:-} timeout 600 ncu --set full --clock-control none --import-source on -k regex:hash_tc -s 6 -c 1 \
      -f -o gpurun_out/${TAG}_$name python bench.py --workload $wl --kernel $kern --steps 1 --warmup 1 --rows ${rows:-3125000} \
      --no-cpu --no-e2e --no-rerank > gpurun_out/${TAG}_$name.log 2>&1
  echo "$name ncu exit $?"
done
ls -la gpurun_out/${TAG}_*
