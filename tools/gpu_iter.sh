#!/bin/bash
# One bring-up iteration on the GPU box: tc_check (parity of the tcgen05 kernel, smallest shapes first),
# then kernel-only bench lines for a list of LSHX_TC_FLAGS / kernel arms.  Usage: tools/gpu_iter.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-it}
for f in ${TCFLAGS_LIST:-default}; do
  if [ "$f" = default ]; then unset LSHX_TC_FLAGS; else export LSHX_TC_FLAGS=$f; fi
  timeout 900 python tools/tc_check.py > gpurun_out/${TAG}_tc_check_$f.log 2>&1
  echo "tc_check flags=$f exit $?"; tail -2 gpurun_out/${TAG}_tc_check_$f.log
  grep -E "FAIL|TIMED" gpurun_out/${TAG}_tc_check_$f.log | head -8
done
unset LSHX_TC_FLAGS
run() {  # name workload kernel flags [split]
  LSHX_TC_SPLIT=${5:-${LSHX_TC_SPLIT:-}} LSHX_TC_FLAGS=${4:-} timeout 300 python bench.py --workload $2 --kernel $3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-rerank \
      > gpurun_out/${TAG}_$1.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$1.json"))
    print("$1", "kernel-only %.1fM" % (d["kernel_only"]["value"] / 1e6), "value %.1fM" % (d["value"] / 1e6),
          "launch ms %.4f" % d["roofline"]["avg_launch_ms"], "frac %.3f" % d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
except Exception as e:
    print("$1 failed", e)
PY
}
for spec in "$@"; do
  [ "$spec" = "$TAG" ] && continue
  IFS=: read name wl kern flags split <<< "$spec"
  run "$name" "$wl" "$kern" "$flags" "$split"
done
