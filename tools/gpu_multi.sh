#!/bin/bash
# Multi-GPU bring-up on one box (run under `gpurun --gpus N`): the relay pair test, then bench.py under torchrun for
# a list of "N:extra flags" specs.  Usage: tools/gpu_multi.sh TAG "2:--relay force --steps 5" "2:" ...
set -u
mkdir -p gpurun_out
TAG=$1; shift
timeout 300 python -m pytest tests/test_fabric_gpu.py -x -q > gpurun_out/${TAG}_fabric_test.log 2>&1
echo "fabric test rc $?"; tail -3 gpurun_out/${TAG}_fabric_test.log
i=0
for spec in "$@"; do
  n=${spec%%:*}; flags=${spec#*:}
  i=$((i+1))
  name=${TAG}_n${n}_$i
  port=$((29500 + i))
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 $flags > gpurun_out/$name.json 2> gpurun_out/$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n $flags > gpurun_out/$name.json 2> gpurun_out/$name.err
  fi
  echo "bench $name ($flags) rc $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json"))
    f = d.get("fabric", {})
    print("  value %.3f G  e2e %.1f M  ms/step %.2f  kernel-only %.3f G  frac %.3f" % (d["value"] / 1e9, d["e2e"]["value"] / 1e6,
          d["ms_per_step"], d["kernel_only"]["value"] / 1e9, d["roofline"]["frac"]))
    print("  devices", f.get("devices"), "plan", (f.get("plan") or {}).get("policy"), (f.get("plan") or {}).get("pairs"),
          "check", f.get("gather_check_all_ranks"), "frac of ceiling", f.get("value_frac_of_ceiling"))
    print("  d2h all", f.get("d2h_gbs_all_ranks_copying"), "fast half", f.get("d2h_gbs_fast_half_copying"), "h2d", f.get("h2d_gbs_all_ranks_copying"))
    print("  ms/step per rank", f.get("ms_per_step_per_rank"), "relay_error", (f.get("plan") or {}).get("relay_error"))
except Exception as e:
    print("  no JSON:", e)
PY
  tail -4 gpurun_out/$name.err | cut -c1-300
done
