#!/usr/bin/env python
"""H2D / D2H rates for subsets of the box's GPUs copying at once (profiling script): does leaving some GPUs on the
slow group idle raise the aggregate?  python tools/fabric_subsets.py > gpurun_out/fabric_subsets.json"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lshrs_b200.fabric import probe_links  # noqa: E402

out = {}
for name, devs in {"all8": list(range(8)), "0,1,4-7": [0, 1, 4, 5, 6, 7], "0,4-7": [0, 4, 5, 6, 7],
                   "0,1,2,4-7": [0, 1, 2, 4, 5, 6, 7], "4-7": [4, 5, 6, 7], "0,2,4-7": [0, 2, 4, 5, 6, 7]}.items():
    r = probe_links(devs, mbytes=512)
    r["h2d_total"] = round(sum(r["h2d_gbs"]), 1)
    r["d2h_total"] = round(sum(r["d2h_gbs"]), 1)
    out[name] = r
    print(name, r, file=sys.stderr, flush=True)
json.dump(out, sys.stdout, indent=1)
