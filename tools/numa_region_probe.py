"""Does it matter WHERE in host memory a pinned buffer lives?  (8-GPU box: GPUs 0-3 write D2H at 18.6 GB/s each when
they copy together, GPUs 4-7 at 45 -- DESIGN section 6.  The guest shows one NUMA node, so placement cannot be asked
for; but if the guest's memory comes from both host sockets, SOME regions may be local to the slow group.)

Allocates R pinned regions of 1 GB, reads the physical address of each from /proc/self/pagemap, and measures the
concurrent D2H rate of GPUs 0-3 and of GPUs 4-7 into every region.  Usage: python tools/numa_region_probe.py [R]"""
import json
import struct
import subprocess
import sys
import threading
import time

import torch

R = int(sys.argv[1]) if len(sys.argv) > 1 else 24
GB = 1 << 30
ngpu = torch.cuda.device_count()
print("gpus", ngpu, flush=True)
print(subprocess.run("grep -E 'MemTotal|MemAvailable|HugePages_Total|Hugepagesize' /proc/meminfo; ls /sys/devices/system/node/ | head; "
                     "cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep -E 'MemTotal|MemFree' | head -8",
                     shell=True, capture_output=True, text=True).stdout, flush=True)


def phys(addr):
    try:
        with open("/proc/self/pagemap", "rb") as f:
            f.seek((addr // 4096) * 8)
            v = struct.unpack("<Q", f.read(8))[0]
        return (v & ((1 << 55) - 1)) * 4096 if v >> 63 else None
    except Exception:  # noqa: BLE001
        return None


regions = []
t0 = time.time()
for r in range(R):
    try:
        regions.append(torch.empty(GB, dtype=torch.uint8, pin_memory=True))
    except RuntimeError as e:
        print("allocation stopped at", r, str(e)[:100])
        break
print(f"{len(regions)} regions pinned in {time.time() - t0:.1f} s", flush=True)
part = GB // 4
srcs = {g: torch.empty(part, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ngpu)}
streams = {g: torch.cuda.Stream(g) for g in range(ngpu)}


def group_rate(gpus, region):
    rates = {}
    bar = threading.Barrier(len(gpus))

    def work(i, g):
        torch.cuda.set_device(g)
        dst = region[i * part:(i + 1) * part]
        best = 0.0
        with torch.cuda.stream(streams[g]):
            for rep in range(3):
                bar.wait()
                t = time.perf_counter()
                dst.copy_(srcs[g], non_blocking=True)
                streams[g].synchronize()
                if rep:
                    best = max(best, part / (time.perf_counter() - t) / 1e9)
        rates[g] = round(best, 1)

    th = [threading.Thread(target=work, args=(i, g)) for i, g in enumerate(gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    return [rates[g] for g in gpus]


out = []
lo, hi = list(range(0, min(4, ngpu))), list(range(4, min(8, ngpu)))
for r, reg in enumerate(regions):
    pa = phys(reg.data_ptr())
    a = group_rate(lo, reg)
    b = group_rate(hi, reg) if hi else []
    out.append({"region": r, "phys_gb": None if pa is None else round(pa / GB, 1), "gpus0-3": a, "gpus4-7": b})
    print(out[-1], flush=True)
print("SUMMARY", json.dumps({"min_sum_lo": min(sum(o["gpus0-3"]) for o in out), "max_sum_lo": max(sum(o["gpus0-3"]) for o in out),
                             "min_sum_hi": min(sum(o["gpus4-7"]) for o in out) if hi else None,
                             "max_sum_hi": max(sum(o["gpus4-7"]) for o in out) if hi else None}))
