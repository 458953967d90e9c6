#!/usr/bin/env python
"""Shape-matched library ceiling for the 768 -> 256-bit projection (profiling script, not product code).

The roofline denominator of the hash kernel comes from an 8192^3 cuBLAS GEMM (MEASURED_PEAKS.json).  This
script times cuBLAS FP16 GEMMs of the kernel's own shape -- M = 781 250 rows, N = 256 columns, K = 768 --
under the same power state (back to back for ~2 s), in the two forms a library user could pick for the
scaled FP16x3 split:
  * three GEMMs (hi.hi, hi.lo, lo.hi) accumulating into one fp32-accumulated output: 3 x [M,768]x[768,256]
  * one GEMM with the three terms concatenated along K: [M,2304] x [2304,256]
and, for comparison in the same run, the square 8192^3 GEMM.  Operands are already FP16 in HBM: the library
number excludes the fp32 -> (hi, lo) conversion and the sign/pack epilogue the kernel fuses, so it is an
upper bound on what a cuBLAS-based pipeline could reach for the MMA part alone.

    python tools/shape_ceiling.py > gpurun_out/shape_ceiling.json
"""
import json
import sys
import time

import torch


def timed(fn, seconds=2.0, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    # calibrate
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    reps = max(5, int(seconds * 1e3 / one))
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    # best single launch too (burst)
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return e0.elapsed_time(e1) / reps, best, reps


def main():
    dev = torch.device("cuda", 0)
    M, N, K = 781_250, 256, 768
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "M": M, "N": N, "K": K}
    a_hi = torch.randn(M, K, device=dev, dtype=torch.float16)
    a_lo = torch.randn(M, K, device=dev, dtype=torch.float16)
    b_hi = torch.randn(K, N, device=dev, dtype=torch.float16)
    b_lo = torch.randn(K, N, device=dev, dtype=torch.float16)
    c = torch.empty(M, N, device=dev, dtype=torch.float16)

    def three():
        torch.mm(a_lo, b_hi, out=c)
        torch.mm(a_hi, b_lo, out=c)
        torch.mm(a_hi, b_hi, out=c)

    ms, best, reps = timed(three)
    flop_exec = 3 * 2.0 * M * N * K
    out["three_gemms_fp16"] = {"ms": ms, "best_ms": best, "reps": reps, "executed_tflops": flop_exec / ms / 1e9,
                               "algorithmic_tflops": flop_exec / 3 / ms / 1e9,
                               "vectors_per_s": M / (ms * 1e-3)}
    del a_lo
    a3 = torch.randn(M, 3 * K, device=dev, dtype=torch.float16)
    b3 = torch.randn(3 * K, N, device=dev, dtype=torch.float16)

    def fused():
        torch.mm(a3, b3, out=c)

    ms, best, reps = timed(fused)
    out["one_gemm_k2304_fp16"] = {"ms": ms, "best_ms": best, "reps": reps, "executed_tflops": flop_exec / ms / 1e9,
                                  "algorithmic_tflops": flop_exec / 3 / ms / 1e9,
                                  "vectors_per_s": M / (ms * 1e-3)}
    del a3, b3, a_hi, c
    # fp32 input read the way the kernel reads it: one TF32 GEMM (single pass, NOT accurate enough) for scale
    x32 = torch.randn(M, K, device=dev, dtype=torch.float32)
    r32 = torch.randn(K, N, device=dev, dtype=torch.float32)
    c32 = torch.empty(M, N, device=dev, dtype=torch.float32)
    torch.backends.cuda.matmul.allow_tf32 = True

    def tf32():
        torch.mm(x32, r32, out=c32)

    ms, best, reps = timed(tf32)
    out["one_gemm_tf32_from_fp32"] = {"ms": ms, "best_ms": best, "reps": reps,
                                      "executed_tflops": 2.0 * M * N * K / ms / 1e9,
                                      "vectors_per_s": M / (ms * 1e-3)}
    torch.backends.cuda.matmul.allow_tf32 = False

    def fp32():
        torch.mm(x32, r32, out=c32)

    ms, best, reps = timed(fp32, seconds=1.0)
    out["one_gemm_fp32_cublas"] = {"ms": ms, "best_ms": best, "reps": reps,
                                   "executed_tflops": 2.0 * M * N * K / ms / 1e9,
                                   "vectors_per_s": M / (ms * 1e-3)}
    del x32, r32, c32
    n = 8192
    p = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    q = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    o = torch.empty(n, n, device=dev, dtype=torch.bfloat16)

    def square():
        torch.mm(p, q, out=o)

    ms, best, reps = timed(square)
    out["square_8192_bf16"] = {"ms": ms, "best_ms": best, "reps": reps, "sustained_tflops": 2.0 * n ** 3 / ms / 1e9,
                               "burst_tflops": 2.0 * n ** 3 / best / 1e9}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"shape_ceiling done in {time.time() - t0:.1f} s", file=sys.stderr)
