// Body of the small-batch hash kernel (hash_ffma.cu: hash_small_kernel) as a device function, so that the device
// band index can fuse it with its lookup / join into ONE launch on the latency path (index_join.cu).
#pragma once
#include <cstdint>

namespace lshx {

constexpr int SMALL_THREADS = 256;

// ONE vector of up to 1024 floats can travel in a kernel's PARAMETER block (sm_70+ launches take 32 KB of
// parameters): it arrives with the launch itself -- no H2D copy in front of the kernel, no fetch from pinned host
// memory by the CTAs (measured: 5 us of a 27 us latency kernel).
constexpr int SMALL_PARAM_FLOATS = 1024;
struct alignas(16) XRowParam { float v[SMALL_PARAM_FLOATS]; };

// One CTA per OUTPUT BYTE (blockIdx.x), one warp per column (= signature bit), the n rows in shared memory `xs`
// ([n][dim] floats), fp32 FMA + shuffle reduction, `> 0`, eight warps -> one byte per row.  X / out / zero_flag
// may be mapped pinned host memory.  `sbits`: 32 words of shared memory.  Ends with the bytes stored (no barrier).
__device__ __forceinline__ void hash_small_body(const float* __restrict__ X, int n, int dim,
                                                const float* __restrict__ Rp, uint8_t* __restrict__ out, int sig_bytes,
                                                uint8_t* __restrict__ zero_flag, float* xs, unsigned int* sbits) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the rows -> shared memory.  This is a latency path (one CTA per output byte, nothing else in flight), so
  // loads are issued in batches before anything consumes them: 16-byte loads when the layout allows (one round
  // trip for a 768-float row, which matters when X is pinned HOST memory read over PCIe), else four in flight.
  const int total = n * dim;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    const float4* X4 = reinterpret_cast<const float4*>(X);
    float4* xs4 = reinterpret_cast<float4*>(xs);
    for (int i = tid; i < (total >> 2); i += 2 * SMALL_THREADS) {
      const int i2 = i + SMALL_THREADS;
      const float4 a = X4[i];
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i2 < (total >> 2)) b = X4[i2];
      xs4[i] = a;
      if (i2 < (total >> 2)) xs4[i2] = b;
    }
  } else {
    for (int i = tid; i < total; i += 4 * SMALL_THREADS) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * SMALL_THREADS < total) ? X[i + u * SMALL_THREADS] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * SMALL_THREADS < total) xs[i + u * SMALL_THREADS] = v[u];
    }
  }
  if (tid < 32) sbits[tid] = 0u;
  __syncthreads();

  const float* rrow = Rp + (int64_t)(blockIdx.x * 8 + warp) * dim;   // column of this warp
  for (int i0 = 0; i0 < n; i0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int k = lane;
    // eight strided elements of the projection row per batch, loaded before the first FMA needs one (a
    // dependent chain of 24 L2 round trips was most of this kernel's time); the order of the FMAs -- k
    // ascending per lane -- is unchanged, so the bits are too
    for (; k + 7 * 32 < dim; k += 8 * 32) {
      float rv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) rv[u] = __ldg(rrow + k + u * 32);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (i0 + j < n) acc[j] = fmaf(rv[u], xs[(i0 + j) * dim + k + u * 32], acc[j]);
      }
    }
    for (; k < dim; k += 32) {
      const float rv = __ldg(rrow + k);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j < n) acc[j] = fmaf(rv, xs[(i0 + j) * dim + k], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && i0 + j < n && v > 0.f) atomicOr(&sbits[i0 + j], 1u << warp);
    }
  }
  __syncthreads();
  if (tid < n) out[(int64_t)tid * sig_bytes + blockIdx.x] = (uint8_t)sbits[tid];

  if (zero_flag != nullptr && blockIdx.x == 0) {
    for (int i = warp; i < n; i += SMALL_THREADS / 32) {
      bool viol = false;
      for (int k = lane; k < dim; k += 32) viol |= !(fabsf(xs[i * dim + k]) <= 1e-8f);
      const unsigned any = __ballot_sync(0xffffffffu, viol);
      if (lane == 0) zero_flag[i] = any ? 0 : 1;
    }
  }
}

}  // namespace lshx
