// FP32 FFMA projection-and-sign kernel (sm_100a).
//
// Replaces LSHHasher._project_and_pack (reference lshrs/hash/lsh.py:200-211)
// for a whole batch: C = X * Rp^T in fp32 (a different summation order from
// the reference's per-band sgemv, same precision), strict `> 0`, and the sign
// bits of the columns packed little-endian -- which, with the zero-row column
// layout described in lshx_common.cuh, is np.packbits(bitorder="little") per
// band.  Also the zero-vector test of LSHRS._prepare_vector (reference
// lshrs/core/main.py:1083) fused into the same read of X.
//
// This is the "FFMA where ncu shows that wins" arm of the north star: it takes
// any dim / any band shape / any alignment, and is the arm the tcgen05 kernel
// is measured against.  128x128 tile per CTA, 8x8 register tile per thread,
// BK = 16, double-buffered shared memory with register prefetch.

#include <cstring>

#include "lshx_common.cuh"
#include "hash_small.cuh"

namespace lshx {
namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int THREADS = 256;
constexpr int LDS = BM + 4;  // row pitch of the transposed tiles (floats); 528 B keeps float4 alignment

struct Frag4 {
  float v[4];
};

// Four consecutive k of one row (row pointer null = row out of range -> zeros).
template <bool VEC4>
__device__ __forceinline__ Frag4 load4(const float* __restrict__ rowp, int dim, int k) {
  Frag4 f;
  f.v[0] = f.v[1] = f.v[2] = f.v[3] = 0.f;
  if (rowp != nullptr) {
    if (VEC4) {
      if (k < dim) {  // dim % 4 == 0 and k % 4 == 0: the whole float4 is in range
        const float4 t = __ldg(reinterpret_cast<const float4*>(rowp + k));
        f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < dim) f.v[j] = __ldg(rowp + k + j);
    }
  }
  return f;
}

template <bool VEC4>
__global__ void __launch_bounds__(THREADS, 2)
hash_ffma_kernel(const float* __restrict__ X, int64_t n, int dim, const float* __restrict__ Rp,
                 int ncols_pad, uint8_t* __restrict__ out, int sig_bytes,
                 uint8_t* __restrict__ zero_flag, int out_word_ok,
                 const int* __restrict__ tile_list, int* tile_count) {
  __shared__ __align__(16) float Xs[2][BK][LDS];
  __shared__ __align__(16) float Rs[2][BK][LDS];
  __shared__ uint32_t bits[BM][BN / 32];

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // column group
  const int ty = tid >> 4;   // row group
  // column tile is the fast index so the CTAs that share an X tile run back to back (L2 reuse)
  const int nt = ncols_pad / BN;
  // Work items: one (row tile, column tile) per CTA, or -- recompute mode, tile_list != nullptr -- a
  // grid-stride loop over the 128-row tiles a tcgen05 launch listed for FP32 recomputation
  // (hash_tc.cu, FP16x3 split); usually none, and the CTAs return at once.
  const int64_t items = (tile_list != nullptr) ? (int64_t)(*tile_count) * nt : (int64_t)gridDim.x;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
  const int ntile = (int)(item % nt);
  const int64_t m0 = (int64_t)((tile_list != nullptr) ? tile_list[item / nt] : item / nt) * BM;
  const int n0 = ntile * BN;

  // loader mapping: 128 rows x 16 k = 512 float4, two per thread
  const int lrow = tid >> 2;        // 0..63 (+64)
  const int lk = (tid & 3) * 4;     // 0,4,8,12

  for (int i = tid; i < BM * (BN / 32); i += THREADS) (&bits[0][0])[i] = 0u;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // "not all |x| <= 1e-8" per loader row (NaN counts as a violation, like np.allclose)
  bool viol0 = false, viol1 = false;

  const int KT = (dim + BK - 1) / BK;
  Frag4 xa, xb, ra, rb;

  // block-uniform 64-bit bases + per-thread 32-bit offsets (fewer live registers than four pointers)
  const float* xbase = X + m0 * (int64_t)dim;
  const float* rbase = Rp + (int64_t)n0 * dim;   // ncols_pad is a multiple of BN: always in range
  const int o0 = lrow * dim, o1 = (lrow + 64) * dim;
  const bool vx0 = (m0 + lrow < n), vx1 = (m0 + lrow + 64 < n);
  auto gload = [&](int kt) {
    const int k = kt * BK + lk;
    xa = load4<VEC4>(vx0 ? xbase + o0 : nullptr, dim, k);
    xb = load4<VEC4>(vx1 ? xbase + o1 : nullptr, dim, k);
    ra = load4<VEC4>(rbase + o0, dim, k);
    rb = load4<VEC4>(rbase + o1, dim, k);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Xs[buf][lk + j][lrow] = xa.v[j];
      Xs[buf][lk + j][lrow + 64] = xb.v[j];
      Rs[buf][lk + j][lrow] = ra.v[j];
      Rs[buf][lk + j][lrow + 64] = rb.v[j];
      viol0 |= !(fabsf(xa.v[j]) <= 1e-8f);
      viol1 |= !(fabsf(xb.v[j]) <= 1e-8f);
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();

  int buf = 0;
  for (int kt = 0; kt < KT; ++kt) {
    if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Rs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Rs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue: strict > 0, assemble each row's 128 column bits ---------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    uint32_t nib0 = 0, nib1 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      nib0 |= (acc[i][j] > 0.f ? 1u : 0u) << j;
      nib1 |= (acc[i][4 + j] > 0.f ? 1u : 0u) << j;
    }
    const int sh = (tx & 7) * 4;
    if (nib0) atomicOr(&bits[row][tx >> 3], nib0 << sh);
    if (nib1) atomicOr(&bits[row][2 + (tx >> 3)], nib1 << sh);
  }
  __syncthreads();

  const int byte0 = n0 / 8;  // first output byte of this column tile
  for (int i = tid; i < BM * (BN / 32); i += THREADS) {
    const int row = i >> 2, w = i & 3;
    const int64_t m = m0 + row;
    if (m >= n) continue;
    const int b = byte0 + w * 4;
    if (b >= sig_bytes) continue;
    const uint32_t word = bits[row][w];
    uint8_t* dst = out + m * (int64_t)sig_bytes + b;
    if (out_word_ok && b + 4 <= sig_bytes) {
      *reinterpret_cast<uint32_t*>(dst) = word;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (b + q < sig_bytes) dst[q] = (uint8_t)(word >> (8 * q));
    }
  }

  if (zero_flag != nullptr && ntile == 0) {
    // combine the four loader threads that share a row (they are adjacent lanes)
    unsigned v0 = viol0, v1 = viol1;
    v0 |= __shfl_xor_sync(0xffffffffu, v0, 1);
    v0 |= __shfl_xor_sync(0xffffffffu, v0, 2);
    v1 |= __shfl_xor_sync(0xffffffffu, v1, 1);
    v1 |= __shfl_xor_sync(0xffffffffu, v1, 2);
    if ((tid & 3) == 0) {
      if (m0 + lrow < n) zero_flag[m0 + lrow] = v0 ? 0 : 1;
      if (m0 + lrow + 64 < n) zero_flag[m0 + lrow + 64] = v1 ? 0 : 1;
    }
  }
  __syncthreads();   // recompute mode: the shared tiles are reused by the next item
  }
  // recompute mode: the last CTA to leave resets the list's counter (tile_count[0]) and the exit ticket
  // (tile_count[1]) for the slot's next launch -- every CTA read the counter before taking its ticket
  __syncthreads();
  if (tile_list != nullptr && tid == 0) {
    __threadfence();
    if (atomicAdd(tile_count + 1, 1) == (int)gridDim.x - 1) {
      tile_count[0] = 0;
      tile_count[1] = 0;
    }
  }
}


// ---- small batches (the per-vector calls of LSHRS.ingest / query) ----------------------------------
// A handful of rows cannot fill a 128-row tile and the call is pure latency, so: one CTA per OUTPUT
// BYTE, one warp per column (= signature bit), the rows in shared memory, fp32 FMA + shuffle
// reduction, `> 0`, eight warps -> one byte per row.  Same arithmetic class as the tiled kernel.
__global__ void __launch_bounds__(SMALL_THREADS)
hash_small_kernel(const float* __restrict__ X, int n, int dim, const float* __restrict__ Rp,
                  uint8_t* __restrict__ out, int sig_bytes, uint8_t* __restrict__ zero_flag) {
  extern __shared__ __align__(16) float xs[];   // [n][dim]
  __shared__ unsigned int sbits[32];     // one byte per row, built with atomicOr
  hash_small_body(X, n, dim, Rp, out, sig_bytes, zero_flag, xs, sbits);
}

// the same for ONE row that travels in the parameter block
__global__ void __launch_bounds__(SMALL_THREADS)
hash_small_one_kernel(const __grid_constant__ XRowParam x, int dim, const float* __restrict__ Rp,
                      uint8_t* __restrict__ out, int sig_bytes, uint8_t* __restrict__ zero_flag) {
  extern __shared__ __align__(16) float xs[];
  __shared__ unsigned int sbits[32];
  hash_small_body(x.v, 1, dim, Rp, out, sig_bytes, zero_flag, xs, sbits);
}

}  // namespace

int launch_hash_ffma(const HashShape& s, const float* d_X, int64_t n, const float* d_Rp,
                     uint8_t* d_out, uint8_t* d_zero_flag, cudaStream_t stream) {
  if (n <= 0) return LSHX_OK;
  const int64_t mt = (n + BM - 1) / BM;
  const int64_t blocks = mt * (s.ncols_pad / BN);
  LSHX_REQUIRE(blocks <= 0x7fffffffLL, "hash batch of %lld rows exceeds one launch", (long long)n);
  dim3 grid((unsigned)blocks);
  const bool vec4 = (s.dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_X) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(d_Rp) & 15) == 0);
  const int word_ok =
      (s.sig_bytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0) ? 1 : 0;
  if (vec4)
    hash_ffma_kernel<true><<<grid, THREADS, 0, stream>>>(d_X, n, s.dim, d_Rp, s.ncols_pad, d_out,
                                                         s.sig_bytes, d_zero_flag, word_ok, nullptr, nullptr);
  else
    hash_ffma_kernel<false><<<grid, THREADS, 0, stream>>>(d_X, n, s.dim, d_Rp, s.ncols_pad, d_out,
                                                          s.sig_bytes, d_zero_flag, word_ok, nullptr, nullptr);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

// Recompute the 128-row tiles listed in d_tile_list[0 .. *d_tile_count) (device memory) in FP32.
int launch_hash_ffma_tiles(const HashShape& s, const float* d_X, int64_t n, const float* d_Rp, uint8_t* d_out,
                           const int* d_tile_list, int* d_tile_count, int num_ctas, cudaStream_t stream) {
  if (n <= 0) return LSHX_OK;
  const bool vec4 = (s.dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_X) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(d_Rp) & 15) == 0);
  const int word_ok =
      (s.sig_bytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0) ? 1 : 0;
  dim3 grid((unsigned)num_ctas);
  if (vec4)
    hash_ffma_kernel<true><<<grid, THREADS, 0, stream>>>(d_X, n, s.dim, d_Rp, s.ncols_pad, d_out, s.sig_bytes,
                                                         nullptr, word_ok, d_tile_list, d_tile_count);
  else
    hash_ffma_kernel<false><<<grid, THREADS, 0, stream>>>(d_X, n, s.dim, d_Rp, s.ncols_pad, d_out, s.sig_bytes,
                                                          nullptr, word_ok, d_tile_list, d_tile_count);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

}  // namespace lshx

namespace lshx {

int hash_small_max_rows(const HashShape& s) {
  const int by_smem = (int)(65536 / ((size_t)s.dim * sizeof(float)));
  return by_smem < 32 ? by_smem : 32;
}

int launch_hash_small_one(const HashShape& s, const float* x_host, const float* d_Rp, uint8_t* out, uint8_t* zero_flag,
                          cudaStream_t stream) {
  LSHX_REQUIRE(s.dim <= SMALL_PARAM_FLOATS, "a row of %d floats does not fit the parameter block", s.dim);
  static thread_local XRowParam row;
  std::memcpy(row.v, x_host, (size_t)s.dim * sizeof(float));
  hash_small_one_kernel<<<s.sig_bytes, SMALL_THREADS, (size_t)s.dim * sizeof(float), stream>>>(row, s.dim, d_Rp, out,
                                                                                               s.sig_bytes, zero_flag);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int launch_hash_small(const HashShape& s, const float* d_X, int n, const float* d_Rp, uint8_t* out,
                      uint8_t* zero_flag, cudaStream_t stream) {
  if (n <= 0) return LSHX_OK;
  LSHX_REQUIRE(n <= hash_small_max_rows(s), "small-batch kernel takes at most %d rows", hash_small_max_rows(s));
  const size_t smem = (size_t)n * s.dim * sizeof(float);
  if (smem > 48 * 1024)
    LSHX_CUDA(cudaFuncSetAttribute(hash_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  hash_small_kernel<<<s.sig_bytes, SMALL_THREADS, smem, stream>>>(d_X, n, s.dim, d_Rp, out, s.sig_bytes,
                                                                  zero_flag);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

}  // namespace lshx
