// Candidate rerank kernel (sm_100a): gather -> normalise -> dot -> top-k / top-p.
//
// Replaces cosine_similarity + top_k_cosine (reference
// lshrs/utils/similarity.py:80-90, 157-183; l2_norm lshrs/utils/norm.py:48-61)
// and the rank-fraction cut of LSHRS.query (reference lshrs/core/main.py:646-658).
//
// One CTA per query.  Each warp streams whole candidate rows with 128-bit
// L1-bypassing loads (two rows x four loads in flight per lane), accumulates
// dot(c, q) and dot(c, c) in fp32, and a lane-0 epilogue turns them into
//     score = dot / (sqrt(cc) * sqrt(qq))
// which is packed with the candidate's position into one 64-bit sort key in
// shared memory.  The CTA then sorts the keys (bitonic, descending score, ties
// by ascending position) and writes the first `limit` of them.  HBM-bound: per
// query the kernel must read n*dim*4 bytes of candidates and nothing else of
// note, so the roofline is bytes / HBM bandwidth.

#include "lshx_common.cuh"

namespace lshx {
namespace {

constexpr int RR_THREADS = 256;
constexpr int RR_WARPS = RR_THREADS / 32;
constexpr int RR_MAX_CAP = 16384;  // sort-buffer entries (128 KB of shared memory)

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Monotone float -> uint32 map (larger score, larger key); NaN ranks below -inf.
__device__ __forceinline__ uint32_t order_key(float s) {
  if (s != s) return 0u;
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
  if (k == 0u) return __int_as_float(0x7fc00000);
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot(v, q) and dot(v, v) for TWO rows at once, warp-cooperative: both rows' loads are issued
// before any FMA so that 8 x 16 B per lane are in flight.  VEC4: dim % 4 == 0, 16 B aligned.
template <bool VEC4>
__device__ __forceinline__ void rows2_dots(const float* __restrict__ va, const float* __restrict__ vb,
                                           const float* __restrict__ qs, int dim, int lane,
                                           float& da, float& ca, float& db, float& cb) {
  da = ca = db = cb = 0.f;
  if (VEC4) {
    const float4* a4 = reinterpret_cast<const float4*>(va);
    const float4* b4 = reinterpret_cast<const float4*>(vb);
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    const int nv = dim >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < nv; i += 128) {
      float4 x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = (i + 32 * u < nv) ? ld_stream(a4 + i + 32 * u) : z;
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = (i + 32 * u < nv) ? ld_stream(b4 + i + 32 * u) : z;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 q = (i + 32 * u < nv) ? q4[i + 32 * u] : z;
        da = fmaf(x[u].x, q.x, da); da = fmaf(x[u].y, q.y, da);
        da = fmaf(x[u].z, q.z, da); da = fmaf(x[u].w, q.w, da);
        ca = fmaf(x[u].x, x[u].x, ca); ca = fmaf(x[u].y, x[u].y, ca);
        ca = fmaf(x[u].z, x[u].z, ca); ca = fmaf(x[u].w, x[u].w, ca);
        db = fmaf(y[u].x, q.x, db); db = fmaf(y[u].y, q.y, db);
        db = fmaf(y[u].z, q.z, db); db = fmaf(y[u].w, q.w, db);
        cb = fmaf(y[u].x, y[u].x, cb); cb = fmaf(y[u].y, y[u].y, cb);
        cb = fmaf(y[u].z, y[u].z, cb); cb = fmaf(y[u].w, y[u].w, cb);
      }
    }
  } else {
    for (int i = lane; i < dim; i += 32) {
      const float x = __ldg(va + i), y = __ldg(vb + i), q = qs[i];
      da = fmaf(x, q, da);
      ca = fmaf(x, x, ca);
      db = fmaf(y, q, db);
      cb = fmaf(y, y, cb);
    }
  }
  da = warp_sum(da);
  ca = warp_sum(ca);
  db = warp_sum(db);
  cb = warp_sum(cb);
}

__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, unsigned cap, int tid) {
  for (unsigned size = 2; size <= cap; size <<= 1) {
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned t = tid; t < (cap >> 1); t += RR_THREADS) {
        const unsigned i = 2 * t - (t & (stride - 1));
        const unsigned j = i + stride;
        const bool desc = ((i & size) == 0);
        const uint64_t a = keys[i], b = keys[j];
        if ((a < b) == desc) {
          keys[i] = b;
          keys[j] = a;
        }
      }
      __syncthreads();
    }
  }
}

template <bool VEC4>
__global__ void __launch_bounds__(RR_THREADS, 4)
rerank_kernel(RerankArgs a, unsigned cap, int q_floats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + (size_t)q_floats * sizeof(float));
  __shared__ float red[RR_WARPS];
  __shared__ int zero_cnt;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dim = a.dim;

  for (int64_t qi = blockIdx.x; qi < a.nq; qi += gridDim.x) {
    // ---- query into shared memory, ||q||^2 ------------------------------------
    if (tid == 0) zero_cnt = 0;
    float part = 0.f;
    const float* qg = a.Q + qi * (int64_t)dim;
    for (int i = tid; i < dim; i += RR_THREADS) {
      const float x = __ldg(qg + i);
      qs[i] = x;
      part = fmaf(x, x, part);
    }
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    float qq = 0.f;
#pragma unroll
    for (int w = 0; w < RR_WARPS; ++w) qq += red[w];
    const float qnorm = sqrtf(qq);

    const int64_t base = a.offs[qi];
    const int64_t n = a.cand_counts ? (int64_t)a.cand_counts[qi] : a.offs[qi + 1] - base;

    // result count: top_k_cosine's k and/or LSHRS.query's max(1, ceil(n * top_p))
    int64_t limit;
    if (a.p > 0.0) {
      int64_t lp = (int64_t)ceil((double)n * a.p);
      if (lp < 1) lp = 1;
      limit = (a.k > 0 && a.k < lp) ? a.k : lp;
    } else {
      limit = a.k;
    }
    if (limit > n) limit = n;
    if (limit > a.out_stride) limit = a.out_stride;

    // ---- score candidates in chunks that fit the sort buffer -------------------
    const bool single = !a.select || (n <= (int64_t)cap);
    const unsigned half = cap >> 1;
    const int64_t chunk = !a.select ? (n > 0 ? n : 1) : (single ? (int64_t)cap : (int64_t)half);
    const unsigned dst0 = single ? 0u : half;
    if (a.select && !single)
      for (unsigned i = tid; i < half; i += RR_THREADS) keys[i] = 0ull;

    for (int64_t c0 = 0; c0 < n || c0 == 0; c0 += chunk) {
      const int64_t cnt = (n - c0 < chunk) ? (n - c0) : chunk;
      for (int64_t j = warp; j < cnt; j += 2 * RR_WARPS) {
        const int64_t jb = j + RR_WARPS;
        const bool has_b = jb < cnt;
        int64_t ra = a.ids ? __ldg(a.ids + base + c0 + j) : (base + c0 + j);
        int64_t rb = has_b ? (a.ids ? __ldg(a.ids + base + c0 + jb) : (base + c0 + jb)) : ra;
        const bool oka = (ra >= 0 && ra < a.n_vectors);
        const bool okb = (rb >= 0 && rb < a.n_vectors);
        if (!oka) ra = 0;
        if (!okb) rb = 0;
        float da, ca, db, cb;  // row b is row a again when the chunk has an odd tail
        rows2_dots<VEC4>(a.V + ra * (int64_t)dim, a.V + rb * (int64_t)dim, qs, dim, lane, da, ca, db, cb);
        if (lane == 0) {
          float sa = da / (sqrtf(ca) * qnorm);
          if (!oka) sa = __int_as_float(0x7fc00000);
          if (ca == 0.f || !oka) atomicAdd(&zero_cnt, 1);
          if (a.all_scores) a.all_scores[base + c0 + j] = sa;
          if (a.big_keys)
            a.big_keys[qi * a.big_stride + c0 + j] =
                ((uint64_t)order_key(sa) << 32) | (uint64_t)(0xffffffffu - (uint32_t)(c0 + j));
          if (a.select)
            keys[dst0 + j] = ((uint64_t)order_key(sa) << 32) | (uint64_t)(0xffffffffu - (uint32_t)(c0 + j));
          if (has_b) {
            float sb = db / (sqrtf(cb) * qnorm);
            if (!okb) sb = __int_as_float(0x7fc00000);
            if (cb == 0.f || !okb) atomicAdd(&zero_cnt, 1);
            if (a.all_scores) a.all_scores[base + c0 + jb] = sb;
            if (a.big_keys)
              a.big_keys[qi * a.big_stride + c0 + jb] =
                  ((uint64_t)order_key(sb) << 32) | (uint64_t)(0xffffffffu - (uint32_t)(c0 + jb));
            if (a.select)
              keys[dst0 + jb] = ((uint64_t)order_key(sb) << 32) | (uint64_t)(0xffffffffu - (uint32_t)(c0 + jb));
          }
        }
      }
      if (a.select) {
        for (int64_t i = cnt + tid; i < chunk; i += RR_THREADS) keys[dst0 + i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(keys, cap, tid);  // ends with __syncthreads()
      }
      if (n == 0) break;
    }
    __syncthreads();

    // ---- results ----------------------------------------------------------------
    if (a.select) {
      for (int64_t i = tid; i < limit; i += RR_THREADS) {
        const uint64_t kv = keys[i];
        a.out_pos[qi * (int64_t)a.out_stride + i] = (int32_t)(0xffffffffu - (uint32_t)(kv & 0xffffffffu));
        a.out_score[qi * (int64_t)a.out_stride + i] = key_score((uint32_t)(kv >> 32));
      }
      if (tid == 0) a.out_count[qi] = (int32_t)limit;
    }
    if (a.big_keys)   // pad the query's slot range up to the power-of-two stride the global sort works on
      for (int64_t i = n + tid; i < a.big_stride; i += RR_THREADS) a.big_keys[qi * a.big_stride + i] = 0ull;
    if (tid == 0 && a.out_zero) a.out_zero[qi] = zero_cnt + ((qq == 0.f) ? 1 : 0);
    __syncthreads();
  }
}

// Oversized selections (more than RR_MAX_CAP candidates for one query AND more than RR_MAX_CAP / 2 results,
// e.g. get_above_p(p = 0.95) on a 16 x 4 index where every bucket holds 1/16 of the corpus): the scoring
// pass above leaves every candidate's 64-bit key in global memory (stride = next power of two) and one
// 1024-thread CTA per query sorts them there -- bitonic network on L2-resident keys, strides below the
// shared-memory tile finished in shared memory -- then writes the first `limit`.  Rare path: the reference
// handles any candidate count (np.argpartition + argsort, similarity.py:174-179), so this must too.
constexpr int BS_THREADS = 1024;
constexpr unsigned BS_TILE = 4096;   // keys sorted / merged per shared-memory tile (32 KB)

__global__ void __launch_bounds__(BS_THREADS)
rerank_bigsort_kernel(RerankArgs a) {
  __shared__ uint64_t tile[BS_TILE];
  const int tid = threadIdx.x;
  const int64_t qi = blockIdx.x;
  uint64_t* keys = a.big_keys + qi * a.big_stride;
  const uint64_t cap = (uint64_t)a.big_stride;
  const int64_t n = a.cand_counts ? (int64_t)a.cand_counts[qi] : a.offs[qi + 1] - a.offs[qi];

  // phase 1: every BS_TILE-key tile fully sorted in shared memory, direction alternating like the network's
  for (uint64_t t0 = 0; t0 < cap; t0 += BS_TILE) {
    for (unsigned i = tid; i < BS_TILE; i += BS_THREADS) tile[i] = keys[t0 + i];
    __syncthreads();
    for (unsigned size = 2; size <= BS_TILE; size <<= 1) {
      for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
        for (unsigned t = tid; t < (BS_TILE >> 1); t += BS_THREADS) {
          const unsigned i = 2 * t - (t & (stride - 1)), j = i + stride;
          const bool desc = (((t0 + i) & size) == 0);
          const uint64_t x = tile[i], y = tile[j];
          if ((x < y) == desc) { tile[i] = y; tile[j] = x; }
        }
        __syncthreads();
      }
    }
    for (unsigned i = tid; i < BS_TILE; i += BS_THREADS) keys[t0 + i] = tile[i];
    __syncthreads();
  }
  // phase 2: merge levels above the tile size; strides >= BS_TILE in global memory, the rest per tile
  for (uint64_t size = 2ull * BS_TILE; size <= cap; size <<= 1) {
    for (uint64_t stride = size >> 1; stride >= BS_TILE; stride >>= 1) {
      for (uint64_t t = tid; t < (cap >> 1); t += BS_THREADS) {
        const uint64_t i = 2 * t - (t & (stride - 1)), j = i + stride;
        const bool desc = ((i & size) == 0);
        const uint64_t x = keys[i], y = keys[j];
        if ((x < y) == desc) { keys[i] = y; keys[j] = x; }
      }
      __syncthreads();
    }
    for (uint64_t t0 = 0; t0 < cap; t0 += BS_TILE) {
      for (unsigned i = tid; i < BS_TILE; i += BS_THREADS) tile[i] = keys[t0 + i];
      __syncthreads();
      const bool desc = ((t0 & size) == 0);
      for (unsigned stride = BS_TILE >> 1; stride > 0; stride >>= 1) {
        for (unsigned t = tid; t < (BS_TILE >> 1); t += BS_THREADS) {
          const unsigned i = 2 * t - (t & (stride - 1)), j = i + stride;
          const uint64_t x = tile[i], y = tile[j];
          if ((x < y) == desc) { tile[i] = y; tile[j] = x; }
        }
        __syncthreads();
      }
      for (unsigned i = tid; i < BS_TILE; i += BS_THREADS) keys[t0 + i] = tile[i];
      __syncthreads();
    }
  }
  int64_t limit;
  if (a.p > 0.0) {
    int64_t lp = (int64_t)ceil((double)n * a.p);
    if (lp < 1) lp = 1;
    limit = (a.k > 0 && a.k < lp) ? a.k : lp;
  } else {
    limit = a.k;
  }
  if (limit > n) limit = n;
  if (limit > a.out_stride) limit = a.out_stride;
  for (int64_t i = tid; i < limit; i += BS_THREADS) {
    const uint64_t kv = keys[i];
    a.out_pos[qi * (int64_t)a.out_stride + i] = (int32_t)(0xffffffffu - (uint32_t)(kv & 0xffffffffu));
    a.out_score[qi * (int64_t)a.out_stride + i] = key_score((uint32_t)(kv >> 32));
  }
  if (tid == 0) a.out_count[qi] = (int32_t)limit;
}

// out[i] = x[i] / ||x[i]||_2 for every row; zero[i] = 1 where the norm is 0 (reference l2_norm raises).
// One warp per row, grid-stride.
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float* __restrict__ X, int64_t n, int dim, float* __restrict__ out,
                    int32_t* __restrict__ zero) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const float* x = X + row * (int64_t)dim;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
      const float v = __ldg(x + i);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float norm = sqrtf(ss);
    for (int i = lane; i < dim; i += 32) out[row * (int64_t)dim + i] = __ldg(x + i) / norm;
    if (lane == 0 && zero != nullptr) zero[row] = (norm == 0.f) ? 1 : 0;
  }
}

unsigned next_pow2(uint64_t v) {
  unsigned p = 2;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

int launch_l2_normalize(const float* d_X, int64_t n, int dim, float* d_out, int32_t* d_zero,
                        cudaStream_t stream) {
  if (n <= 0) return LSHX_OK;
  const int64_t blocks = (n + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  l2_normalize_kernel<<<grid, 256, 0, stream>>>(d_X, n, dim, d_out, d_zero);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

// Worst-case result count of a launch (the largest query decides).
static int64_t worst_limit(const RerankArgs& a, int64_t mc) {
  int64_t worst = a.k > 0 ? a.k : mc;
  if (a.p > 0.0) {
    int64_t lp = (int64_t)ceil((double)mc * a.p);
    if (lp < 1) lp = 1;
    worst = (a.k > 0 && a.k < lp) ? a.k : lp;
  }
  return worst < mc ? worst : mc;
}

static bool needs_bigsort(const RerankArgs& a) {
  return a.select && a.max_cand > RR_MAX_CAP && worst_limit(a, a.max_cand) > RR_MAX_CAP / 2;
}

size_t rerank_scratch_bytes(const RerankArgs& a, int64_t* queries_per_pass) {
  if (queries_per_pass) *queries_per_pass = a.nq;
  if (!needs_bigsort(a)) return 0;
  uint64_t stride = BS_TILE;
  while (stride < (uint64_t)a.max_cand) stride <<= 1;
  const size_t per_query = (size_t)stride * sizeof(uint64_t);
  int64_t group = (int64_t)((1ull << 30) / per_query);   // at most 1 GiB of keys in flight
  if (group < 1) group = 1;
  if (group > a.nq) group = a.nq;
  if (queries_per_pass) *queries_per_pass = group;
  return per_query * (size_t)group;
}

int launch_rerank(const RerankArgs& a_in, cudaStream_t stream) {
  RerankArgs a = a_in;
  if (a.nq <= 0) return LSHX_OK;
  const int q_floats = (a.dim + 3) & ~3;
  const bool vec4 = (a.dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.V) & 15) == 0);
  auto kern = vec4 ? rerank_kernel<true> : rerank_kernel<false>;

  if (needs_bigsort(a)) {
    int64_t group = 0;
    const size_t need = rerank_scratch_bytes(a, &group);
    LSHX_REQUIRE(a.big_keys != nullptr && a.big_bytes >= need,
                 "rerank: %lld candidates for one query need %zu bytes of sort scratch", (long long)a.max_cand, need);
    const int64_t stride = (int64_t)(need / sizeof(uint64_t) / (size_t)group);
    LSHX_REQUIRE(a.max_cand < (1ll << 32), "rerank: more than 2^32 candidates for one query");
    const size_t smem = (size_t)q_floats * sizeof(float) + 2 * sizeof(uint64_t);
    LSHX_REQUIRE(smem <= 227 * 1024, "rerank: dim %d too large for shared memory", a.dim);
    LSHX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t q0 = 0; q0 < a_in.nq; q0 += group) {
      RerankArgs c = a_in;
      c.nq = (a_in.nq - q0 < group) ? (a_in.nq - q0) : group;
      c.Q = a_in.Q + q0 * (int64_t)a_in.dim;
      c.offs = a_in.offs + q0;
      c.out_pos = a_in.out_pos + q0 * (int64_t)a_in.out_stride;
      c.out_score = a_in.out_score + q0 * (int64_t)a_in.out_stride;
      c.out_count = a_in.out_count + q0;
      c.out_zero = a_in.out_zero ? a_in.out_zero + q0 : nullptr;
      c.cand_counts = a_in.cand_counts ? a_in.cand_counts + q0 : nullptr;
      c.big_stride = stride;
      RerankArgs score = c;       // pass 1: score every candidate, keys to global memory, no in-SM selection
      score.select = false;
      kern<<<(unsigned)c.nq, RR_THREADS, smem, stream>>>(score, 2u, q_floats);
      count_launch();
      LSHX_CUDA(cudaGetLastError());
      rerank_bigsort_kernel<<<(unsigned)c.nq, BS_THREADS, 0, stream>>>(c);   // pass 2: sort + cut
      count_launch();
      LSHX_CUDA(cudaGetLastError());
    }
    return LSHX_OK;
  }

  a.big_keys = nullptr;
  unsigned cap = 2;
  if (a.select) {
    const int64_t mc = a.max_cand < 1 ? 1 : a.max_cand;
    // above RR_MAX_CAP candidates: chunked running top-k that keeps the best cap/2 keys between chunks
    cap = mc <= RR_MAX_CAP ? next_pow2((uint64_t)mc) : (unsigned)RR_MAX_CAP;
  }
  const size_t smem = (size_t)q_floats * sizeof(float) + (size_t)cap * sizeof(uint64_t);
  LSHX_REQUIRE(smem <= 227 * 1024, "rerank: dim %d too large for shared memory", a.dim);
  LSHX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(a.nq < (1 << 20) ? a.nq : (1 << 20));
  kern<<<grid, RR_THREADS, smem, stream>>>(a, cap, q_floats);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

}  // namespace lshx
