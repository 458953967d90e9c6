// Device-side band index and candidate join (sm_100a) -- SURVEY section 8f rank 4.
//
// Replaces, for batches of queries, the bucket lookups and collision counting of
// LSHRS._candidate_counts (reference lshrs/core/main.py:1088-1111: for every band, SMEMBERS of the
// bucket named by the query's band key; counts[id] += 1) and the ordering of LSHRS.query
// (main.py:614: sorted by (-collisions, id)).  The reference keeps buckets in Redis SETs and walks them
// in Python; here the (band, key, id) triples that index() produces stay in HBM:
//
//   entries   per band b a segment keys[b][0..n) / ids[b][0..n), sorted by (key, id); key = the band's
//             bytes as a little-endian integer (<= 8 bytes per band, i.e. rows_per_band <= 64)
//   add       append to the segments' tails; the next query radix-sorts the segments again
//             (LSD, 8-bit digits, stable; only the bytes that max(id) and the key width need)
//   remove    ids are overwritten by a tombstone (-1) in place (Redis SREM from every bucket)
//   query     lookup  : binary search of every (query, band) key -> [lo, lo + cnt)
//             scan    : per-query raw candidate counts -> offsets, total, maximum
//             join    : one CTA per query gathers its <= num_bands ranges, sorts the ids, run-length
//                       counts them (= collisions; duplicates inside one bucket and tombstones are
//                       skipped, SET semantics), sorts by (-collisions, id) and writes the list
//   the lists feed lshx_rerank_topk on the device (candidate id = row of the resident corpus).
//
// Integer work end to end: the candidate lists, their order and the collision counts are EXACTLY the
// storage path's.  HBM / L2-latency bound; nothing here belongs on tensor cores.

#include <cstring>

#include "lshx_common.cuh"
#include "hash_small.cuh"

namespace lshx {
namespace {

constexpr uint64_t EMPTY = ~0ull;            // sorts last
constexpr int ID_BITS = 56;                  // id < 2^56; collisions ride in the top byte of the order key
constexpr uint64_t ID_MASK = (1ull << ID_BITS) - 1;

// ---------------------------------------------------------------------------------------------------
// append
// ---------------------------------------------------------------------------------------------------
// ids: one per vector, or (per_band) one per (vector, band) with -1 = "no entry in this band" (stored as a
// tombstone, which no query ever returns): how single (band, key, id) operations reach the equal-length segments
__global__ void index_append_kernel(const uint8_t* __restrict__ sig, const int64_t* __restrict__ ids, int64_t n,
                                    int nb, int bpb, uint64_t* __restrict__ keys, int64_t* __restrict__ out_ids,
                                    int64_t cap, int64_t at, unsigned long long* __restrict__ max_id,
                                    int* __restrict__ bad, int per_band) {
  const int64_t total = n * nb;
  unsigned long long local_max = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / nb;
    const int b = (int)(t % nb);
    const uint8_t* src = sig + (i * nb + b) * (int64_t)bpb;
    uint64_t key = 0;
    for (int j = 0; j < bpb; ++j) key |= (uint64_t)src[j] << (8 * j);
    const int64_t id = per_band ? ids[t] : ids[i];
    if ((id < 0 && !(per_band && id == -1)) || (id >= 0 && (uint64_t)id > ID_MASK)) *bad = 1;
    keys[b * cap + at + i] = key;
    out_ids[b * cap + at + i] = id;
    if (id > 0 && (unsigned long long)id > local_max) local_max = (unsigned long long)id;
  }
  if (local_max) atomicMax(max_id, local_max);
}

// ---------------------------------------------------------------------------------------------------
// LSD radix sort of every band segment by (key, id); grid = (tiles, bands)
// ---------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__device__ __forceinline__ unsigned digit_of(uint64_t key, int64_t id, int pass_byte, int id_bytes) {
  // passes [0, id_bytes) walk the id's bytes, the following ones the key's
  return (pass_byte < id_bytes) ? (unsigned)(((uint64_t)id >> (8 * pass_byte)) & 0xFFu)
                                : (unsigned)((key >> (8 * (pass_byte - id_bytes))) & 0xFFu);
}

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ ids, int64_t n, int64_t cap,
                  int pass_byte, int id_bytes, unsigned* __restrict__ hist, int ntiles) {
  __shared__ unsigned h[256];
  const int band = blockIdx.y, tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)tile * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t e = base + r * RS_THREADS + threadIdx.x;
    if (e < n) atomicAdd(&h[digit_of(keys[band * cap + e], ids[band * cap + e], pass_byte, id_bytes)], 1u);
  }
  __syncthreads();
  hist[((int64_t)band * 256 + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// exclusive scan of each band's 256 * ntiles counters (digit-major), one CTA per band
__global__ void __launch_bounds__(1024)
radix_scan_kernel(unsigned* __restrict__ hist, int ntiles) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
  unsigned* h = hist + (int64_t)blockIdx.x * 256 * ntiles;
  const int64_t total = (int64_t)256 * ntiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t c0 = 0; c0 < total; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const unsigned v = (i < total) ? h[i] : 0u;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;   // inclusive totals of the warps
    }
    __syncthreads();
    const unsigned before = carry_s + (warp ? warp_tot[warp - 1] : 0u) + inc - v;
    if (i < total) h[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const int64_t* __restrict__ ids_in,
                     uint64_t* __restrict__ keys_out, int64_t* __restrict__ ids_out, int64_t n, int64_t cap,
                     int pass_byte, int id_bytes, const unsigned* __restrict__ hist, int ntiles) {
  __shared__ unsigned base[256];
  __shared__ unsigned warp_cnt[RS_THREADS / 32][256];
  const int band = blockIdx.y, tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  base[threadIdx.x] = hist[((int64_t)band * 256 + threadIdx.x) * ntiles + tile];
  const int64_t tile0 = (int64_t)tile * RS_TILE;
  for (int r = 0; r < RS_ITEMS; ++r) {
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t e = tile0 + r * RS_THREADS + threadIdx.x;
    const bool valid = e < n;
    uint64_t key = 0;
    int64_t id = 0;
    unsigned d = 256u + (unsigned)lane;     // invalid lanes never match a real digit (or each other)
    if (valid) {
      key = keys_in[band * cap + e];
      id = ids_in[band * cap + e];
      d = digit_of(key, id, pass_byte, id_bytes);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[warp][d] = __popc(peers);
    __syncthreads();
    unsigned pos = 0;
    if (valid) {
      pos = base[d] + rank;
      for (int w = 0; w < warp; ++w) pos += warp_cnt[w][d];
    }
    __syncthreads();
    {
      unsigned add = 0;
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; ++w) add += warp_cnt[w][threadIdx.x];
      base[threadIdx.x] += add;
    }
    if (valid) {
      keys_out[band * cap + pos] = key;
      ids_out[band * cap + pos] = id;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// remove: tombstone every entry whose id is in the (sorted) removal list
// ---------------------------------------------------------------------------------------------------
__global__ void index_tombstone_kernel(int64_t* __restrict__ ids, int64_t n, int64_t cap, int nb,
                                       const int64_t* __restrict__ gone, int64_t ngone) {
  const int64_t total = n * nb;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / n);
    const int64_t e = t % n;
    const int64_t id = ids[b * cap + e];
    if (id < 0) continue;
    int64_t lo = 0, hi = ngone;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (gone[mid] < id) lo = mid + 1; else hi = mid;
    }
    if (lo < ngone && gone[lo] == id) ids[b * cap + e] = -1;
  }
}

// ---------------------------------------------------------------------------------------------------
// query
// ---------------------------------------------------------------------------------------------------
// Runs: the main run [0, main_n) and, when main_n < n, the delta run [main_n, n) of every band segment are each in
// (key, id) order (a small add is sorted on its own instead of re-sorting the whole segment).  A query searches both:
// slot v = run * nb + band of query q -> lo (absolute entry index), cnt.
__device__ __forceinline__ void search_run(const uint64_t* __restrict__ k, int64_t begin, int64_t end, uint64_t key,
                                           int64_t& lo_out, int64_t& cnt_out) {
  int64_t lo = begin, hi = end;
  while (lo < hi) {                       // first entry with k >= key
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(k + mid) < key) lo = mid + 1; else hi = mid;
  }
  int64_t lo2 = lo, hi2 = end;
  while (lo2 < hi2) {                     // first entry with k > key
    const int64_t mid = (lo2 + hi2) >> 1;
    if (__ldg(k + mid) <= key) lo2 = mid + 1; else hi2 = mid;
  }
  lo_out = lo;
  cnt_out = lo2 - lo;
}

// The same search by a GROUP of W lanes (latency path: one query, nothing else in flight to hide a 20-step chain
// of dependent loads behind): W probes per step split the range W + 1 ways -- 5 steps for a million entries with
// W = 16 -- and a bucket shorter than W entries ends in one more.  Two groups share a warp (each ballots under
// its own mask), so the 16 slots of a 16-band query are searched in ONE round of the CTA's 8 warps.
// LE = false: first entry >= key; true: first entry > key.
template <bool LE, int W>
__device__ __forceinline__ int64_t group_bound(const uint64_t* __restrict__ k, int64_t lo, int64_t hi, uint64_t key,
                                               int gl /* lane within the group */, unsigned mask) {
  const int shift = __ffs(mask) - 1;
  while (hi > lo) {
    const int64_t len = hi - lo;
    const int64_t p = (len <= W) ? lo + gl : lo + (len * (gl + 1)) / (W + 1);   // ascending in the lane, p < hi
    bool before = false;
    if (p < hi) {
      const uint64_t v = __ldg(k + p);
      before = LE ? (v <= key) : (v < key);
    }
    const int t = __popc((__ballot_sync(mask, before) & mask) >> shift);   // sorted: the first t probes are "before"
    if (len <= W) return lo + t;
    const int64_t p_prev = lo + (len * t) / (W + 1);                 // probe t - 1 (for t > 0)
    const int64_t p_next = lo + (len * (t + 1)) / (W + 1);           // probe t     (for t < W)
    const int64_t nlo = t > 0 ? p_prev + 1 : lo;
    hi = t < W ? p_next : hi;
    lo = nlo;
  }
  return lo;
}

__global__ void index_lookup_kernel(const uint8_t* __restrict__ sig, int64_t nq, int nb, int bpb,
                                    const uint64_t* __restrict__ keys, int64_t main_n, int64_t n, int64_t cap,
                                    int64_t* __restrict__ lo_out, int* __restrict__ cnt_out,
                                    int* __restrict__ raw_count) {
  const int nruns = n > main_n ? 2 : 1;
  const int nv = nb * nruns;
  const int64_t total = nq * nv;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / nv;
    const int v = (int)(t % nv);
    const int b = v % nb, run = v / nb;
    const uint8_t* src = sig + (q * nb + b) * (int64_t)bpb;
    uint64_t key = 0;
    for (int j = 0; j < bpb; ++j) key |= (uint64_t)src[j] << (8 * j);
    int64_t lo, c;
    search_run(keys + b * cap, run ? main_n : 0, run ? n : main_n, key, lo, c);
    lo_out[t] = lo;
    cnt_out[t] = (int)(c > 0x7fffffff ? 0x7fffffff : c);
    if (c) atomicAdd(raw_count + q, (int)c);
  }
}

__device__ __forceinline__ unsigned pow2_at_least(unsigned v) {
  unsigned p = 2;
  while (p < v) p <<= 1;
  return p;
}

// exclusive scans of raw_count (candidate slots) and of pow2(raw_count) (sort workspace slots);
// meta = {total raw, max raw, total workspace}
__global__ void __launch_bounds__(1024)
index_scan_kernel(const int* __restrict__ raw_count, int64_t nq, int64_t* __restrict__ raw_off,
                  int64_t* __restrict__ ws_off, int64_t* __restrict__ meta) {
  __shared__ long long wt_a[32], wt_b[32];
  __shared__ long long carry_a, carry_b;
  __shared__ int max_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; max_s = 0; }
  __syncthreads();
  for (int64_t c0 = 0; c0 < nq; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const int v = (i < nq) ? raw_count[i] : 0;
    const long long a = v, b = v > 0 ? (long long)pow2_at_least((unsigned)v) : 0;
    long long ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { wt_a[warp] = ia; wt_b[warp] = ib; }
    atomicMax(&max_s, v);
    __syncthreads();
    if (warp == 0) {
      long long wa = wt_a[lane], wb = wt_b[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long ta = __shfl_up_sync(0xffffffffu, wa, o), tb = __shfl_up_sync(0xffffffffu, wb, o);
        if (lane >= o) { wa += ta; wb += tb; }
      }
      wt_a[lane] = wa; wt_b[lane] = wb;
    }
    __syncthreads();
    const long long ea = carry_a + (warp ? wt_a[warp - 1] : 0) + ia - a;
    const long long eb = carry_b + (warp ? wt_b[warp - 1] : 0) + ib - b;
    if (i < nq) { raw_off[i] = ea; ws_off[i] = eb; }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_a = ea + a; carry_b = eb + b; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    raw_off[nq] = carry_a;
    ws_off[nq] = carry_b;
    meta[0] = carry_a;
    meta[1] = max_s;
    meta[2] = carry_b;
  }
}

constexpr int JN_THREADS = 256;
constexpr unsigned JN_SMEM_CAP = 4096;   // raw candidates of one query sorted in shared memory (2 x 32 KB)

__device__ __forceinline__ void bitonic_asc(uint64_t* a, unsigned P, int tid) {
  if (P <= 64) {
    // a short list (the usual case of a single query): one warp sorts it between two barriers instead of the
    // whole CTA meeting at a barrier after each of the network's 21 stages
    if (tid < 32) {
      for (unsigned size = 2; size <= P; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
          const unsigned t = (unsigned)tid;
          if (t < (P >> 1)) {
            const unsigned i = 2 * t - (t & (stride - 1)), j = i + stride;
            const bool asc = ((i & size) == 0);
            const uint64_t x = a[i], y = a[j];
            if ((x > y) == asc) { a[i] = y; a[j] = x; }
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    return;
  }
  for (unsigned size = 2; size <= P; size <<= 1) {
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned t = tid; t < (P >> 1); t += JN_THREADS) {
        const unsigned i = 2 * t - (t & (stride - 1)), j = i + stride;
        const bool asc = ((i & size) == 0);
        const uint64_t x = a[i], y = a[j];
        if ((x > y) == asc) { a[i] = y; a[j] = x; }
      }
      __syncthreads();
    }
  }
}

struct JoinArgs {
  int64_t nq;
  int nb;                  // bands
  int nv;                  // slots per query: nb (one run) or 2 * nb (main + delta run)
  const int64_t* ids;      // [nb][cap]
  int64_t cap;
  const int64_t* lo;       // [nq][nb]
  const int* cnt;          // [nq][nb]
  const int* raw_count;    // [nq]
  const int64_t* raw_off;  // [nq + 1]
  const int64_t* ws_off;   // [nq + 1] (global workspace mode)
  uint64_t* ws;            // global workspace, 2 x total_ws entries (nullptr: shared memory)
  int64_t ws_total;
  int64_t* out_ids;        // candidate ids at raw_off[q] .. + uniq[q], ordered by (-collisions, id)
  int* out_coll;           // their collision counts (may be null)
  int* uniq;               // [nq]
};

template <bool SMEM>
__global__ void __launch_bounds__(JN_THREADS)
index_join_kernel(JoinArgs a) {
  extern __shared__ __align__(16) uint64_t sm[];
  __shared__ int band_off[512];
  __shared__ int heads;
  const int tid = threadIdx.x;
  for (int64_t q = blockIdx.x; q < a.nq; q += gridDim.x) {
    const int n_raw = a.raw_count[q];
    if (n_raw == 0) {
      if (tid == 0) a.uniq[q] = 0;
      continue;
    }
    const unsigned P = pow2_at_least((unsigned)n_raw);
    uint64_t* buf = SMEM ? sm : a.ws + a.ws_off[q];
    uint64_t* buf2 = SMEM ? sm + JN_SMEM_CAP : a.ws + a.ws_total + a.ws_off[q];
    if (tid == 0) {
      int acc = 0;
      for (int v = 0; v < a.nv; ++v) { band_off[v] = acc; acc += a.cnt[q * a.nv + v]; }
      heads = 0;
    }
    __syncthreads();
    // gather (id, band) pairs of every matching bucket: after the sort an id repeated inside one bucket (indexed
    // twice with the same band key -- in one run or once in each) is a repeated PAIR and counts once, like a
    // Redis SET member; a tombstone not at all
    for (int v = 0; v < a.nv; ++v) {
      const int b = v % a.nb;
      const int c = a.cnt[q * a.nv + v];
      const int64_t* src = a.ids + b * a.cap + a.lo[q * a.nv + v];
      for (int i = tid; i < c; i += JN_THREADS) {
        const int64_t id = src[i];
        buf[band_off[v] + i] = id < 0 ? EMPTY : (((uint64_t)id << 8) | (uint64_t)b);
      }
    }
    for (unsigned i = n_raw + tid; i < P; i += JN_THREADS) buf[i] = EMPTY;
    __syncthreads();
    bitonic_asc(buf, P, tid);
    // distinct bands per id = collisions; order key = (255 - collisions) << 56 | id  (ascending = (-collisions, id))
    int mine = 0;
    for (unsigned i = tid; i < P; i += JN_THREADS) {
      const uint64_t v = buf[i];
      uint64_t key2 = EMPTY;
      if (v != EMPTY && (i == 0 || (buf[i - 1] >> 8) != (v >> 8))) {
        unsigned coll = 1;
        for (unsigned j = i + 1; j < P && (buf[j] >> 8) == (v >> 8); ++j) coll += (buf[j] != buf[j - 1]) ? 1u : 0u;
        key2 = ((uint64_t)(255u - coll) << ID_BITS) | (v >> 8);
        ++mine;
      }
      buf2[i] = key2;
    }
    if (mine) atomicAdd(&heads, mine);
    __syncthreads();
    bitonic_asc(buf2, P, tid);
    const int u = heads;
    const int64_t o = a.raw_off[q];
    for (int i = tid; i < u; i += JN_THREADS) {
      const uint64_t k2 = buf2[i];
      a.out_ids[o + i] = (int64_t)(k2 & ID_MASK);
      if (a.out_coll) a.out_coll[o + i] = 255 - (int)(k2 >> ID_BITS);
    }
    if (tid == 0) a.uniq[q] = u;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// latency path: lookup + join + emit of a FEW queries in one launch (LSHRS.query / get_top_k hash ONE vector per
// call, main.py:524-658).  One CTA per query; thread b searches band b; the lists go straight into mapped pinned
// host memory (no copy call), count = -1 when the raw candidates do not fit the shared-memory sort (the caller
// then takes the batched path).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// dbg (diagnostics, normally null): nanosecond stamps of the phases of query 0, see lshx_index_debug_timeline
__device__ __forceinline__ void query_small_body(unsigned long long* dbg, uint64_t* sm, const int64_t q, const uint8_t* sig, int nb, int bpb, const uint64_t* __restrict__ keys,
                         const int64_t* __restrict__ ids, int64_t main_n, int64_t n, int64_t cap, int out_cap,
                         int raw_cap, int64_t* __restrict__ out_ids, int* __restrict__ out_coll,
                         int* __restrict__ out_count, int* __restrict__ out_count_clamped,
                         int64_t* __restrict__ out_offs) {
  __shared__ int64_t s_lo[512];
  __shared__ int s_cnt[512];
  __shared__ int band_off[512];
  __shared__ int heads, n_raw_s;
  const int tid = threadIdx.x;
  const int nv = nb * (n > main_n ? 2 : 1);
  constexpr int GW = 16;                                        // lanes per search group
  const int gl = tid & (GW - 1);
  const unsigned gmask = ((1u << GW) - 1u) << ((tid & 31) & ~(GW - 1));
  for (int v = tid / GW; v < nv; v += JN_THREADS / GW) {        // one group per (run, band) slot
    const int b = v % nb, run = v / nb;
    const uint8_t* src = sig + (q * nb + b) * (int64_t)bpb;
    uint64_t key = 0;
    for (int j = 0; j < bpb; ++j) key |= (uint64_t)__ldcg(src + j) << (8 * j);
    const uint64_t* k = keys + b * cap;
    const int64_t end = run ? n : main_n;
    const int64_t lo = group_bound<false, GW>(k, run ? main_n : 0, end, key, gl, gmask);
    // most buckets are short: one probe of the GW entries behind lo usually finds the end
    const int64_t near = (end - lo < GW) ? end : lo + GW;
    int64_t up = group_bound<true, GW>(k, lo, near, key, gl, gmask);
    if (up == near && near < end) up = group_bound<true, GW>(k, near, end, key, gl, gmask);
    if (gl == 0) {
      const int64_t c = up - lo;
      s_lo[v] = lo;
      s_cnt[v] = (int)(c > 0x7fffffff ? 0x7fffffff : c);
    }
  }
  __syncthreads();
  if (dbg && tid == 0 && q == 0) dbg[3] = gtime();     // searches done
  if (tid == 0) {
    long long acc = 0;
    for (int v = 0; v < nv; ++v) { band_off[v] = (int)(acc > 0x7fffffff ? 0x7fffffff : acc); acc += s_cnt[v]; }
    n_raw_s = acc > (long long)raw_cap ? -1 : (int)acc;
    heads = 0;
    if (out_offs) out_offs[q] = q * (int64_t)out_cap;       // CSR base of this query's slots for the rerank kernel
  }
  __syncthreads();
  const int n_raw = n_raw_s;
  if (n_raw <= 0) {
    if (tid == 0) {
      out_count[q] = n_raw;     // 0 = no candidate, -1 = too many for this path
      if (out_count_clamped) out_count_clamped[q] = 0;
    }
    return;
  }
  const unsigned P = pow2_at_least((unsigned)n_raw);
  uint64_t* buf = sm;
  uint64_t* buf2 = sm + JN_SMEM_CAP;
  // gather, flattened over the matched entries: every load is in flight at once (a loop over the slots costs one
  // memory round trip per slot -- 6.5 us of a 27 us kernel, lshx_index_debug_timeline)
  for (int e = tid; e < n_raw; e += JN_THREADS) {
    int lo_v = 0, hi_v = nv - 1;          // last slot whose first entry is <= e (empty slots share an offset)
    while (lo_v < hi_v) {
      const int mid = (lo_v + hi_v + 1) >> 1;
      if (band_off[mid] <= e) lo_v = mid; else hi_v = mid - 1;
    }
    const int b = lo_v % nb;
    const int64_t id = ids[b * cap + s_lo[lo_v] + (e - band_off[lo_v])];
    buf[e] = id < 0 ? EMPTY : (((uint64_t)id << 8) | (uint64_t)b);   // (id, band), as in the join
  }
  for (unsigned i = n_raw + tid; i < P; i += JN_THREADS) buf[i] = EMPTY;
  __syncthreads();
  if (dbg && tid == 0 && q == 0) dbg[4] = gtime();     // gathered
  bitonic_asc(buf, P, tid);
  if (dbg && tid == 0 && q == 0) dbg[5] = gtime();     // first sort
  int mine = 0;
  for (unsigned i = tid; i < P; i += JN_THREADS) {
    const uint64_t v = buf[i];
    uint64_t key2 = EMPTY;
    if (v != EMPTY && (i == 0 || (buf[i - 1] >> 8) != (v >> 8))) {
      unsigned coll = 1;
      for (unsigned j = i + 1; j < P && (buf[j] >> 8) == (v >> 8); ++j) coll += (buf[j] != buf[j - 1]) ? 1u : 0u;
      key2 = ((uint64_t)(255u - coll) << ID_BITS) | (v >> 8);
      ++mine;
    }
    buf2[i] = key2;
  }
  if (mine) atomicAdd(&heads, mine);
  __syncthreads();
  bitonic_asc(buf2, P, tid);
  if (dbg && tid == 0 && q == 0) dbg[6] = gtime();     // counted + second sort
  const int u = heads;
  const int take = u < out_cap ? u : out_cap;
  for (int i = tid; i < take; i += JN_THREADS) {
    const uint64_t k2 = buf2[i];
    out_ids[q * out_cap + i] = (int64_t)(k2 & ID_MASK);
    if (out_coll) out_coll[q * out_cap + i] = 255 - (int)(k2 >> ID_BITS);
  }
  if (tid == 0) {
    out_count[q] = u;
    if (out_count_clamped) out_count_clamped[q] = take;
    if (dbg && q == 0) { dbg[7] = gtime(); dbg[8] = (unsigned long long)n_raw; }
  }
}


__global__ void __launch_bounds__(JN_THREADS)
index_query_small_kernel(const uint8_t* __restrict__ sig, int nb, int bpb, const uint64_t* __restrict__ keys,
                         const int64_t* __restrict__ ids, int64_t main_n, int64_t n, int64_t cap, int out_cap,
                         int raw_cap, int64_t* __restrict__ out_ids, int* __restrict__ out_coll,
                         int* __restrict__ out_count, int* __restrict__ out_count_clamped,
                         int64_t* __restrict__ out_offs) {
  extern __shared__ __align__(16) uint64_t sm[];
  query_small_body(nullptr, sm, blockIdx.x, sig, nb, bpb, keys, ids, main_n, n, cap, out_cap, raw_cap, out_ids, out_coll, out_count,
                   out_count_clamped, out_offs);
}

// Hash AND query in one launch: one CTA per signature byte hashes the nq vectors (hash_small_body; X may be mapped
// pinned host memory, so no copy precedes the launch); the CTA that finishes last -- a ticket counter, no CTA
// ever waits for another -- reads the finished signatures and runs the lookup / join / emit of every query.
static_assert(SMALL_THREADS == JN_THREADS, "the fused latency kernel runs both bodies with one block size");

struct HashQueryArgs {
  int nq, dim;
  const float* Rp;
  uint8_t* sig;
  int sig_bytes;
  uint8_t* zero_flag;
  unsigned* ticket;
  int nb, bpb;
  const uint64_t* keys;
  const int64_t* ids;
  int64_t main_n, n, cap;
  int out_cap, raw_cap;
  int64_t* out_ids;
  int* out_coll;
  int* out_count;
  int* out_count_clamped;
  int64_t* out_offs;
  unsigned long long* dbg;
};

__device__ __forceinline__ void hash_query_small(const float* X, const HashQueryArgs& a, uint64_t* sm) {
  __shared__ unsigned int sbits[32];
  __shared__ int last;
  unsigned long long t_in = 0;
  if (a.dbg && threadIdx.x == 0) t_in = gtime();
  hash_small_body(X, a.nq, a.dim, a.Rp, a.sig, a.sig_bytes, a.zero_flag, reinterpret_cast<float*>(sm), sbits);
  unsigned long long t_hashed = 0;
  if (a.dbg && threadIdx.x == 0) t_hashed = gtime();
  __threadfence();                      // this CTA's signature bytes are visible before its ticket is
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(a.ticket, 1u);
    last = (t == gridDim.x - 1) ? 1 : 0;
    if (last) *a.ticket = 0u;           // ready for the next launch (launches on one stream do not overlap)
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (a.dbg && threadIdx.x == 0) { a.dbg[0] = t_in; a.dbg[1] = t_hashed; a.dbg[2] = gtime(); }   // last CTA: in, hashed, ticket
  for (int q = 0; q < a.nq; ++q) {
    query_small_body(a.dbg, sm, q, a.sig, a.nb, a.bpb, a.keys, a.ids, a.main_n, a.n, a.cap, a.out_cap, a.raw_cap,
                     a.out_ids, a.out_coll, a.out_count, a.out_count_clamped, a.out_offs);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(JN_THREADS)
index_hash_query_small_kernel(const float* __restrict__ X, const HashQueryArgs a) {
  extern __shared__ __align__(16) uint64_t sm[];
  hash_query_small(X, a, sm);
}

// ONE vector travels in the kernel's parameter block (XRowParam, hash_small.cuh)
__global__ void __launch_bounds__(JN_THREADS)
index_hash_query_one_kernel(const __grid_constant__ XRowParam x, const HashQueryArgs a) {
  extern __shared__ __align__(16) uint64_t sm[];
  hash_query_small(x.v, a, sm);
}

// dense [nq][k] prefix of the candidate lists (get_top_k mode, main.py:616-623)
__global__ void index_topk_kernel(const int64_t* __restrict__ cand, const int64_t* __restrict__ raw_off,
                                  const int* __restrict__ uniq, int64_t nq, int k, int64_t* __restrict__ out,
                                  int* __restrict__ out_count) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nq * k; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / k;
    const int i = (int)(t % k);
    const int u = uniq[q];
    out[t] = (i < u) ? cand[raw_off[q] + i] : -1;
    if (i == 0) out_count[q] = u < k ? u : k;
  }
}

// rerank positions -> candidate ids
__global__ void index_pos_to_id_kernel(const int64_t* __restrict__ cand, const int64_t* __restrict__ raw_off,
                                       const int32_t* __restrict__ pos, const int32_t* __restrict__ count, int64_t nq,
                                       int stride, int64_t* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nq * stride; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / stride;
    const int i = (int)(t % stride);
    out[t] = (i < count[q]) ? cand[raw_off[q] + pos[t]] : -1;
  }
}

// ---------------------------------------------------------------------------------------------------
// single buckets (the storage protocol's get_bucket, reference lshrs/storage/redis.py get_bucket = SMEMBERS)
// ---------------------------------------------------------------------------------------------------
__global__ void index_bucket_lookup_kernel(const int* __restrict__ band_ids, const uint64_t* __restrict__ want, int64_t m,
                                           const uint64_t* __restrict__ keys, int64_t n, int64_t cap,
                                           int64_t* __restrict__ lo_out, int64_t* __restrict__ cnt_out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = want[t];
    const uint64_t* k = keys + band_ids[t] * cap;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(k + mid) < key) lo = mid + 1; else hi = mid;
    }
    int64_t lo2 = lo, hi2 = n;
    while (lo2 < hi2) {
      const int64_t mid = (lo2 + hi2) >> 1;
      if (__ldg(k + mid) <= key) lo2 = mid + 1; else hi2 = mid;
    }
    lo_out[t] = lo;
    cnt_out[t] = lo2 - lo;
  }
}

// one CTA per bucket: its id range, as stored (ascending ids, duplicates adjacent, tombstones last), to out + off[t]
__global__ void index_bucket_gather_kernel(const int* __restrict__ band_ids, const int64_t* __restrict__ lo,
                                           const int64_t* __restrict__ off, int64_t m, const int64_t* __restrict__ ids,
                                           int64_t cap, int64_t* __restrict__ out) {
  for (int64_t t = blockIdx.x; t < m; t += gridDim.x) {
    const int64_t c = off[t + 1] - off[t];
    const int64_t* src = ids + band_ids[t] * cap + lo[t];
    int64_t* dst = out + off[t];
    for (int64_t i = threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
  }
}

unsigned grid_for(int64_t work, int threads) {
  int64_t g = (work + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)g;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------------
int index_append(const uint8_t* d_sig, const int64_t* d_ids, int64_t n, int nb, int bpb, uint64_t* keys, int64_t* ids,
                 int64_t cap, int64_t at, unsigned long long* d_max_id, int* d_bad, int per_band, cudaStream_t st) {
  if (n <= 0) return LSHX_OK;
  index_append_kernel<<<grid_for(n * nb, 256), 256, 0, st>>>(d_sig, d_ids, n, nb, bpb, keys, ids, cap, at, d_max_id, d_bad,
                                                             per_band);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

// Sorts entries [first, first + n) of every band segment by (key, id); *cur names the buffer that holds them and is
// flipped once per pass -- the caller moves the range back when it must stay beside entries that were not sorted.
int index_sort(uint64_t* keys_in[2], int64_t* ids_in[2], int* cur, int64_t first, int64_t n, int64_t cap, int nb,
               int key_bytes, int id_bytes, unsigned* d_hist, size_t hist_entries, cudaStream_t st) {
  if (n <= 1) return LSHX_OK;
  uint64_t* keys[2] = {keys_in[0] + first, keys_in[1] + first};
  int64_t* ids[2] = {ids_in[0] + first, ids_in[1] + first};
  const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  LSHX_REQUIRE((size_t)nb * 256 * ntiles <= hist_entries, "radix histogram scratch too small");
  LSHX_REQUIRE(n < (1ll << 32), "more than 2^32 entries per band are not supported");
  dim3 grid((unsigned)ntiles, (unsigned)nb);
  for (int pass = 0; pass < id_bytes + key_bytes; ++pass) {
    const int in = *cur, out = in ^ 1;
    radix_hist_kernel<<<grid, RS_THREADS, 0, st>>>(keys[in], ids[in], n, cap, pass, id_bytes, d_hist, ntiles);
    radix_scan_kernel<<<(unsigned)nb, 1024, 0, st>>>(d_hist, ntiles);
    radix_scatter_kernel<<<grid, RS_THREADS, 0, st>>>(keys[in], ids[in], keys[out], ids[out], n, cap, pass, id_bytes,
                                                      d_hist, ntiles);
    count_launch(3);
    LSHX_CUDA(cudaGetLastError());
    *cur = out;
  }
  return LSHX_OK;
}
size_t index_sort_hist_entries(int64_t n, int nb) {
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  return (size_t)nb * 256 * (size_t)(ntiles < 1 ? 1 : ntiles);
}

int index_tombstone(int64_t* ids, int64_t n, int64_t cap, int nb, const int64_t* d_gone_sorted, int64_t ngone,
                    cudaStream_t st) {
  if (n <= 0 || ngone <= 0) return LSHX_OK;
  index_tombstone_kernel<<<grid_for(n * nb, 256), 256, 0, st>>>(ids, n, cap, nb, d_gone_sorted, ngone);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_lookup_scan(const uint8_t* d_sig, int64_t nq, int nb, int bpb, const uint64_t* keys, int64_t main_n, int64_t n,
                      int64_t cap, int64_t* d_lo, int* d_cnt, int* d_raw_count, int64_t* d_raw_off, int64_t* d_ws_off,
                      int64_t* d_meta, cudaStream_t st) {
  LSHX_CUDA(cudaMemsetAsync(d_raw_count, 0, (size_t)nq * sizeof(int), st));
  index_lookup_kernel<<<grid_for(nq * nb * (n > main_n ? 2 : 1), 256), 256, 0, st>>>(d_sig, nq, nb, bpb, keys, main_n, n,
                                                                                       cap, d_lo, d_cnt, d_raw_count);
  index_scan_kernel<<<1, 1024, 0, st>>>(d_raw_count, nq, d_raw_off, d_ws_off, d_meta);
  count_launch(2);
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_bucket_lookup(const int* d_band_ids, const uint64_t* d_want, int64_t m, const uint64_t* keys, int64_t n,
                        int64_t cap, int64_t* d_lo, int64_t* d_cnt, cudaStream_t st) {
  if (m <= 0) return LSHX_OK;
  index_bucket_lookup_kernel<<<grid_for(m, 128), 128, 0, st>>>(d_band_ids, d_want, m, keys, n, cap, d_lo, d_cnt);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_bucket_gather(const int* d_band_ids, const int64_t* d_lo, const int64_t* d_off, int64_t m, const int64_t* ids,
                        int64_t cap, int64_t* d_out, cudaStream_t st) {
  if (m <= 0) return LSHX_OK;
  index_bucket_gather_kernel<<<(unsigned)(m < 148 * 8 ? m : 148 * 8), 256, 0, st>>>(d_band_ids, d_lo, d_off, m, ids, cap,
                                                                                    d_out);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

unsigned index_join_smem_cap() { return JN_SMEM_CAP; }

int index_join(int64_t nq, int nb, int nruns, const int64_t* ids, int64_t cap, const int64_t* d_lo, const int* d_cnt,
               const int* d_raw_count, const int64_t* d_raw_off, const int64_t* d_ws_off, uint64_t* d_ws,
               int64_t ws_total, int64_t* d_out_ids, int* d_out_coll, int* d_uniq, cudaStream_t st) {
  if (nq <= 0) return LSHX_OK;
  JoinArgs a{nq, nb, nb * nruns, ids, cap, d_lo, d_cnt, d_raw_count, d_raw_off, d_ws_off, d_ws, ws_total, d_out_ids, d_out_coll,
             d_uniq};
  const unsigned grid = (unsigned)(nq < 148 * 8 ? nq : 148 * 8);
  if (d_ws == nullptr) {
    const size_t smem = 2 * (size_t)JN_SMEM_CAP * sizeof(uint64_t);
    LSHX_CUDA(cudaFuncSetAttribute(index_join_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    index_join_kernel<true><<<grid, JN_THREADS, smem, st>>>(a);
  } else {
    index_join_kernel<false><<<grid, JN_THREADS, 0, st>>>(a);
  }
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_query_small(const uint8_t* d_sig, int nq, int nb, int bpb, const uint64_t* keys, const int64_t* ids,
                      int64_t main_n, int64_t n, int64_t cap, int out_cap, int raw_cap, int64_t* out_ids, int* out_coll, int* out_count,
                      int* out_count_clamped, int64_t* out_offs, cudaStream_t st) {
  if (nq <= 0) return LSHX_OK;
  if (raw_cap <= 0 || raw_cap > (int)JN_SMEM_CAP) raw_cap = (int)JN_SMEM_CAP;
  const size_t smem = 2 * (size_t)JN_SMEM_CAP * sizeof(uint64_t);
  LSHX_CUDA(cudaFuncSetAttribute(index_query_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  index_query_small_kernel<<<(unsigned)nq, JN_THREADS, smem, st>>>(d_sig, nb, bpb, keys, ids, main_n, n, cap, out_cap,
                                                                  raw_cap, out_ids, out_coll, out_count, out_count_clamped,
                                                                  out_offs);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_hash_query_small(const float* X, const float* X_host, int nq, int dim, const float* d_Rp, uint8_t* d_sig,
                           int sig_bytes, uint8_t* zero_flag, unsigned* d_ticket, int nb, int bpb, const uint64_t* keys,
                           const int64_t* ids, int64_t main_n, int64_t n, int64_t cap, int out_cap, int raw_cap,
                           int64_t* out_ids, int* out_coll, int* out_count, int* out_count_clamped, int64_t* out_offs,
                           unsigned long long* d_dbg, cudaStream_t st) {
  if (nq <= 0) return LSHX_OK;
  if (raw_cap <= 0 || raw_cap > (int)JN_SMEM_CAP) raw_cap = (int)JN_SMEM_CAP;
  const size_t smem = 2 * (size_t)JN_SMEM_CAP * sizeof(uint64_t);     // >= nq * dim floats (hash_small_max_rows)
  LSHX_REQUIRE((size_t)nq * dim * sizeof(float) <= smem, "too many rows for the fused latency kernel");
  static thread_local int attr_device = -1;   // once per host thread and device: the call is on the latency path
  int dev_now = -1;
  LSHX_CUDA(cudaGetDevice(&dev_now));
  if (attr_device != dev_now) {
    LSHX_CUDA(cudaFuncSetAttribute(index_hash_query_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LSHX_CUDA(cudaFuncSetAttribute(index_hash_query_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_device = dev_now;
  }
  const HashQueryArgs a{nq, dim, d_Rp, d_sig, sig_bytes, zero_flag, d_ticket, nb, bpb, keys, ids, main_n, n, cap, out_cap,
                        raw_cap, out_ids, out_coll, out_count, out_count_clamped, out_offs, d_dbg};
  if (nq == 1 && dim <= SMALL_PARAM_FLOATS && X_host != nullptr) {
    static thread_local XRowParam row;       // (the tail beyond dim is never read)
    std::memcpy(row.v, X_host, (size_t)dim * sizeof(float));
    index_hash_query_one_kernel<<<(unsigned)sig_bytes, JN_THREADS, smem, st>>>(row, a);
  } else {
    index_hash_query_small_kernel<<<(unsigned)sig_bytes, JN_THREADS, smem, st>>>(X, a);
  }
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_topk(const int64_t* d_cand, const int64_t* d_raw_off, const int* d_uniq, int64_t nq, int k, int64_t* d_out,
               int* d_out_count, cudaStream_t st) {
  if (nq <= 0 || k <= 0) return LSHX_OK;
  index_topk_kernel<<<grid_for(nq * k, 256), 256, 0, st>>>(d_cand, d_raw_off, d_uniq, nq, k, d_out, d_out_count);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

int index_pos_to_id(const int64_t* d_cand, const int64_t* d_raw_off, const int32_t* d_pos, const int32_t* d_count,
                    int64_t nq, int stride, int64_t* d_out, cudaStream_t st) {
  if (nq <= 0 || stride <= 0) return LSHX_OK;
  index_pos_to_id_kernel<<<grid_for(nq * stride, 256), 256, 0, st>>>(d_cand, d_raw_off, d_pos, d_count, nq, stride, d_out);
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

}  // namespace lshx
