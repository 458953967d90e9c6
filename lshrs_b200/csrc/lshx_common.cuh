// Shared declarations of liblshx.so (internal; the public ABI is include/lshx.h).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>

#include "lshx.h"

namespace lshx {

// Thread-local error text returned by lshx_last_error().
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define LSHX_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ::lshx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                        \
      return (e_ == cudaErrorMemoryAllocation) ? LSHX_ERR_OOM : LSHX_ERR_CUDA;            \
    }                                                                                     \
  } while (0)

#define LSHX_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      ::lshx::set_error(__VA_ARGS__); \
      return LSHX_ERR_INVALID_ARG;    \
    }                                 \
  } while (0)

// Device-side description of one hasher (what the kernels need).
//
// Column layout: the kernels compute one sign bit per COLUMN of a padded
// projection matrix Rp[ncols_pad][dim].  Band b owns columns
// [b*8*bpb, b*8*bpb + rows_per_band); the remaining columns of each band (when
// rows_per_band is not a multiple of 8) and the tail padding are ZERO rows, so
// their projection is exactly 0, `> 0` is false, and the packed byte stream of
// the columns is bit-for-bit np.packbits(bitorder="little") per band
// (reference lshrs/hash/lsh.py:204-211).
struct HashShape {
  int dim;
  int num_bands;
  int rows_per_band;
  int bpb;        // bytes per band = ceil(rows_per_band / 8)
  int sig_bytes;  // num_bands * bpb
  int ncols;      // sig_bytes * 8
  int ncols_pad;  // ncols rounded up to a multiple of 128
  int dim_pad;    // dim rounded up to a multiple of 32 (K padding of the split copies)
};

// ---- hash kernels (hash_ffma.cu / hash_tc.cu) --------------------------------
// Rp: [ncols_pad][dim] fp32 with zero padding rows.
int launch_hash_ffma(const HashShape& s, const float* d_X, int64_t n, const float* d_Rp,
                     uint8_t* d_out, uint8_t* d_zero_flag, cudaStream_t stream);

// Recompute mode of the same kernel: the 128-row tiles listed in d_tile_list[0 .. d_tile_count[0]).  The
// kernel zeroes d_tile_count[0] (and its exit ticket d_tile_count[1]) when its last CTA leaves.
int launch_hash_ffma_tiles(const HashShape& s, const float* d_X, int64_t n, const float* d_Rp, uint8_t* d_out,
                           const int* d_tile_list, int* d_tile_count, int num_ctas, cudaStream_t stream);

// Small batches (<= hash_small_max_rows): one CTA per output byte, one warp per column.  `out` /
// `zero_flag` may be mapped pinned host memory (the kernel stores straight into it).
int hash_small_max_rows(const HashShape& s);
int launch_hash_small(const HashShape& s, const float* d_X, int n, const float* d_Rp, uint8_t* out,
                      uint8_t* zero_flag, cudaStream_t stream);
// ONE host row (dim <= 1024) carried in the kernel's parameter block: no copy in front of the launch
int launch_hash_small_one(const HashShape& s, const float* x_host, const float* d_Rp, uint8_t* out, uint8_t* zero_flag,
                          cudaStream_t stream);

struct TcPlan;  // opaque tcgen05 state (TMA descriptor of the split projections, ...)
bool tc_shape_supported(const HashShape& s);
int tc_plan_create(const HashShape& s, const float* d_Rp, TcPlan** out);
void tc_plan_destroy(TcPlan* p);
bool tc_plan_f16_ok(const TcPlan* p);       // false: a projection row has no representable FP16 scale
int tc_plan_default_split(const TcPlan* p); // what split < 0 resolves to (2 unless LSHX_TC_SPLIT overrides)
int tc_plan_flags(const TcPlan* p);         // TC_FLAG_* in effect (LSHX_TC_FLAGS overrides the default)
// diagnostics (tests): keep the first tile's raw fp32 accumulators of every following launch
int tc_plan_set_debug(TcPlan* p, bool on);
const float* tc_plan_debug_buffer(const TcPlan* p, int* cols);
// split: 0 = 3xTF32, 1 = TF32 hi.hi + BF16 cross terms, 2 = scaled FP16x3, < 0 = the plan's choice
int launch_hash_tc(const HashShape& s, TcPlan* plan, int split, const float* d_X, int64_t n, uint8_t* d_out,
                   uint8_t* d_zero_flag, cudaStream_t stream);

// ---- rerank kernel (rerank.cu) ------------------------------------------------
struct RerankArgs {
  const float* Q;
  int64_t nq;
  const float* V;
  int64_t n_vectors;
  const int64_t* offs;
  const int64_t* ids;  // may be null
  const int32_t* cand_counts;  // may be null: query i has cand_counts[i] candidates starting at offs[i]
                               // (lists that do not fill their slot range: the device join's output)
  int dim;
  int k;
  double p;
  int out_stride;
  int32_t* out_pos;
  float* out_score;
  int32_t* out_count;
  int32_t* out_zero;
  float* all_scores;  // optional: every candidate's score (cosine_similarity)
  int64_t max_cand;   // largest per-query candidate count (sizes the sort buffer)
  bool select;        // false: scores only
  // oversized selections (see rerank_scratch_bytes): global sort scratch owned by the caller
  uint64_t* big_keys;   // null unless the launch needs the global-memory sort
  size_t big_bytes;
  int64_t big_stride;   // set by launch_rerank
};
// Bytes of device scratch launch_rerank needs for `a` (0 for all but oversized selections: a query with more
// than 16384 candidates AND more than 8192 results); queries_per_pass = queries sorted per scratch fill.
size_t rerank_scratch_bytes(const RerankArgs& a, int64_t* queries_per_pass);
int launch_rerank(const RerankArgs& a, cudaStream_t stream);
int launch_l2_normalize(const float* d_X, int64_t n, int dim, float* d_out, int32_t* d_zero,
                        cudaStream_t stream);

// ---- device band index / candidate join (index_join.cu) ---------------------------
// Segments: keys[b * cap + e], ids[b * cap + e] for band b, entry e < n.
int index_append(const uint8_t* d_sig, const int64_t* d_ids, int64_t n, int nb, int bpb, uint64_t* keys, int64_t* ids,
                 int64_t cap, int64_t at, unsigned long long* d_max_id, int* d_bad, int per_band, cudaStream_t st);
int index_bucket_lookup(const int* d_band_ids, const uint64_t* d_want, int64_t m, const uint64_t* keys, int64_t n,
                        int64_t cap, int64_t* d_lo, int64_t* d_cnt, cudaStream_t st);
int index_bucket_gather(const int* d_band_ids, const int64_t* d_lo, const int64_t* d_off, int64_t m, const int64_t* ids,
                        int64_t cap, int64_t* d_out, cudaStream_t st);
size_t index_sort_hist_entries(int64_t n, int nb);
int index_sort(uint64_t* keys[2], int64_t* ids[2], int* cur, int64_t first, int64_t n, int64_t cap, int nb,
               int key_bytes, int id_bytes, unsigned* d_hist, size_t hist_entries, cudaStream_t st);
int index_tombstone(int64_t* ids, int64_t n, int64_t cap, int nb, const int64_t* d_gone_sorted, int64_t ngone,
                    cudaStream_t st);
int index_lookup_scan(const uint8_t* d_sig, int64_t nq, int nb, int bpb, const uint64_t* keys, int64_t main_n, int64_t n,
                      int64_t cap, int64_t* d_lo, int* d_cnt, int* d_raw_count, int64_t* d_raw_off, int64_t* d_ws_off,
                      int64_t* d_meta, cudaStream_t st);
unsigned index_join_smem_cap();
int index_query_small(const uint8_t* d_sig, int nq, int nb, int bpb, const uint64_t* keys, const int64_t* ids,
                      int64_t main_n, int64_t n, int64_t cap, int out_cap, int raw_cap, int64_t* out_ids, int* out_coll, int* out_count,
                      int* out_count_clamped, int64_t* out_offs, cudaStream_t st);
int index_join(int64_t nq, int nb, int nruns, const int64_t* ids, int64_t cap, const int64_t* d_lo, const int* d_cnt,
               const int* d_raw_count, const int64_t* d_raw_off, const int64_t* d_ws_off, uint64_t* d_ws,
               int64_t ws_total, int64_t* d_out_ids, int* d_out_coll, int* d_uniq, cudaStream_t st);
// hash (small-batch FP32 kernel body) + lookup / join / emit of nq <= hash_small_max_rows vectors in ONE launch
// (X: device-readable, e.g. mapped pinned memory; X_host: the same rows readable by the host, or null -- ONE row of
// up to 1024 floats then travels in the kernel's parameter block)
int index_hash_query_small(const float* X, const float* X_host, int nq, int dim, const float* d_Rp, uint8_t* d_sig,
                           int sig_bytes, uint8_t* zero_flag, unsigned* d_ticket, int nb, int bpb, const uint64_t* keys,
                           const int64_t* ids, int64_t main_n, int64_t n, int64_t cap, int out_cap, int raw_cap,
                           int64_t* out_ids, int* out_coll, int* out_count, int* out_count_clamped, int64_t* out_offs,
                           unsigned long long* d_dbg, cudaStream_t st);
int index_topk(const int64_t* d_cand, const int64_t* d_raw_off, const int* d_uniq, int64_t nq, int k, int64_t* d_out,
               int* d_out_count, cudaStream_t st);
int index_pos_to_id(const int64_t* d_cand, const int64_t* d_raw_off, const int32_t* d_pos, const int32_t* d_count,
                    int64_t nq, int stride, int64_t* d_out, cudaStream_t st);

}  // namespace lshx
