/*
 * lshx.h -- C ABI of liblshx.so: the B200 (sm_100a) hot path of lshrs.
 *
 * The reference (mxngjxa/lshrs, pure Python) has no FFI; its "operator
 * interface" for this path is the Python surface
 *     LSHHasher.__init__ / hash_vector / hash_batch / _project_and_pack
 *         (reference lshrs/hash/lsh.py:51-94, 96-134, 136-169, 171-211)
 *     l2_norm / cosine_similarity / top_k_cosine
 *         (reference lshrs/utils/norm.py:4-61, lshrs/utils/similarity.py:26-90, 93-183)
 * and the call sites LSHRS.ingest / index / query
 *         (reference lshrs/core/main.py:386-411, 442-518, 524-658).
 * Each entry point below names the reference lines it replaces.  A maintainer
 * binds it with ctypes (INTEGRATION.md shows the stub); lshrs_b200/_native.py
 * is that binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative lshx_status otherwise;
 *     lshx_last_error() returns a thread-local message for the last failure.
 *   - all matrices are row-major, contiguous, float32; signatures are uint8 in
 *     exactly the reference's byte order (np.packbits(bitorder="little") per
 *     band, bands concatenated: uint8[n][num_bands][ceil(rows_per_band/8)]).
 *   - the caller owns every buffer it passes; the library owns only what hangs
 *     off the opaque handles and frees it in *_destroy.  Nothing is retained
 *     across calls.
 *   - a handle is bound to one CUDA device.  Calls on one handle are
 *     serialised by an internal mutex (the reference hasher is called from
 *     many Python threads, tests/test_concurrency.py:31-42); use one handle
 *     per GPU for multi-GPU sharding.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the CUDA default
 *     stream, which is also torch's default stream).  Calls taking HOST
 *     buffers are synchronous; calls whose buffers are all DEVICE pointers
 *     only enqueue work on `stream`.
 *   - there is NO CPU fallback: without a usable sm_100 device *_create fails.
 */
#ifndef LSHX_H_
#define LSHX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSHX_ABI_VERSION 3

typedef enum lshx_status {
  LSHX_OK = 0,
  LSHX_ERR_INVALID_ARG = -1, /* bad size / null pointer / unsupported shape        */
  LSHX_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                  */
  LSHX_ERR_NO_DEVICE = -3,   /* no CUDA device, or not compute capability 10.x       */
  LSHX_ERR_OOM = -4,         /* device or pinned-host allocation failed              */
  LSHX_ERR_ZERO_VECTOR = -5  /* rerank: a zero-norm query or candidate (l2_norm)     */
} lshx_status;

/* Which projection kernel lshx_hash_batch launches. */
typedef enum lshx_hash_kernel {
  LSHX_KERNEL_AUTO = 0,    /* tcgen05 when the shape allows it, else FFMA            */
  LSHX_KERNEL_FFMA = 1,    /* FP32 FFMA register-tiled kernel                        */
  LSHX_KERNEL_TCGEN05 = 2, /* tcgen05, TMA-staged, TMEM accumulators, operands split
                              hi + lo; default arithmetic: scaled FP16x3 (power-of-two
                              scale per vector and per projection row, three FP16 MMAs
                              per product, 22 significand bits like 3xTF32; vectors
                              outside the scaled FP16 range are recomputed in FP32)   */
  LSHX_KERNEL_SMALL = 3,   /* reported by lshx_hasher_last_kernel only: the per-vector
                              latency kernel (FP32 FMA, one warp per signature bit) that
                              AUTO uses for a handful of host rows                      */
  LSHX_KERNEL_TCGEN05_3XTF32 = 4,  /* the same kernel, all three terms in TF32 (3xTF32:
                              2x the tensor time of the default)                       */
  LSHX_KERNEL_TCGEN05_TF32BF16 = 5 /* the same kernel, TF32 hi.hi + BF16 cross terms
                              (1.33x the tensor time, fp32-sgemm accuracy)             */
} lshx_hash_kernel;

typedef struct lshx_hasher lshx_hasher;   /* opaque */
typedef struct lshx_reranker lshx_reranker; /* opaque */
typedef struct lshx_index lshx_index;       /* opaque */

/* ---- library ---------------------------------------------------------- */

int lshx_abi_version(void);
/* Thread-local text of the last error raised on this thread ("" if none). */
const char* lshx_last_error(void);
/* Number of visible CUDA devices with compute capability 10.x (0 if none). */
int lshx_device_count(void);
/* Kernels launched by this library in this process since load (all handles). */
uint64_t lshx_launch_count(void);

/* ---- hasher: replaces LSHHasher (reference lshrs/hash/lsh.py) ---------- */

/*
 * Replaces the device-side half of LSHHasher.__init__ (lsh.py:51-94).  The
 * projection matrices are still drawn on the host with numpy's PCG64 stream
 * (lsh.py:93-94) and passed here as ONE row-major float32 matrix
 * R[num_bands*rows_per_band][dim] (band b = rows [b*r, (b+1)*r)).
 * Fails with LSHX_ERR_INVALID_ARG for non-positive sizes (lsh.py:78-83).
 */
int lshx_hasher_create(int device, int dim, int num_bands, int rows_per_band,
                       const float* projections_host, lshx_hasher** out);

/*
 * Re-upload the projections.  Replaces the attribute rebinding
 * `hasher.projections = [...]` done by LSHRS.load_from_disk / __setstate__
 * (reference lshrs/core/main.py:981, 1044).
 */
int lshx_hasher_set_projections(lshx_hasher* h, const float* projections_host);

/* Select the projection kernel (default LSHX_KERNEL_AUTO). */
int lshx_hasher_set_kernel(lshx_hasher* h, int kernel /* lshx_hash_kernel */);
/* The kernel the last lshx_hash_batch on this handle actually launched. */
int lshx_hasher_last_kernel(const lshx_hasher* h);

/* Bytes of signature per vector: num_bands * ceil(rows_per_band / 8). */
int lshx_hasher_signature_bytes(const lshx_hasher* h);

/*
 * Replaces LSHHasher.hash_batch / hash_vector / _project_and_pack
 * (lsh.py:136-169, 96-134, 171-211): for each of the n rows of X,
 *     bit(b, j) = (R[b*r + j] . x) > 0          strict; 0 and NaN give 0
 * packed little-endian per band into out[n][num_bands][ceil(r/8)].
 *
 *   X            n x dim float32, host (x_is_device = 0) or device pointer
 *   out          n x signature_bytes uint8, host or device pointer
 *   zero_flag    optional (may be NULL) n bytes, same memory space as `out`:
 *                1 where every |x_i| <= 1e-8 -- the np.allclose(arr, 0,
 *                atol=1e-8) test of LSHRS._prepare_vector (main.py:1083)
 *                fused into the same pass; NaN elements give 0 like numpy.
 *   stream       cudaStream_t or NULL.
 * n == 0 is a no-op.  Host buffers are staged through handle-owned device
 * buffers in chunks (H2D, kernel, D2H overlapped on two streams); pinned host
 * memory makes those copies asynchronous.
 */
int lshx_hash_batch(lshx_hasher* h, const float* X, int64_t n, int x_is_device,
                    uint8_t* out, int out_is_device, uint8_t* zero_flag, void* stream);

/* Element type of a host batch handed to lshx_hash_batch_typed. */
typedef enum lshx_dtype {
  LSHX_DTYPE_F32 = 0,
  LSHX_DTYPE_F16 = 1, /* IEEE binary16 (numpy float16)                                  */
  LSHX_DTYPE_U8 = 2,  /* e.g. SIFT descriptors                                           */
  LSHX_DTYPE_I8 = 3
} lshx_dtype;

/*
 * lshx_hash_batch for a HOST batch that is not float32.  The reference casts on
 * the host first -- np.asarray(vectors, dtype=np.float32), lsh.py:162 / 243 --
 * which for these types is exact; here the raw rows cross PCIe (half / a quarter
 * of the float32 bytes) and are cast on the device, so the signatures are the
 * same bytes.  X is n x dim elements of `dtype`, row-major, contiguous; out and
 * zero_flag as in lshx_hash_batch with host pointers.  Synchronous.
 */
int lshx_hash_batch_typed(lshx_hasher* h, const void* X, int dtype /* lshx_dtype */, int64_t n,
                          uint8_t* out, uint8_t* zero_flag);

/*
 * Lower-case hex of the band bytes, the variable part of
 * RedisStorage.bucket_key (reference lshrs/storage/redis.py:225:
 * f"{prefix}:{band_id}:bucket:{hash_val.hex()}").  Host helper:
 * sig[n][sig_bytes] -> hex[n][2*sig_bytes] ASCII (no terminator).
 */
int lshx_signatures_to_hex(const uint8_t* sig, int64_t n, int sig_bytes, char* hex_out);

int lshx_hasher_destroy(lshx_hasher* h);

/*
 * Bit mask of the tuning / bring-up environment overrides that are set in this
 * process: 1 LSHX_TC_FLAGS, 2 LSHX_TC_SPLIT (kernel variant / operand split of
 * the tcgen05 kernel), 4 LSHX_COPY_THREADS, 8 LSHX_BOUNCE_MB (pageable-input
 * staging), 16 LSHX_TRACE_PAGEABLE (phase times of that staging on stderr; adds
 * synchronisations).  No reference counterpart; bench.py records it so that a published
 * number states that the product dispatch was not overridden.
 */
int lshx_env_overrides(void);

/*
 * DIAGNOSTICS (tests only; no reference counterpart).  Runs the tcgen05 kernel
 * over X (n rows, HOST float32) and returns the raw fp32 TMEM accumulators of
 * the first tile of pass 0 -- before the sign test that LSHHasher._project_and_pack
 * applies (lsh.py:204) -- row-major acc_out[out_rows][out_cols]:
 * out_rows = min(n, 128) (min(n, 256) when the 2-CTA variant ran, n >= 256),
 * out_cols = accumulator columns of a pass (column c = signature bit c when
 * rows_per_band % 8 == 0).  With the default scaled FP16x3 arm entry (i, c)
 * is s_x[i] * s_r[c] * (x_i . r_c) for exact powers of two s_x, s_r; the
 * tests use it to measure the arithmetic's error on the hardware.
 */
int lshx_hasher_debug_accumulators(lshx_hasher* h, const float* X_host, int64_t n, float* acc_out,
                                   int64_t acc_capacity, int* out_rows, int* out_cols);

/* ---- reranker: replaces top_k_cosine (reference lshrs/utils/similarity.py) */

/* Per-device workspace for rerank calls on vectors of length `dim`. */
int lshx_rerank_create(int device, int dim, lshx_reranker** out);

/*
 * Replaces cosine_similarity + top_k_cosine (similarity.py:80-90, 157-183) and
 * the rank-fraction cut of LSHRS.query (reference lshrs/core/main.py:646-658)
 * for a batch of nq queries.
 *
 *   Q             nq x dim float32 queries
 *   vectors       float32 rows of length dim that candidates are taken from
 *   cand_offsets  nq+1 int64; query i owns candidate slots
 *                 [cand_offsets[i], cand_offsets[i+1])
 *   cand_ids      optional int64 per slot: row of `vectors` holding that
 *                 candidate (gather mode, corpus resident in HBM).  NULL means
 *                 slot s IS row s of `vectors` (packed candidates, what
 *                 top_k_cosine(query, candidates) receives).
 *   n_vectors     rows in `vectors` (ids are range-checked against it; an
 *                 out-of-range id scores NaN, ranks last and is counted in
 *                 out_zero)
 *   max_candidates  largest per-query candidate count; only read when
 *                 on_device == 1 (the offsets are not host-readable then),
 *                 otherwise computed from cand_offsets
 *   k             > 0: keep the k best per query (k >= n_i keeps all n_i)
 *                 <= 0: use p
 *   p             in (0, 1] when k <= 0: keep max(1, ceil(n_i * p)), evaluated
 *                 in double like Python's math.ceil(len * top_p)
 *                 (main.py:650); when k > 0 and p > 0 both apply:
 *                 min(k, max(1, ceil(n_i * p))) (main.py:653-656)
 *   out_stride    slots reserved per query in out_pos / out_score
 *                 (>= the largest per-query result count)
 *   out_pos       nq x out_stride int32: POSITION within the query's candidate
 *                 list (what top_k_cosine returns), best first; ties by
 *                 ascending position
 *   out_score     nq x out_stride float32 cosine similarities, descending
 *   out_count     nq int32: results written for query i
 *   out_zero      nq int32: number of zero-norm vectors met for query i
 *                 (its own query vector counts); l2_norm raises ValueError
 *                 ("Cannot normalize zero vector", norm.py:57) in that case,
 *                 the host wrapper does the same when out_zero[i] != 0.
 *   on_device     0: every pointer is a HOST pointer (synchronous call, staged)
 *                 1: every pointer is a DEVICE pointer (enqueue on `stream`)
 *                 2: Q / offsets / ids / outputs are HOST, `vectors` is DEVICE
 *                    (queries against a corpus resident in HBM)
 * Scores: fp32 dot(c, q) / (||c|| ||q||), within 1e-5 of the reference's
 * normalise-then-dot order (BASELINE.json north_star tolerance).
 * Any number of candidates and results per query (< 2^32): up to 16384 candidates are
 * sorted in shared memory, more are scored in chunks with a running top-8192, and a
 * launch that needs more than 8192 results from more than 16384 candidates sorts the
 * keys in a handle-owned global scratch (slower; the reference's argpartition + argsort
 * has no size limit either, similarity.py:174-179).
 */
int lshx_rerank_topk(lshx_reranker* r, const float* Q, int64_t nq,
                     const float* vectors, int64_t n_vectors,
                     const int64_t* cand_offsets, const int64_t* cand_ids,
                     int64_t max_candidates, int k, double p, int out_stride,
                     int32_t* out_pos, float* out_score, int32_t* out_count,
                     int32_t* out_zero, int on_device, void* stream);

/*
 * Replaces cosine_similarity alone (similarity.py:80-90): scores[s] for every
 * candidate slot, no selection.  Same argument meaning as lshx_rerank_topk;
 * out_scores has total_candidates == cand_offsets[nq] entries.
 */
int lshx_rerank_scores(lshx_reranker* r, const float* Q, int64_t nq,
                       const float* vectors, int64_t n_vectors,
                       const int64_t* cand_offsets, const int64_t* cand_ids,
                       int64_t total_candidates, float* out_scores, int32_t* out_zero,
                       int on_device, void* stream);

/*
 * Replaces l2_norm (reference lshrs/utils/norm.py:48-61) for n rows at once:
 * out[i] = X[i] / ||X[i]||_2 in float32; zero_rows[i] = 1 where the norm is 0
 * (l2_norm raises ValueError("Cannot normalize zero vector") there, norm.py:57;
 * the host wrapper does the same).  on_device: 0 = host pointers (synchronous),
 * 1 = device pointers (enqueued on `stream`).  The rerank kernel does NOT call
 * this -- it fuses the normalisation; the entry point exists for the public
 * helper.
 */
int lshx_l2_normalize(lshx_reranker* r, const float* X, int64_t n, float* out,
                      int32_t* zero_rows, int on_device, void* stream);

int lshx_rerank_destroy(lshx_reranker* r);

/* ---- device band index: batched candidate generation ----
 *
 * Replaces, for BATCHES of queries, the bucket lookups and the collision counting of
 * LSHRS._candidate_counts (reference lshrs/core/main.py:1088-1111: per band one
 * SMEMBERS of the bucket the query's band key names, counts[id] += 1) and the
 * ordering of LSHRS.query (main.py:614: (-collisions, id)).  A mirror of what
 * index() sends to the bucket store -- the (band, key, id) triples -- stays in
 * HBM, sorted by (band, key, id); Redis remains the system of record (every
 * operation still goes to the storage backend unchanged).  Integer work: the
 * candidate lists, their order and the collision counts are exactly the storage
 * path's.  Band keys of at most 8 bytes (rows_per_band <= 64), at most 255 bands,
 * ids in [0, 2^56).
 */
int lshx_index_create(int device, int num_bands, int bytes_per_band, lshx_index** out);
int lshx_index_destroy(lshx_index* ix);
/* Vectors added so far (removed ones included until they are compacted away). */
int64_t lshx_index_size(const lshx_index* ix);

/*
 * Mirror of RedisStorage.batch_add for n vectors (reference lshrs/storage/redis.py
 * batch_add, called from LSHRS.flush, main.py:413-440): signatures as
 * lshx_hash_batch wrote them (n x num_bands x bytes_per_band), ids[n].  on_device:
 * both pointers are device pointers produced on `stream` (signatures need not
 * leave HBM); otherwise host pointers.  SET semantics: adding the same
 * (band, key, id) twice counts once.
 */
int lshx_index_add(lshx_index* ix, const uint8_t* signatures, const int64_t* ids, int64_t n,
                   int on_device, void* stream);
/*
 * Single (band, key, id) operations (RedisStorage.add_to_bucket / a batch_add whose
 * operations are not whole vectors, reference lshrs/storage/redis.py:227-262): n rows
 * of num_bands entries each, ids_per_band[n][num_bands] with -1 = "nothing for this
 * band" (the row's key bytes there are ignored).  Host pointers.
 */
int lshx_index_add_entries(lshx_index* ix, const uint8_t* signatures, const int64_t* ids_per_band, int64_t n);
/* Mirror of RedisStorage.remove_indices (LSHRS.delete, main.py:740-771): host ids. */
int lshx_index_remove(lshx_index* ix, const int64_t* ids_host, int64_t n);
/* Mirror of RedisStorage.clear (LSHRS.clear, main.py:773-791). */
int lshx_index_clear(lshx_index* ix);

/*
 * Candidate lists of nq queries at once: for query i every id that shares at
 * least one band key with signatures[i], ordered by (-collisions, id).  The
 * result stays in handle-owned device buffers until the next query / add / remove /
 * clear on this handle; total_candidates receives the number of candidate SLOTS
 * (an upper bound of the sum of the list lengths: the buffer sizes lshx_index_fetch
 * needs), max_candidates the largest per-query slot count.
 */
int lshx_index_query(lshx_index* ix, const uint8_t* signatures, int64_t nq, int on_device, void* stream,
                     int64_t* total_candidates, int64_t* max_candidates);
/*
 * Latency path of LSHRS.query / get_top_k / get_above_p (reference lshrs/core/main.py:
 * 524-658 hash ONE vector per call and read num_bands buckets): hash nq <= 32 host
 * vectors with `h` (the FP32 small-batch arithmetic of lshx_hash_batch), look their band
 * keys up and join, all in ONE launch and one synchronisation: one CTA per signature
 * byte hashes, reading the vectors from pinned memory in place, and the CTA that
 * finishes last runs the lookup / join and stores the lists into mapped memory.  Query i: its first
 * min(out_count[i], capacity) candidates, ordered by (-collisions, id), in
 * out_ids / out_collisions[i * capacity ..]; out_count[i] = the full list length, or
 * -1 when the query matches more than 4096 bucket entries (take lshx_index_query +
 * lshx_index_fetch then).  zero_flag (optional, nq bytes) as in lshx_hash_batch.
 * capacity <= 4096.  Host pointers; drops the last lshx_index_query result.
 */
int lshx_index_query_vectors(lshx_index* ix, lshx_hasher* h, const float* X, int nq, int capacity,
                             int64_t* out_ids, int32_t* out_collisions, int32_t* out_count,
                             uint8_t* zero_flag);
/*
 * The same latency path with the rerank fused in (LSHRS.query(top_p=...) /
 * get_above_p, main.py:625-658, when the indexed vectors are resident in HBM):
 * hash -> lookup/join -> cosine rerank against corpus_device (candidate id = row)
 * -> ids, three launches and one synchronisation, results stored straight into
 * mapped pinned memory.  Query i keeps min(k, max(1, ceil(n_i * p))) of its n_i
 * candidates (k <= 0: no k; p <= 0: no p), at most out_stride (<= 1024):
 * out_ids / out_score[i * out_stride ..], out_count[i].  out_candidates[i] = n_i, or
 * -1 when the query matches more than 1024 bucket entries (nothing is ranked then:
 * take lshx_index_query + lshx_index_rerank).  out_zero (optional): zero-norm or
 * out-of-range candidate vectors met; zero_flag (optional) as in lshx_hash_batch.
 */
int lshx_index_query_rerank_vectors(lshx_index* ix, lshx_hasher* h, lshx_reranker* r, const float* X, int nq,
                                    const float* corpus_device, int64_t n_vectors, int k, double p,
                                    int out_stride, int64_t* out_ids, float* out_score, int32_t* out_count,
                                    int32_t* out_zero, int32_t* out_candidates, uint8_t* zero_flag);
/*
 * DIAGNOSTICS (no reference counterpart, not on any product path): with enable != 0 the
 * fused latency kernel of lshx_index_query_vectors records %globaltimer stamps (ns) of its
 * phases for query 0 -- last CTA entered, hashed, ticket taken, bands searched, candidates
 * gathered, first sort, counted + second sort, list stored -- and the number of bucket
 * entries matched; stamps_out (9 words, may be NULL) receives the last launch's.
 */
int lshx_index_debug_timeline(lshx_index* ix, int enable, uint64_t* stamps_out);
/*
 * RedisStorage.get_bucket (SMEMBERS, reference lshrs/storage/redis.py:264-301) for m
 * buckets at once -- what makes the index usable as the bucket STORE, not only as
 * a mirror: bucket t = (band_ids[t], keys[t * bytes_per_band ..]); its members (live
 * ids, each once, ascending) go to ids_out[offsets[t] .. offsets[t + 1]).  With
 * ids_out == NULL only sizes are computed: offsets then bound the member counts from
 * above (removed ids and repeats still counted) and *needed is the ids_out capacity
 * a second call needs.  Host pointers; drops the last lshx_index_query result.
 */
int lshx_index_get_buckets(lshx_index* ix, const int32_t* band_ids, const uint8_t* keys, int64_t m,
                           int64_t* offsets, int64_t* ids_out, int64_t ids_capacity, int64_t* needed);
/*
 * Everything the index holds, for persistence (the reference leaves bucket data to
 * Redis' own RDB/AOF; an HBM store needs a way out): *n_out = entries per band;
 * keys_out[num_bands][n][bytes_per_band], ids_out[num_bands][n] (-1 = removed), both
 * host buffers of `capacity` entries per band >= n, or both NULL for the size only.
 * Re-import with lshx_index_add_entries (one row per entry).
 */
int lshx_index_export(lshx_index* ix, uint8_t* keys_out, int64_t* ids_out, int64_t capacity, int64_t* n_out);
/*
 * lshx_index_query for nq HOST vectors in one pass over PCIe (LSHRS.query_batch): the
 * vectors are uploaded once, hashed on the device with `h` (same kernel choice as
 * lshx_hash_batch makes for a device batch), joined -- the signatures never visit
 * the host -- and stay in the handle, so that lshx_index_rerank can be called with
 * Q = NULL for this result.  zero_flag (optional, nq bytes, host) as in
 * lshx_hash_batch.
 */
int lshx_index_query_host_vectors(lshx_index* ix, lshx_hasher* h, const float* X, int64_t nq,
                                  uint8_t* zero_flag, int64_t* total_candidates, int64_t* max_candidates);
/*
 * Host copies of the last query's lists: query i owns ids[offsets[i] ..
 * offsets[i] + counts[i]) and, when collisions != NULL, the matching collision
 * counts.  offsets has nq + 1 entries, ids / collisions total_candidates.  Any
 * pointer may be NULL.
 */
int lshx_index_fetch(lshx_index* ix, int64_t* offsets, int32_t* counts, int64_t* ids, int32_t* collisions);
/*
 * get_top_k for the last query (main.py:616-623): the first min(top_k, counts[i])
 * candidates of every list into out_ids[nq][top_k] (-1 padded), out_count[nq]. Host.
 */
int lshx_index_topk(lshx_index* ix, int top_k, int64_t* out_ids, int32_t* out_count);
/*
 * get_above_p / query(top_p=...) for the last query (main.py:625-658): rerank every
 * list by cosine against `corpus_device` (float32 [n_vectors][dim] resident in HBM,
 * candidate id = row: the device-side stand-in for vector_fetch_fn) with the rerank
 * kernel of `r`, keep min(k, max(1, ceil(n_i * p))) (k <= 0: no k; p <= 0: no p)
 * and return IDS (not positions): out_ids[nq][out_stride] (-1 padded), out_score,
 * out_count, out_zero (zero-norm vectors met, as lshx_rerank_topk).  Q: the nq
 * query vectors, host (q_on_device = 0) or device, or NULL after
 * lshx_index_query_host_vectors (the handle still holds them).  Outputs are host pointers.
 */
int lshx_index_rerank(lshx_index* ix, lshx_reranker* r, const float* Q, int q_on_device,
                      const float* corpus_device, int64_t n_vectors, int k, double p, int out_stride,
                      int64_t* out_ids, float* out_score, int32_t* out_count, int32_t* out_zero);

/* ---- multi-process plumbing for the signature relay (no reference counterpart) ----
 *
 * One rank per GPU (torchrun) cannot share device memory with, or order its streams
 * against, another process without CUDA IPC.  lshrs_b200/fabric.py uses these to let
 * a rank on a slow host link hand its signatures over NVLink to a partner on a fast
 * one.  Plumbing only -- no arithmetic.  IPC handles are the
 * 64 opaque bytes of cudaIpcMemHandle_t / cudaIpcEventHandle_t; `device` is the
 * caller's GPU; `stream` a cudaStream_t of that GPU.
 */
int lshx_ipc_mem_alloc(int device, size_t bytes, void** dptr, unsigned char* handle_out /* 64 bytes */);
/* Map another process's allocation; peer access from `device` is enabled lazily. */
int lshx_ipc_mem_open(int device, const unsigned char* handle, void** dptr);
int lshx_ipc_mem_close(void* dptr);
int lshx_ipc_mem_free(int device, void* dptr);
int lshx_ipc_event_create(int device, void** event, unsigned char* handle_out /* 64 bytes */);
int lshx_ipc_event_open(int device, const unsigned char* handle, void** event);
int lshx_ipc_event_record(int device, void* event, void* stream);
/* cudaStreamWaitEvent: waits for the most recent record ISSUED before this call. */
int lshx_ipc_event_wait(int device, void* event, void* stream);
int lshx_ipc_event_destroy(void* event);
/* cudaMemcpyAsync(cudaMemcpyDefault): peer-to-peer (NVLink), D2H or H2D on `stream`. */
int lshx_memcpy_async(int device, void* dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSHX_H_ */
