// tcgen05 projection-and-sign kernel (sm_100a): split-operand MMAs, TMA-staged, accumulators in TMEM.
//
// Replaces LSHHasher._project_and_pack (reference lshrs/hash/lsh.py:200-211) for a whole
// batch, like hash_ffma.cu, but on the 5th-generation tensor cores.  Both operands are split
//
//     x = x_hi + x_lo, r = r_hi + r_lo,      x.r ~= x_lo.r_hi + x_hi.r_lo + x_hi.r_hi   (fp32 accumulate)
//
// which keeps ~22 significand bits per operand (the dropped x_lo.r_lo term is <= 2^-22 relative), i.e.
// at least the accuracy of an fp32 sgemm with a different summation order -- far inside the 1e-5
// relative margin the parity contract allows, where single-pass TF32 is not.
// Three arithmetic arms (template parameter kSplit):
//     2  scaled FP16x3 (default): power-of-two scales per vector / projection row, three kind::f16 MMAs
//     1  TF32 hi.hi + BF16 cross terms in one K-doubled kind::f16 MMA
//     0  3xTF32
//
// Data flow per CTA (persistent, one CTA per SM, 384 threads; hash_tc2_kernel pairs two CTAs with
// cta_group::2 so that each stages only half of the projection columns):
//
//   warp 0      X TMA producer: tile chunks of 128 rows x 32 floats (SWIZZLE_128B)
//   warp 3      projection TMA producer: per chunk, two 16-k halves of the pre-split projection planes for
//               every column of the pass (rows of 64 B, SWIZZLE_64B)
//   warps 4-7   converters: thread t owns row t of the tile; reads its 128 B of the X chunk from
//               shared memory, splits hi/lo, writes them to TMEM with tcgen05.st (A operand lives
//               in TMEM, so the MMA never re-reads X from shared memory); fuses the zero-vector
//               test of LSHRS._prepare_vector (reference lshrs/core/main.py:1083)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma, M=128 (256 over a CTA pair),
//               N=256 (or 128 / the compact column count), A from TMEM, B (projection chunk) from shared memory
//   warps 8-11  epilogue: tcgen05.ld the fp32 accumulators, strict `> 0`, one bit per column into
//               per-row words, 16 B of signature per 128 columns, coalesced store
//   warp 2      TMEM allocation
//
// TMEM (512 columns): accumulators in columns [0,256) (one 256-column tile, or two stages of a
// 128-column tile), A-operand stages in [256,512): 4 stages x 64 columns.
// Shared memory: 192 KB of X stages (16 KB each) and projection stages.
//
// Measured on B200 (profiles/): 917-991 M vectors/s at dim 768 / 256 bits with the default arm, tensor
// pipe 87 % active under the board's power cap, DRAM traffic equal to the algorithmic bytes.

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lshx_common.cuh"

namespace lshx {

namespace {

constexpr int TC_THREADS = 384;
constexpr int TC_THREADS_2CONV = 512;   // + a second converter warpgroup (warps 12-15), 1-CTA kernel only
constexpr int TM = 128;          // rows per tile
constexpr int TK = 32;           // floats per K chunk (128 B: one SWIZZLE_128B row)
constexpr int TN = 128;          // accumulator columns per column tile (16 signature bytes)
constexpr int TKB = 16;          // floats per projection stage along K (64 B rows, SWIZZLE_64B)
constexpr int XS_MAX = 10;       // X stages: as many 16 KB stages as fit beside the projection stages
constexpr int BS = 4;            // projection stages
constexpr int AS = 4;            // A-operand TMEM stages
constexpr uint32_t X_STAGE_BYTES = TM * TK * 4;        // 16384
constexpr uint32_t SMEM_BYTES = 196608;                // 4 x 16 KB of X + 4 x 32 KB of projections at N = 256
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t A_COL0 = 256;
constexpr uint32_t A_STAGE_COLS = 64;
// dispatch switches (TcParams::flags).  The default is B_WARP | CG2; LSHX_TC_FLAGS in the environment
// overrides it when a plan is created -- the tests use that to drive every variant on small batches.
constexpr int TC_FLAG_B_WARP = 1;       // projections are TMA-loaded by warp 3 instead of warp 0
constexpr int TC_FLAG_CG2 = 8;          // 2-CTA kernel (cta_group::2) where the shape allows it
constexpr int TC_FLAG_CG2_ALWAYS = 32;  // ... even for batches smaller than one 256-row tile per SM pair
constexpr int TC_FLAG_CONV2 = 64;       // 1-CTA kernel: two converter warpgroups even when the heuristic says one
constexpr int TC_FLAG_CONV1 = 128;      // 1-CTA kernel: never two

// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N=n (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// the same for A=B=BF16 (K = 16 per instruction): the cross terms of the TF32+BF16 split
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// and for A=B=FP16 (format 0): the scaled FP16x3 split
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// ---- PTX wrappers -----------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// One lane of a fully converged warp (elect.sync): lets ptxas keep the operands of the uniform-datapath
// instructions (UTMALDG, UTCHMMA, UTCBAR) in uniform registers instead of a per-lane waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]; kind::tf32, M=128, N from idesc, K=8
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with kind::f16 (operand types from the instruction descriptor: BF16), K=16
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two fp32 -> one word of two BF16 (round to nearest even), `even` in the low half
__device__ __forceinline__ uint32_t pack_bf16(float even, float odd) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(odd), "f"(even));
  return d;
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 64 B
// (16 floats = two K=8 steps), 8-row swizzle atoms 512 B apart (SBO = 32 x 16 B), LBO unused for
// swizzled K-major, version 1 (Blackwell), layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) |
         (4ull << 61);
}

// round-to-nearest (ties away) to TF32, as cvt.rna.tf32.f32 does, with integer ops
__device__ __forceinline__ uint32_t tf32_rna(uint32_t u) { return (u + 0x1000u) & 0xFFFFE000u; }

struct Ring {  // stage index + phase bit of one mbarrier ring
  uint32_t idx = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++idx == n) {
      idx = 0;
      phase ^= 1;
    }
  }
};

struct TcParams {
  int64_t n;        // rows
  int kc;           // K chunks = dim_pad / 32
  int ncols_pass;   // accumulator columns per pass = N of the MMA (multiple of 16, <= 256)
  int xs;           // X stages in use (4 at N = 256, up to XS_MAX for narrow shapes: more bytes in flight)
  int b_resident;   // 1: the whole split projection matrix of the (single) pass stays in shared memory
                    //    for the life of the CTA (small shapes); 0: streamed through BS stages per tile
  int npass;        // passes over X
  int repack;       // 1: columns are compact (band b = columns [b*r, (b+1)*r) of its pass) and the
                    //    epilogue expands every band to whole bytes (rows_per_band % 8 != 0)
  int r, bpb, bpp, num_bands;  // rows per band, bytes per band, bands per pass (repack mode)
  int64_t mtiles;   // ceil(n / 128)
  int sig_bytes;
  int out_vec_ok;   // 16-byte stores allowed
  int flags;        // TC_FLAG_*
  uint8_t* out;
  uint8_t* zero_flag;
  // FP16x3: 128-row tiles holding a vector that does not fit the scaled FP16 range are appended here
  // (one entry per converter warp that saw one) and recomputed in FP32 right after (launch_hash_tc)
  int* redo_count;
  int* redo_list;
  // diagnostics (lshx_hasher_debug_accumulators, tests only; nullptr in every product launch): the raw fp32
  // TMEM accumulators of the first tile, pass 0, row-major [tile rows][ncols_pass]
  float* dbg_acc;
};

// One half (16 floats) of thread t's 128 B row of an X chunk in shared memory -> A-operand words.
// hi: TF32(x).  lo: 3xTF32 -> TF32 of the residual; TF32+BF16 -> words [0,8) = BF16 pairs of the residual,
// words [8,16) = BF16 pairs of hi (the layout of the cross plane, split_cross_kernel).  `viol` collects
// the zero-vector test of LSHRS._prepare_vector (some |x| > 1e-8, or NaN).
template <int kSplit>
__device__ __forceinline__ void convert_half(uint32_t row, int t, int h, uint32_t (&hi)[16],
                                             uint32_t (&lo)[16], bool& viol) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int chunk = h * 4 + c;                       // logical 16 B chunk of the row
    const uint32_t addr = row + (uint32_t)((chunk ^ (t & 7)) << 4);  // SWIZZLE_128B
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    const float f[4] = {v.x, v.y, v.z, v.w};
    float lf[4], hc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = f[e];
      viol |= !(fabsf(x) <= 1e-8f);
      const uint32_t hu = tf32_rna(__float_as_uint(x));
      const float hf = __uint_as_float(hu);
      // residual is exact in fp32; an infinite hi has no residual (inf - inf would be NaN)
      lf[e] = (fabsf(hf) == INFINITY) ? 0.f : (x - hf);
      hi[c * 4 + e] = hu;
      if (kSplit == 0) lo[c * 4 + e] = tf32_rna(__float_as_uint(lf[e]));
      // a hi that would round to a BF16 infinity (or is one) takes no part in the cross term:
      // its sign is already decided by hi.hi
      else hc[e] = ((hu & 0x7FFFFFFFu) >= 0x7F7F8000u) ? 0.f : hf;
    }
    if (kSplit != 0) {
      lo[c * 2 + 0] = pack_bf16(lf[0], lf[1]);
      lo[c * 2 + 1] = pack_bf16(lf[2], lf[3]);
      lo[8 + c * 2 + 0] = pack_bf16(hc[0], hc[1]);
      lo[8 + c * 2 + 1] = pack_bf16(hc[2], hc[3]);
    }
  }
}

// two fp32 -> one word of two FP16 (round to nearest even), `even` in the low half
__device__ __forceinline__ uint32_t pack_f16(float even, float odd) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(odd), "f"(even));
  return d;
}
__device__ __forceinline__ void unpack_f16(uint32_t w, float& even, float& odd) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
      : "=f"(even), "=f"(odd)
      : "r"(w));
}

// ---- scaled FP16x3 split (kSplit == 2) ------------------------------------------------------------
// Only the SIGN of x.r is kept, and it is invariant under a positive scale per vector and per projection
// row.  With power-of-two scales (exact) that bring the operands into FP16's range,
//     y = s_x * x = y_hi + y_lo,  q = s_r * r = q_hi + q_lo     (hi = FP16(.), lo = FP16 of the residual)
//     y.q ~= y_lo.q_hi + y_hi.q_lo + y_hi.q_hi                  (three kind::f16 MMAs, K = 16 each)
// keeps 11 + 11 significand bits per operand like 3xTF32 (measured max error 1.5e-8 |x||r|) at HALF the
// tensor time per term: 1.5 tensor-time units per product, and half the projection bytes.
//   s_r: per projection row, from its largest |r| (static, split_f16_kernel).
//   s_x: per vector, fixed by the first K chunk that holds a non-zero element so that its largest |x|
//        lands in [2, 4): later elements may be up to 2^14 times larger before FP16 overflows, and
//        elements far below it only meet FP16's absolute floor (2^-25 per element, i.e. at most
//        sqrt(dim) * 2^-26 of |x||r| in total by Cauchy-Schwarz: 4e-7 at dim 768).  A vector that does
//        overflow (or whose scale is not representable: fp32 denormals, inf) has its 128-row tile
//        appended to TcParams::redo_list and recomputed by the FP32 FFMA kernel right after this one
//        (launch_hash_tc) -- never silently wrong.

// largest |x| of thread t's 32 floats of an X chunk
__device__ __forceinline__ float chunk_absmax(uint32_t row, int t) {
  float m = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t addr = row + (uint32_t)((c ^ (t & 7)) << 4);
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));   // fmaxf drops NaN
  }
  return m;
}
// power of two s with m * s in [2, 4); 0 when it is not representable (m denormal or inf)
__device__ __forceinline__ float row_scale_for(float m) {
  const int e = (int)((__float_as_uint(m) >> 23) & 0xFFu);
  const int se = 254 + 1 - e;                       // biased exponent of the scale
  return (e >= 1 && e < 255) ? __uint_as_float((uint32_t)se << 23) : 0.f;
}
// NaN-propagating max(m, |x|): one instruction per element carries both per-row tests of the FP16x3
// converters -- "some |x| > 1e-8 or NaN" (the zero-vector test) and "some |s_x * x| > 65504" (FP16 overflow).
__device__ __forceinline__ float absmax_nan(float m, float x) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(m), "f"(fabsf(x)));
  return d;
}
// One half (16 floats) of a row -> 16 A-operand words: [0,8) = FP16 pairs of y_hi, [8,16) = pairs of y_lo.
// `mx` accumulates max |x| over the row (NaN sticks).
__device__ __forceinline__ void convert_half_f16(uint32_t row, int t, int h, float sc, uint32_t (&w)[16],
                                                 float& mx) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int chunk = h * 4 + c;
    const uint32_t addr = row + (uint32_t)((chunk ^ (t & 7)) << 4);  // SWIZZLE_128B
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    const float f[4] = {v.x, v.y, v.z, v.w};
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mx = absmax_nan(mx, f[e]);
      y[e] = f[e] * sc;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t hw = pack_f16(y[2 * q], y[2 * q + 1]);
      float h0, h1;
      unpack_f16(hw, h0, h1);
      // residuals are exact in fp32; rows with an infinite hi are recomputed anyway
      w[c * 2 + q] = hw;
      w[8 + c * 2 + q] = pack_f16(y[2 * q] - h0, y[2 * q + 1] - h1);
    }
  }
}
// the two row tests from the running max: viol = some |x| > 1e-8 (or NaN); overflow = the scaled row
// leaves FP16's range (s_x is fixed before the first non-zero element is converted, so max|x| * s_x is
// the largest scaled magnitude; NaN compares false: a NaN row hashes to zeros like numpy)
__device__ __forceinline__ bool row_viol(float mx) { return !(mx <= 1e-8f); }
__device__ __forceinline__ bool row_overflow(float mx, float sc) { return mx * sc > 65504.f; }

// Accumulator row of this thread (TMEM lane) -> one sign bit per column: words[c >> 5] bit (c & 31) = D[c] > 0.
__device__ __forceinline__ void read_sign_words(uint32_t tmem_row, uint32_t N, uint32_t (&words)[8]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {   // 64 columns per step: two 32-column loads in flight per wait
    uint32_t w0 = 0, w1 = 0;
    if ((uint32_t)(g * 64) < N) {
      uint32_t v0[32], v1[32];
      const uint32_t src = tmem_row + g * 64;
      tc_ld32(src, v0);
      tc_ld32(src + 32, v1);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        w0 |= (__uint_as_float(v0[i]) > 0.f ? 1u : 0u) << i;
        w1 |= (__uint_as_float(v1[i]) > 0.f ? 1u : 0u) << i;
      }
      // columns at or past N were never written by this pass's MMAs
      const uint32_t left = N - (uint32_t)(g * 64);
      if (left < 32u) w0 &= (1u << left) - 1u;
      if (left <= 32u) w1 = 0u;
      else if (left < 64u) w1 &= (1u << (left - 32u)) - 1u;
    }
    words[2 * g] = w0;
    words[2 * g + 1] = w1;
  }
}

// Diagnostics: this thread's accumulator row as it sits in TMEM (before the sign test) -> dst[0 .. N).
__device__ __noinline__ void dump_accumulator_row(uint32_t tmem_row, uint32_t N, float* dst) {
  for (uint32_t c = 0; c < N; c += 32) {
    uint32_t v[32];
    tc_ld32(tmem_row + c, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c + (uint32_t)i < N) dst[c + i] = __uint_as_float(v[i]);
  }
}

// Sign words of row m (thread t of the tile), pass `pass` -> signature bytes in the reference's layout.
__device__ __forceinline__ void store_signature(const TcParams& p, uint32_t N, int pass, int64_t m, int t,
                                                const uint32_t (&words)[8], uint32_t* repack_sm) {
  if (p.repack && p.r == 4) {
    // 4-row bands (BASELINE config 5): every nibble of the compact bits becomes one byte; pure
    // register bit-spreading, 16 bits -> 4 bytes per step
    if (m < p.n) {
      const int band0 = pass * p.bpp;
      const int nb = (p.num_bands - band0 < p.bpp) ? (p.num_bands - band0) : p.bpp;  // = output bytes
      uint8_t* dst = p.out + m * (int64_t)p.sig_bytes + band0;
      const bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {          // 64 compact bits -> 16 output bytes per step
        if (g * 16 < nb) {
          uint32_t o[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t x = (words[2 * g + (h >> 1)] >> (16 * (h & 1))) & 0xFFFFu;
            x = (x | (x << 8)) & 0x00FF00FFu;
            x = (x | (x << 4)) & 0x0F0F0F0Fu;
            o[h] = x;
          }
          if (vec_ok && g * 16 + 16 <= nb) {
            *reinterpret_cast<uint4*>(dst + g * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (g * 16 + e < nb) dst[g * 16 + e] = (uint8_t)(o[e >> 2] >> (8 * (e & 3)));
          }
        }
      }
    }
  } else if (p.repack) {
    // compact column bits -> every band padded to whole bytes (np.packbits zero high bits)
    uint32_t* mine = repack_sm + t * 9;
#pragma unroll
    for (int i = 0; i < 8; ++i) mine[i] = words[i];
    mine[8] = 0u;
    if (m < p.n) {
      const int band0 = pass * p.bpp;
      const int nb = (p.num_bands - band0 < p.bpp) ? (p.num_bands - band0) : p.bpp;
      const int out_bytes = nb * p.bpb;                       // bytes this pass contributes to the row
      uint8_t* dst = p.out + m * (int64_t)p.sig_bytes + (int64_t)band0 * p.bpb;
      const bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
      int j = 0, q = 0;                                       // band within the pass, byte within the band
      for (int ob = 0; ob < out_bytes; ob += 16) {            // 16 output bytes per step, in registers
        uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (ob + e < out_bytes) {
            const int src = j * p.r + 8 * q;
            const int nbits = (p.r - 8 * q < 8) ? (p.r - 8 * q) : 8;
            const uint32_t v = __funnelshift_r(mine[src >> 5], mine[(src >> 5) + 1], src & 31) &
                               ((1u << nbits) - 1u);
            o[e >> 2] |= v << (8 * (e & 3));
            if (++q == p.bpb) {
              q = 0;
              ++j;
            }
          }
        }
        if (vec_ok && ob + 16 <= out_bytes) {
          *reinterpret_cast<uint4*>(dst + ob) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (ob + e < out_bytes) dst[ob + e] = (uint8_t)(o[e >> 2] >> (8 * (e & 3)));
        }
      }
    }
  } else if (m < p.n) {
    const int byte0 = pass * (int)(N / 8);  // 16 signature bytes per 128 columns
    uint8_t* dst = p.out + m * (int64_t)p.sig_bytes + byte0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if ((uint32_t)(j * TN) < N) {
        const int b = byte0 + j * 16;
        if (p.out_vec_ok && b + 16 <= p.sig_bytes) {
          *reinterpret_cast<uint4*>(dst + j * 16) =
              make_uint4(words[j * 4], words[j * 4 + 1], words[j * 4 + 2], words[j * 4 + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (b + q < p.sig_bytes) dst[j * 16 + q] = (uint8_t)(words[j * 4 + (q >> 2)] >> (8 * (q & 3)));
        }
      }
    }
  }
}

constexpr int BS2 = 6;            // projection stages of (N/2) x 64 B x 2 planes (16 KB at N = 256)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive(cta_id) does: what the waiter must see are TMEM writes, ordered by
// tcgen05.wait::st + tcgen05.fence::before_thread_sync, not generic-proxy stores.  (A
// .release.cluster arrive costs MEMBAR.ALL + ERRBAR per call -- half of the converters' time when it
// was tried -- and .acquire.cluster waits invalidate L1 on every pass.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are counted on a barrier of the
// leader CTA (`bar_cluster` is a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far are done
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma2_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MMAs of one K half (16 k) of a chunk.  kCG = cta_group (1: M = 128, 2: M = 256 over the CTA pair).
// a_stage: first TMEM column of the A stage; bs: shared-memory address of the projection stage;
// plane_bytes: size of one plane of it.  +32 B per K step inside a 64 B swizzle row = start-address
// field += 2.  Small terms first.
template <int kSplit, int kCG>
__device__ __forceinline__ void issue_khalf(uint32_t d_base, uint32_t a_stage, int hk, uint32_t bs,
                                            uint32_t plane_bytes, uint32_t N, bool first) {
  constexpr uint32_t MSEL = (uint32_t)((kCG * TM) >> 4) << 24;
  const uint32_t mmask = ~(31u << 24);
  auto mma_tf32 = [&](uint32_t a, uint64_t b, uint32_t acc) {
    const uint32_t id = (make_idesc(N) & mmask) | MSEL;
    if (kCG == 1) tc_mma_ts(d_base, a, b, id, acc); else tc_mma2_ts(d_base, a, b, id, acc);
  };
  auto mma_16 = [&](uint32_t a, uint64_t b, uint32_t id0, uint32_t acc) {
    const uint32_t id = (id0 & mmask) | MSEL;
    if (kCG == 1) tc_mma_ts_f16(d_base, a, b, id, acc); else tc_mma2_ts_f16(d_base, a, b, id, acc);
  };
  const uint32_t acc0 = first ? 0u : 1u;
  if (kSplit == 2) {
    // A words of this half: [0,8) = y_hi, [8,16) = y_lo; projection row: [0,32 B) = q_hi, [32,64 B) = q_lo
    const uint64_t desc = make_b_desc(bs);
    const uint32_t a_hi = a_stage + (uint32_t)(hk * TKB), a_lo = a_hi + 8;
    mma_16(a_lo, desc, make_idesc_f16(N), acc0);        // y_lo . q_hi
    mma_16(a_hi, desc + 2, make_idesc_f16(N), 1u);      // y_hi . q_lo
    mma_16(a_hi, desc, make_idesc_f16(N), 1u);          // y_hi . q_hi
  } else {
    const uint64_t desc_hi = make_b_desc(bs);
    const uint64_t desc_lo = make_b_desc(bs + plane_bytes);
    const uint32_t a_hi = a_stage, a_lo = a_stage + 32;
#pragma unroll
    for (int s = 0; s < TKB / 8; ++s) {
      const uint32_t ka = (uint32_t)(hk * TKB + 8 * s);
      const uint32_t acc = (first && s == 0) ? 0u : 1u;
      if (kSplit == 0) {
        mma_tf32(a_lo + ka, desc_hi + (uint64_t)(2 * s), acc);   // x_lo . r_hi
        mma_tf32(a_hi + ka, desc_lo + (uint64_t)(2 * s), 1u);    // x_hi . r_lo
      } else {
        // s = 0: bf(x_lo) . bf(r_hi), s = 1: bf(x_hi) . bf(r_lo), each over the 16 k of this half
        mma_16(a_lo + ka, desc_lo + (uint64_t)(2 * s), make_idesc_bf16(N), acc);
      }
    }
#pragma unroll
    for (int s = 0; s < TKB / 8; ++s)                            // x_hi . r_hi
      mma_tf32(a_hi + (uint32_t)(hk * TKB + 8 * s), desc_hi + (uint64_t)(2 * s), 1u);
  }
}

// kSplit == 0: 3xTF32 (tm_rlo = TF32 residual plane).
// kSplit == 1: TF32 hi.hi + BF16 cross terms (tm_rlo = the cross plane, see split_cross_kernel):
//   x.r ~= [bf(x_lo) | bf(x_hi)] . [bf(r_hi) ; bf(r_lo)]   (kind::f16, K = 16 per MMA: 32 BF16 per 16 k)
//          + x_hi . r_hi                                    (kind::tf32, K = 8)
//   i.e. two tensor-time units per product instead of three.  The BF16 roundings act on terms that are
//   already 2^-11 down: worst case 4 * 2^-8 * 2^-11 = 2^-17 relative to sum|x_i r_i|, measured max
//   1.3e-7 * |x||r| on Gaussian data -- the same as an fp32 sgemm, two orders inside the 1e-5 margin.
template <int kSplit>
__global__ void __launch_bounds__(TC_THREADS_2CONV, 1)
hash_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_rhi,
               const __grid_constant__ CUtensorMap tm_rlo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[XS_MAX * 2 + BS * 2 + AS * 2 + 4];
  __shared__ uint32_t tmem_base_slot;

  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B: 1024 B alignment
  uint32_t bar0 = smem_u32(bars);
  // opaque to the compiler from here on: otherwise it re-derives the shared-window addresses (S2R
  // SR_CgaCtaId + LEA) in front of every mbarrier operation of the role loops instead of keeping them
  asm volatile("" : "+r"(smem_base), "+r"(bar0));
  const uint32_t x_smem = smem_base;
  const uint32_t XS = (uint32_t)p.xs;
  constexpr uint32_t kPlanes = (kSplit == 2) ? 1u : 2u;              // FP16x3: q_hi | q_lo share one 64 B row
  const uint32_t B_HALF_BYTES = (uint32_t)p.ncols_pass * TKB * 4u;   // one plane of one stage: N x 64 B
  const uint32_t B_STAGE_BYTES = kPlanes * B_HALF_BYTES;
  const uint32_t b_smem = smem_base + XS * X_STAGE_BYTES;
  auto x_full = [&](uint32_t i) { return bar0 + 8u * i; };
  auto x_empty = [&](uint32_t i) { return bar0 + 8u * (XS_MAX + i); };
  auto b_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + i); };
  auto b_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + BS + i); };
  auto a_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS + i); };
  auto a_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS + AS + i); };
  auto d_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS + 2 * AS + i); };
  auto d_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS + 2 * AS + 2 + i); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  __shared__ uint32_t repack_sm[TM * 9];  // per-thread 8 words (+1 zero) of compact column bits
  // per-row flags of a tile (bit 0: some |x| > 1e-8, bit 1: recompute in FP32), converters -> epilogue:
  // written by each converter group before the a_full arrive of its last chunk of the tile, read by the
  // epilogue after d_full.  8 slots: the converters run at most 4 chunks + 2 accumulator stages ahead.
  __shared__ uint8_t conv_flags[8][2][TM];
  // One or two converter warpgroups (launch with TC_THREADS or TC_THREADS_2CONV).  With two, the groups
  // take alternate K chunks, so two load -> split -> tcgen05.st -> wait chains are in flight per SM: for
  // shapes with few K chunks per tile (128 -> 64 bits: HBM-bound) the converters, not the tensor pipe, set
  // the pace (ncu: half of their time was per-chunk barrier / fence overhead).  Both groups still wait for
  // every chunk (which keeps their x_empty arrivals in the right mbarrier phase) and derive the vector's
  // FP16 scale from the same chunk, so they agree on it without talking to each other.
  const uint32_t ngroups = (blockDim.x > (unsigned)TC_THREADS) ? 2u : 1u;
  const uint32_t N = (uint32_t)p.ncols_pass;
  const uint32_t dstages = (N <= (uint32_t)TN) ? 2u : 1u;
  const int64_t work_items = p.mtiles * p.npass;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_rhi);
    tma_prefetch_desc(&tm_rlo);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < XS_MAX; ++i) { mbar_init(x_full(i), 1); mbar_init(x_empty(i), 4 * ngroups); }
    for (uint32_t i = 0; i < BS; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (uint32_t i = 0; i < AS; ++i) { mbar_init(a_full(i), 4); mbar_init(a_empty(i), 1); }   // the chunk's owner group
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(d_full(i), 1); mbar_init(d_empty(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loops, one elected lane issues) ====
    Ring xr, br;
    if (p.b_resident && (int64_t)blockIdx.x < work_items) {
      // small shape: load every K slice of R_hi / R_lo once; stage (kc, hk) lives at index 2*kc + hk
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full(0), (uint32_t)p.kc * (TK / TKB) * B_STAGE_BYTES);
        for (int kc = 0; kc < p.kc; ++kc)
          for (int hk = 0; hk < TK / TKB; ++hk) {
            const uint32_t dst = b_smem + (uint32_t)(kc * (TK / TKB) + hk) * B_STAGE_BYTES;
            if (kPlanes == 2) tma_load_2d(&tm_rhi, b_full(0), dst, kc * TK + hk * TKB, 0);
            tma_load_2d(&tm_rlo, b_full(0), dst + (kPlanes - 1) * B_HALF_BYTES, kc * TK + hk * TKB, 0);
          }
      }
      __syncwarp();
    }
    for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x) {
      const int64_t mt = (p.npass == 1) ? w : w / p.npass;   // (no 64-bit division on the common path)
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      const int row0 = (int)(mt * TM);
      for (int kc = 0; kc < p.kc; ++kc) {
        mbar_wait(x_empty(xr.idx), xr.phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(x_full(xr.idx), X_STAGE_BYTES);
          tma_load_2d(&tm_x, x_full(xr.idx), x_smem + xr.idx * X_STAGE_BYTES, kc * TK, row0);
        }
        __syncwarp();
        xr.advance(XS);
        for (int hk = 0; hk < ((p.b_resident || (p.flags & TC_FLAG_B_WARP)) ? 0 : TK / TKB); ++hk) {   // two 16-float halves of the chunk
          const int col0 = pass * (int)N;
          mbar_wait(b_empty(br.idx), br.phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(b_full(br.idx), B_STAGE_BYTES);
            const uint32_t dst = b_smem + br.idx * B_STAGE_BYTES;
            if (kPlanes == 2) tma_load_2d(&tm_rhi, b_full(br.idx), dst, kc * TK + hk * TKB, col0);
            tma_load_2d(&tm_rlo, b_full(br.idx), dst + (kPlanes - 1) * B_HALF_BYTES, kc * TK + hk * TKB, col0);
          }
          __syncwarp();
          br.advance(BS);
        }
      }
    }
  } else if (warp == 3 && (p.flags & TC_FLAG_B_WARP) && !p.b_resident) {
    // ===================== projection producer: its own warp, so that the X prefetch depth (XS
    // stages) is not throttled by the projection ring (b_empty follows the MMAs closely) ==========
    Ring br;
    for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x) {
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      const int col0 = pass * (int)N;
      for (int kc = 0; kc < p.kc; ++kc) {
        for (int hk = 0; hk < TK / TKB; ++hk) {
          mbar_wait(b_empty(br.idx), br.phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(b_full(br.idx), B_STAGE_BYTES);
            const uint32_t dst = b_smem + br.idx * B_STAGE_BYTES;
            if (kPlanes == 2) tma_load_2d(&tm_rhi, b_full(br.idx), dst, kc * TK + hk * TKB, col0);
            tma_load_2d(&tm_rlo, b_full(br.idx), dst + (kPlanes - 1) * B_HALF_BYTES, kc * TK + hk * TKB, col0);
          }
          __syncwarp();
          br.advance(BS);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loops, one elected lane issues) =====
    Ring ar, br, dr;
    if (p.b_resident && (int64_t)blockIdx.x < work_items) {
      mbar_wait(b_full(0), 0);
      tc_fence_after();
    }
    for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x) {
      mbar_wait(d_empty(dr.idx), dr.phase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_base = tmem_base + dr.idx * TN;  // dstages == 2 only when N <= 128
      for (int kc = 0; kc < p.kc; ++kc) {
        mbar_wait(a_full(ar.idx), ar.phase);
        const uint32_t a_hi = tmem_base + A_COL0 + ar.idx * A_STAGE_COLS;
        for (int hk = 0; hk < TK / TKB; ++hk) {
          if (!p.b_resident) {
            mbar_wait(b_full(br.idx), br.phase);
            tc_fence_after();
          }
          const uint32_t bs =
              b_smem + (p.b_resident ? (uint32_t)(kc * (TK / TKB) + hk) : br.idx) * B_STAGE_BYTES;
          if (elect_one()) {
            issue_khalf<kSplit, 1>(d_base, a_hi, hk, bs, B_HALF_BYTES, N, kc == 0 && hk == 0);
            if (!p.b_resident) tc_commit(b_empty(br.idx));  // same thread as the MMAs (commit tracks its own ops)
            if (hk == TK / TKB - 1) tc_commit(a_empty(ar.idx));
            if (hk == TK / TKB - 1 && kc == p.kc - 1) tc_commit(d_full(dr.idx));
          }
          __syncwarp();
          br.advance(BS);
        }
        ar.advance(AS);
      }
      dr.advance(dstages);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ===================== converters: X fp32 (smem) -> split A operand (TMEM) =================
    const uint32_t grp = (warp >= 12) ? 1u : 0u;
    const int t = (warp & 3) * 32 + lane;                 // row within the tile == TMEM lane
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    Ring xr, ar;
    uint32_t tcount = 0;   // tiles (work items) this CTA has converted
    uint32_t seq = 0;      // running chunk number: chunk seq belongs to group (seq & 1)
    for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x) {
      bool viol = false;  // some |x| > 1e-8 (or NaN): not a zero vector
      bool redo = false;  // FP16x3: this vector does not fit the scaled FP16 range
      float sc = 0.f;     // FP16x3: the vector's power-of-two scale (0 until a non-zero chunk is met)
      float mx = 0.f;     // FP16x3: max |x| over the chunks this group converted (NaN sticks)
      for (int kc = 0; kc < p.kc; ++kc, ++seq) {
        const bool mine = (ngroups == 1u) || ((seq & 1u) == grp);
        mbar_wait(x_full(xr.idx), xr.phase);
        const uint32_t row = x_smem + xr.idx * X_STAGE_BYTES + (uint32_t)t * 128u;
        if (kSplit == 2 && sc == 0.f && !redo) {
          const float m = chunk_absmax(row, t);   // both groups look at the same chunk: same scale
          if (m > 0.f) {
            sc = row_scale_for(m);
            redo = (sc == 0.f);
          }
        }
        if (mine) {
          mbar_wait(a_empty(ar.idx), ar.phase ^ 1);
          tc_fence_after();
          const uint32_t a_dst = tmem_base + lane_field + A_COL0 + ar.idx * A_STAGE_COLS;
#pragma unroll
          for (int h = 0; h < 2; ++h) {   // two halves of 16 floats keep the register count down
            if (kSplit == 2) {
              uint32_t wd[16];
              convert_half_f16(row, t, h, sc, wd, mx);
              tc_st16(a_dst + h * 16, wd);
            } else {
              uint32_t hi[16], lo[16];
              convert_half<kSplit>(row, t, h, hi, lo, viol);
              tc_st16(a_dst + h * 16, hi);
              tc_st16(a_dst + 32 + h * 16, lo);
            }
          }
          tc_wait_st();
          // this group's last chunk of the tile: hand its share of the row flags to the epilogue (the
          // a_full arrive below releases the store; the epilogue reads after d_full)
          if (kc + (int)ngroups >= p.kc) {
            if (kSplit == 2) {
              viol = row_viol(mx);
              redo |= row_overflow(mx, sc);
            }
            conv_flags[tcount & 7u][grp][t] = (uint8_t)((viol ? 1 : 0) | (redo ? 2 : 0));
          }
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) {
          if (mine) mbar_arrive(a_full(ar.idx));
          mbar_arrive(x_empty(xr.idx));
        }
        xr.advance(XS);
        ar.advance(AS);
      }
      ++tcount;
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================== epilogue: accumulators -> sign bits -> signature bytes ==============
    const int t = (warp - 8) * 32 + lane;
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    Ring dr;
    uint32_t tcount = 0;
    for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x, ++tcount) {
      const int64_t mt = (p.npass == 1) ? w : w / p.npass;   // (no 64-bit division on the common path)
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      mbar_wait(d_full(dr.idx), dr.phase);
      tc_fence_after();
      // row flags from the converter group(s) that own chunks of this tile
      const uint32_t first_owner = (tcount * (uint32_t)p.kc) & 1u;
      uint32_t fl = 0;
      if (ngroups == 1u || p.kc >= 2 || first_owner == 0u) fl |= conv_flags[tcount & 7u][0][t];
      if (ngroups == 2u && (p.kc >= 2 || first_owner == 1u)) fl |= conv_flags[tcount & 7u][1][t];
      uint32_t words[8];
      if (p.dbg_acc != nullptr && w == 0)
        dump_accumulator_row(tmem_base + lane_field + dr.idx * TN, N, p.dbg_acc + (size_t)t * N);
      read_sign_words(tmem_base + lane_field + dr.idx * TN, N, words);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(dr.idx));
      dr.advance(dstages);

      const int64_t m = mt * TM + t;
      store_signature(p, N, pass, m, t, words, repack_sm);
      if (pass == 0) {
        if (p.zero_flag != nullptr && m < p.n) p.zero_flag[m] = (fl & 1u) ? 0 : 1;
        if (kSplit == 2) {
          const unsigned any = __ballot_sync(0xffffffffu, (fl & 2u) != 0u && m < p.n);
          if (any != 0u && lane == 0) p.redo_list[atomicAdd(p.redo_count, 1)] = (int)mt;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// =================================================================================================
// 2-CTA variant (cta_group::2): two SMs of one TPC work on a 256-row tile.  Each CTA converts its own
// 128 rows into its own TMEM and keeps its own 128 x N accumulator rows, but stages only HALF of the
// projection columns of a pass in its shared memory; the pair's tensor cores read both halves.  That
// halves the projection bytes every SM pulls from L2 per tile -- the limit of the 1-CTA kernel once
// the split needs only two tensor-time units (measured: 80 KB per K chunk per SM = 10.9 TB/s of
// L2->SM traffic at 724 M vectors/s, against a ~12 TB/s fabric cap) -- and the shared-memory reads
// of the MMAs.  Only the leader CTA (cluster rank 0) issues MMAs; barriers that gate them (a_full,
// b_full, d_empty) live in the leader and are arrived on across the cluster, barriers that the MMAs
// release (a_empty, b_empty, d_full) are signalled in both CTAs by a multicast tcgen05.commit.
// =================================================================================================
template <int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
hash_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_rhi,
                const __grid_constant__ CUtensorMap tm_rlo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[XS_MAX * 2 + BS2 * 2 + AS * 2 + 4];
  __shared__ uint32_t tmem_base_slot;

  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // same offset in both CTAs
  uint32_t bar0 = smem_u32(bars);
  asm volatile("" : "+r"(smem_base), "+r"(bar0));   // keep both in registers (see hash_tc_kernel)
  const uint32_t x_smem = smem_base;
  const uint32_t XS = (uint32_t)p.xs;
  const uint32_t N = (uint32_t)p.ncols_pass;
  const uint32_t NH = N / 2;                                // projection columns staged by this CTA
  constexpr uint32_t kPlanes = (kSplit == 2) ? 1u : 2u;
  const uint32_t B_HALF_BYTES = NH * TKB * 4u;              // one plane of one stage: N/2 rows x 64 B
  const uint32_t B_STAGE_BYTES = kPlanes * B_HALF_BYTES;
  const uint32_t b_smem = smem_base + XS * X_STAGE_BYTES;
  auto x_full = [&](uint32_t i) { return bar0 + 8u * i; };
  auto x_empty = [&](uint32_t i) { return bar0 + 8u * (XS_MAX + i); };
  auto b_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + i); };
  auto b_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + BS2 + i); };
  auto a_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS2 + i); };
  auto a_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS2 + AS + i); };
  auto d_full = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS2 + 2 * AS + i); };
  auto d_empty = [&](uint32_t i) { return bar0 + 8u * (2 * XS_MAX + 2 * BS2 + 2 * AS + 2 + i); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t dstages = (N <= (uint32_t)TN) ? 2u : 1u;
  const int64_t work_items = p.mtiles * p.npass;            // mtiles = 256-row tiles here

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_rhi);
    tma_prefetch_desc(&tm_rlo);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < XS_MAX; ++i) { mbar_init(x_full(i), 1); mbar_init(x_empty(i), 4); }
    for (uint32_t i = 0; i < BS2; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
    for (uint32_t i = 0; i < AS; ++i) { mbar_init(a_full(i), 8); mbar_init(a_empty(i), 1); }   // 4 warps x 2 CTAs
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(d_full(i), 1); mbar_init(d_empty(i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // both CTAs take part in the pair-wide allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated, before anything crosses
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== X producer (this CTA's 128 rows of the pair's tile) ==================
    Ring xr;
    for (int64_t w = pair; w < work_items; w += npairs) {
      const int64_t mt = (p.npass == 1) ? w : w / p.npass;
      const int row0 = (int)(mt * (2 * TM) + rank * TM);
      for (int kc = 0; kc < p.kc; ++kc) {
        mbar_wait(x_empty(xr.idx), xr.phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(x_full(xr.idx), X_STAGE_BYTES);
          tma_load_2d(&tm_x, x_full(xr.idx), x_smem + xr.idx * X_STAGE_BYTES, kc * TK, row0);
        }
        __syncwarp();
        xr.advance(XS);
      }
    }
  } else if (warp == 3) {
    // ===================== projection producer: this CTA's half of the pass's columns ===========
    Ring br;
    const uint32_t half = rank;   // the leader's shared memory feeds accumulator columns [0, N/2)
    for (int64_t w = pair; w < work_items; w += npairs) {
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      const int col0 = pass * (int)N + (int)(half * NH);
      for (int kc = 0; kc < p.kc; ++kc) {
        for (int hk = 0; hk < TK / TKB; ++hk) {
          mbar_wait(b_empty(br.idx), br.phase ^ 1);
          if (elect_one()) {
            // the leader's barrier counts the bytes of both CTAs' loads
            if (rank == 0) mbar_arrive_expect_tx(b_full(br.idx), 2u * B_STAGE_BYTES);
            const uint32_t full0 = mapa_rank(b_full(br.idx), 0);
            const uint32_t dst = b_smem + br.idx * B_STAGE_BYTES;
            if (kPlanes == 2) tma_load_2d_cg2(&tm_rhi, full0, dst, kc * TK + hk * TKB, col0);
            tma_load_2d_cg2(&tm_rlo, full0, dst + (kPlanes - 1) * B_HALF_BYTES, kc * TK + hk * TKB, col0);
          }
          __syncwarp();
          br.advance(BS2);
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only): M = 256 over the pair ===================
    Ring ar, br, dr;
    for (int64_t w = pair; w < work_items; w += npairs) {
      mbar_wait(d_empty(dr.idx), dr.phase ^ 1);   // both epilogues have drained this stage
      tc_fence_after();
      const uint32_t d_base = tmem_base + dr.idx * TN;
      for (int kc = 0; kc < p.kc; ++kc) {
        mbar_wait(a_full(ar.idx), ar.phase);      // both CTAs' converters
        const uint32_t a_hi = tmem_base + A_COL0 + ar.idx * A_STAGE_COLS;
        for (int hk = 0; hk < TK / TKB; ++hk) {
          mbar_wait(b_full(br.idx), br.phase);
          tc_fence_after();
          const uint32_t bs = b_smem + br.idx * B_STAGE_BYTES;
          if (elect_one()) {
            issue_khalf<kSplit, 2>(d_base, a_hi, hk, bs, B_HALF_BYTES, N, kc == 0 && hk == 0);
            tc_commit_pair(b_empty(br.idx));
            if (hk == TK / TKB - 1) tc_commit_pair(a_empty(ar.idx));
            if (hk == TK / TKB - 1 && kc == p.kc - 1) tc_commit_pair(d_full(dr.idx));
          }
          __syncwarp();
          br.advance(BS2);
        }
        ar.advance(AS);
      }
      dr.advance(dstages);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== converters (as in the 1-CTA kernel; a_full lives in the leader) =======
    const int t = (warp - 4) * 32 + lane;
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    Ring xr, ar;
    for (int64_t w = pair; w < work_items; w += npairs) {
      const int64_t mt = (p.npass == 1) ? w : w / p.npass;
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      bool viol = false, redo = false;
      float sc = 0.f, mx = 0.f;
      for (int kc = 0; kc < p.kc; ++kc) {
        mbar_wait(x_full(xr.idx), xr.phase);
        const uint32_t row = x_smem + xr.idx * X_STAGE_BYTES + (uint32_t)t * 128u;
        if (kSplit == 2 && sc == 0.f && !redo) {
          const float m = chunk_absmax(row, t);
          if (m > 0.f) {
            sc = row_scale_for(m);
            redo = (sc == 0.f);
          }
        }
        mbar_wait(a_empty(ar.idx), ar.phase ^ 1);
        tc_fence_after();
        const uint32_t a_dst = tmem_base + lane_field + A_COL0 + ar.idx * A_STAGE_COLS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (kSplit == 2) {
            uint32_t wd[16];
            convert_half_f16(row, t, h, sc, wd, mx);
            tc_st16(a_dst + h * 16, wd);
          } else {
            uint32_t hi[16], lo[16];
            convert_half<kSplit>(row, t, h, hi, lo, viol);
            tc_st16(a_dst + h * 16, hi);
            tc_st16(a_dst + 32 + h * 16, lo);
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_rank(a_full(ar.idx), 0));
          mbar_arrive(x_empty(xr.idx));
        }
        xr.advance(XS);
        ar.advance(AS);
      }
      if (kSplit == 2) {
        viol = row_viol(mx);
        redo |= row_overflow(mx, sc);
      }
      if (pass == 0) {
        const int64_t m = mt * (2 * TM) + rank * TM + t;
        if (p.zero_flag != nullptr && m < p.n) p.zero_flag[m] = viol ? 0 : 1;
        if (kSplit == 2) {
          const unsigned any = __ballot_sync(0xffffffffu, redo && m < p.n);
          if (any != 0u && lane == 0) p.redo_list[atomicAdd(p.redo_count, 1)] = (int)(mt * 2 + rank);
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: this CTA's 128 accumulator rows =============================
    const int t = (warp - 8) * 32 + lane;
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    Ring dr;
    for (int64_t w = pair; w < work_items; w += npairs) {
      const int64_t mt = (p.npass == 1) ? w : w / p.npass;
      const int pass = (p.npass == 1) ? 0 : (int)(w % p.npass);
      mbar_wait(d_full(dr.idx), dr.phase);
      tc_fence_after();
      uint32_t words[8];
      if (p.dbg_acc != nullptr && w == 0)
        dump_accumulator_row(tmem_base + lane_field + dr.idx * TN, N, p.dbg_acc + (size_t)(rank * TM + t) * N);
      read_sign_words(tmem_base + lane_field + dr.idx * TN, N, words);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(d_empty(dr.idx), 0));
      dr.advance(dstages);
      store_signature(p, N, pass, mt * (2 * TM) + rank * TM + t, t, words, nullptr);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nothing may still target the peer's barriers / shared memory / TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// Rp [.][dim] fp32 -> Rhi / Rlo [rows][dim_pad] TF32-rounded split; row i of the output is row
// rowmap[i] of Rp (or a zero row when rowmap[i] < 0); K is zero-padded to dim_pad.
__global__ void split_projections_kernel(const float* __restrict__ Rp, const int* __restrict__ rowmap,
                                         float* __restrict__ hi, float* __restrict__ lo, int rows, int dim,
                                         int dim_pad) {
  const int64_t total = (int64_t)rows * dim_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / dim_pad), k = (int)(i % dim_pad);
    const int src = rowmap[row];
    float h = 0.f, l = 0.f;
    if (k < dim && src >= 0) {
      const float r = Rp[(int64_t)src * dim + k];
      h = __uint_as_float(tf32_rna(__float_as_uint(r)));
      const float res = (fabsf(h) == INFINITY) ? 0.f : (r - h);
      l = __uint_as_float(tf32_rna(__float_as_uint(res)));
    }
    hi[i] = h;
    lo[i] = l;
  }
}

// Cross plane of the TF32+BF16 split, [rows][dim_pad] 32-bit words.  The 16 words of K block g
// (k in [16g, 16g+16)) of a row hold 32 BF16: words [0,8) = bf(r_hi[k]) pairs, words [8,16) =
// bf(r_lo[k]) pairs (even k in the low half), r_hi = TF32(r), r_lo = r - r_hi -- one 64 B SWIZZLE_64B
// row = two K=16 kind::f16 steps, matched by the converters' A words [bf(x_lo) | bf(x_hi)].
__global__ void split_cross_kernel(const float* __restrict__ Rp, const int* __restrict__ rowmap,
                                   uint32_t* __restrict__ cross, int rows, int dim, int dim_pad) {
  const int64_t total = (int64_t)rows * dim_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / dim_pad), wd = (int)(i % dim_pad);
    const int src = rowmap[row];
    const int g = wd / 16, j = wd % 16;
    const int k0 = 16 * g + 2 * (j & 7);
    float v[2] = {0.f, 0.f};
    for (int e = 0; e < 2; ++e) {
      const int k = k0 + e;
      if (k < dim && src >= 0) {
        const float r = Rp[(int64_t)src * dim + k];
        const uint32_t hu = tf32_rna(__float_as_uint(r));
        const float h = __uint_as_float(hu);
        const float res = (fabsf(h) == INFINITY) ? 0.f : (r - h);
        v[e] = (j < 8) ? (((hu & 0x7FFFFFFFu) >= 0x7F7F8000u) ? 0.f : h) : res;
      }
    }
    cross[i] = pack_bf16(v[0], v[1]);
  }
}

// FP16x3 plane, [rows][dim_pad] 32-bit words, one block per row.  q = s_r * r with the power of two s_r
// that puts the row's largest |r| in [2^13, 2^14); the 16 words of K block g hold 32 FP16: words
// [0,8) = q_hi pairs, [8,16) = q_lo pairs (q_hi = FP16(q), q_lo = FP16(q - q_hi), even k in the low half).
__global__ void split_f16_kernel(const float* __restrict__ Rp, const int* __restrict__ rowmap,
                                 uint32_t* __restrict__ plane, int dim, int dim_pad, int* __restrict__ bad_rows) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int src = rowmap[row];
  const float* r = (src >= 0) ? Rp + (int64_t)src * dim : nullptr;
  float m = 0.f;
  if (r != nullptr)
    for (int k = threadIdx.x; k < dim; k += blockDim.x) m = fmaxf(m, fabsf(r[k]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  const int e = (int)((__float_as_uint(m) >> 23) & 0xFFu);
  const int se = 254 + 13 - e;
  const bool representable = (e >= 13 && e < 255 && se >= 1);
  const float sc = representable ? __uint_as_float((uint32_t)se << 23) : 1.f;
  // a NON-ZERO row without a representable scale (largest |r| below 2^-114, infinite, or NaN -- only
  // reachable through hasher.projections = [...]) would be flushed by FP16: the plan then refuses the FP16x3
  // arm and the hasher runs the FP32 kernel instead (tc_plan_create), never silently wrong bits
  bool has_nan = false;
  if (r != nullptr)
    for (int k = threadIdx.x; k < dim; k += blockDim.x) has_nan |= (r[k] != r[k]);
  if (__syncthreads_or((int)has_nan) || (!representable && m != 0.f)) {
    if (threadIdx.x == 0) atomicAdd(bad_rows, 1);
  }
  for (int wd = threadIdx.x; wd < dim_pad; wd += blockDim.x) {
    const int g = wd / 16, j = wd % 16;
    const int k0 = 16 * g + 2 * (j & 7);
    float v[2] = {0.f, 0.f};
    for (int e2 = 0; e2 < 2; ++e2) {
      const int k = k0 + e2;
      if (r != nullptr && k < dim) {
        const float q = r[k] * sc;
        float h0, h1;
        unpack_f16(pack_f16(q, 0.f), h0, h1);
        v[e2] = (j < 8) ? q : (q - h0);
      }
    }
    plane[(int64_t)row * dim_pad + wd] = pack_f16(v[0], v[1]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D row-major fp32 tensor [rows][cols], box = box_cols floats x box_rows rows, swizzle span equal
// to the box row (128 B or 64 B), zero OOB fill
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
             uint32_t box_cols, uint32_t box_rows, CUtensorMapL2promotion promo) {
  EncodeTiledFn enc = get_encode_fn();
  LSHX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const CUtensorMapSwizzle swz = (box_cols * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu pitch %llu)", (int)rc,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes);
    return LSHX_ERR_CUDA;
  }
  return LSHX_OK;
}

}  // namespace

struct TcPlan {
  float* d_hi = nullptr;
  float* d_lo = nullptr;     // TF32 residual plane (3xTF32 arm)
  uint32_t* d_x = nullptr;   // BF16 cross plane (TF32+BF16 arm)
  uint32_t* d_h = nullptr;   // scaled FP16 hi|lo plane (FP16x3 arm)
  const float* d_Rp = nullptr;  // the hasher's padded FP32 projections (not owned): FP32 recomputation
  CUtensorMap tm_rhi, tm_rlo, tm_rx, tm_rh;
  CUtensorMap tm2_rhi, tm2_rlo, tm2_rx, tm2_rh;   // the same planes with boxes of ncols_pass / 2 rows (2-CTA kernel)
  bool cg2_ok = false;
  int flags = 0;
  int auto_split = 1;   // what split < 0 resolves to for this shape
  int num_sms = 0;
  // column layout of the split projections (see tc_plan_create)
  int ncols_pass = 0, npass = 0, repack = 0, bpp = 0;
  bool f16_ok = true;   // false: some projection row has no representable FP16 scale (split_f16_kernel)
  // FP16x3: handle-owned scratch for the lists of 128-row tiles to recompute in FP32.  A small ring, one
  // slot per launch in flight: [0] = counter (reset by the recompute kernel itself), [1] = its exit ticket,
  // [4 ..] = up to four entries per tile.  `done` orders a slot's next user after its last one when the two
  // launches are on different streams.
  struct RedoSlot {
    int* buf = nullptr;
    size_t tiles = 0;
    cudaEvent_t done = nullptr;
    cudaStream_t last = nullptr;
    bool used = false;
  };
  static constexpr int kRedoSlots = 4;
  RedoSlot redo[kRedoSlots];
  int next_redo = 0;
  float* dbg_acc = nullptr;   // diagnostics only (lshx_hasher_debug_accumulators): 256 x 256 floats when set
};

bool tc_shape_supported(const HashShape& s) {
  // TMA needs a 16-byte row pitch for X; everything else is padded by the plan
  return s.dim % 4 == 0 && get_encode_fn() != nullptr;
}

int tc_plan_create(const HashShape& s, const float* d_Rp, TcPlan** out) {
  *out = nullptr;
  TcPlan* pl = new TcPlan();
  auto fail = [&](int code) {
    tc_plan_destroy(pl);
    return code;
  };
  // Column layout.  rows_per_band % 8 == 0: the padded layout of lshx_common.cuh is already dense
  // (column c = signature bit c), passes of 256 (or 128) columns.  Otherwise computing the zero
  // padding columns would waste up to 8x of the tensor work (16 bands x 4 rows: 128 columns for 64
  // bits), so the columns are COMPACT -- pass p holds bands [p*bpp, (p+1)*bpp) back to back, padded
  // to a multiple of 16 columns with zero rows -- and the epilogue expands each band to bytes.
  std::vector<int> rowmap;
  if (s.rows_per_band % 8 == 0 || s.rows_per_band > 128) {
    const int ntiles = s.ncols_pad / TN;
    pl->ncols_pass = (ntiles % 2 == 0) ? 2 * TN : TN;
    pl->npass = s.ncols_pad / pl->ncols_pass;
    pl->repack = 0;
    rowmap.resize((size_t)s.ncols_pad);
    for (int i = 0; i < s.ncols_pad; ++i) rowmap[i] = i;
  } else {
    const int r = s.rows_per_band;
    pl->bpp = (256 / r < s.num_bands) ? 256 / r : s.num_bands;
    pl->npass = (s.num_bands + pl->bpp - 1) / pl->bpp;
    pl->ncols_pass = (pl->bpp * r + 15) / 16 * 16;
    pl->repack = 1;
    rowmap.assign((size_t)pl->npass * pl->ncols_pass, -1);
    for (int b = 0; b < s.num_bands; ++b)
      for (int j = 0; j < r; ++j)
        rowmap[(size_t)(b / pl->bpp) * pl->ncols_pass + (size_t)(b % pl->bpp) * r + j] = b * 8 * s.bpb + j;
  }
  const int rows = pl->npass * pl->ncols_pass;
  const size_t bytes = (size_t)rows * s.dim_pad * sizeof(float);
  int* d_rowmap = nullptr;
  if (cudaMalloc(&pl->d_hi, bytes) != cudaSuccess || cudaMalloc(&pl->d_lo, bytes) != cudaSuccess ||
      cudaMalloc(&pl->d_x, bytes) != cudaSuccess || cudaMalloc(&pl->d_h, bytes) != cudaSuccess ||
      cudaMalloc(&d_rowmap, rowmap.size() * sizeof(int)) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes for the split projections failed", bytes);
    (void)cudaGetLastError();
    if (d_rowmap) cudaFree(d_rowmap);
    return fail(LSHX_ERR_OOM);
  }
  int* d_bad = nullptr;
  cudaError_t serr = cudaMemcpy(d_rowmap, rowmap.data(), rowmap.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (serr == cudaSuccess) serr = cudaMalloc(&d_bad, sizeof(int));
  if (serr == cudaSuccess) serr = cudaMemset(d_bad, 0, sizeof(int));
  int bad_rows = 0;
  if (serr == cudaSuccess) {
    split_projections_kernel<<<256, 256>>>(d_Rp, d_rowmap, pl->d_hi, pl->d_lo, rows, s.dim, s.dim_pad);
    split_cross_kernel<<<256, 256>>>(d_Rp, d_rowmap, pl->d_x, rows, s.dim, s.dim_pad);
    split_f16_kernel<<<rows, 256>>>(d_Rp, d_rowmap, pl->d_h, s.dim, s.dim_pad, d_bad);
    count_launch(3);
    serr = cudaMemcpy(&bad_rows, d_bad, sizeof(int), cudaMemcpyDeviceToHost);   // synchronises
  }
  pl->d_Rp = d_Rp;
  cudaFree(d_rowmap);
  if (d_bad) cudaFree(d_bad);
  if (serr != cudaSuccess) {
    set_error("splitting the projections failed: %s", cudaGetErrorString(serr));
    (void)cudaGetLastError();
    return fail(LSHX_ERR_CUDA);
  }
  pl->f16_ok = (bad_rows == 0);
  const uint64_t pitch = (uint64_t)s.dim_pad * sizeof(float);
  const uint32_t brows = (uint32_t)pl->ncols_pass;  // columns per pass = rows of one projection box
  int rc = make_map(&pl->tm_rhi, pl->d_hi, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, brows,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc != LSHX_OK) return fail(rc);
  rc = make_map(&pl->tm_rlo, pl->d_lo, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, brows,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc != LSHX_OK) return fail(rc);
  rc = make_map(&pl->tm_rx, pl->d_x, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, brows,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc != LSHX_OK) return fail(rc);
  rc = make_map(&pl->tm_rh, pl->d_h, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, brows,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc != LSHX_OK) return fail(rc);
  pl->cg2_ok = !pl->repack && pl->ncols_pass % 32 == 0;
  if (pl->cg2_ok) {
    const uint32_t hrows = brows / 2;
    if ((rc = make_map(&pl->tm2_rhi, pl->d_hi, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, hrows,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != LSHX_OK ||
        (rc = make_map(&pl->tm2_rlo, pl->d_lo, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, hrows,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != LSHX_OK ||
        (rc = make_map(&pl->tm2_rx, pl->d_x, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, hrows,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != LSHX_OK ||
        (rc = make_map(&pl->tm2_rh, pl->d_h, (uint64_t)rows, (uint64_t)s.dim_pad, pitch, TKB, hrows,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != LSHX_OK)
      return fail(rc);
  }
  pl->flags = TC_FLAG_B_WARP | TC_FLAG_CG2;
  if (const char* e = getenv("LSHX_TC_FLAGS")) if (*e) pl->flags = atoi(e);
  pl->auto_split = 2;   // scaled FP16x3: fastest on every measured shape (768/256, 1536/512, 128/64)
  if (const char* e = getenv("LSHX_TC_SPLIT")) if (*e) pl->auto_split = atoi(e);
  {
    // keep freed scratch of the stream-ordered allocator cached instead of returning it at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, 0) == cudaSuccess) {
      int dev_now = 0;
      cudaGetDevice(&dev_now);
      if (cudaDeviceGetDefaultMemPool(&pool, dev_now) == cudaSuccess) {
        uint64_t keep = 64ull << 20;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
    }
    (void)cudaGetLastError();
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(hash_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(hash_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(hash_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(hash_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(hash_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess ||
      cudaFuncSetAttribute(hash_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(SMEM_BYTES + 1024)) != cudaSuccess) {
    set_error("cannot reserve %u bytes of shared memory for the tcgen05 kernel", SMEM_BYTES + 1024);
    (void)cudaGetLastError();
    return fail(LSHX_ERR_CUDA);
  }
  *out = pl;
  return LSHX_OK;
}

bool tc_plan_f16_ok(const TcPlan* p) { return p != nullptr && p->f16_ok; }
int tc_plan_default_split(const TcPlan* p) { return p ? p->auto_split : -1; }
int tc_plan_flags(const TcPlan* p) { return p ? p->flags : 0; }

int tc_plan_set_debug(TcPlan* p, bool on) {
  if (on && p->dbg_acc == nullptr) {
    LSHX_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->dbg_acc), 2 * TM * 2 * TN * sizeof(float)));
    LSHX_CUDA(cudaMemset(p->dbg_acc, 0, 2 * TM * 2 * TN * sizeof(float)));
  } else if (!on && p->dbg_acc != nullptr) {
    cudaFree(p->dbg_acc);
    p->dbg_acc = nullptr;
  }
  return LSHX_OK;
}
const float* tc_plan_debug_buffer(const TcPlan* p, int* cols) {
  if (cols) *cols = p->ncols_pass;
  return p->dbg_acc;
}

void tc_plan_destroy(TcPlan* p) {
  if (!p) return;
  for (auto& slot : p->redo) {
    if (slot.buf) cudaFree(slot.buf);
    if (slot.done) cudaEventDestroy(slot.done);
  }
  if (p->dbg_acc) cudaFree(p->dbg_acc);
  if (p->d_hi) cudaFree(p->d_hi);
  if (p->d_lo) cudaFree(p->d_lo);
  if (p->d_x) cudaFree(p->d_x);
  if (p->d_h) cudaFree(p->d_h);
  delete p;
}

int launch_hash_tc(const HashShape& s, TcPlan* plan, int split, const float* d_X, int64_t n, uint8_t* d_out,
                   uint8_t* d_zero_flag, cudaStream_t stream) {
  if (n <= 0) return LSHX_OK;
  if (split < 0) split = plan->auto_split;
  LSHX_REQUIRE(split >= 0 && split <= 2, "unknown split %d", split);
  LSHX_REQUIRE((reinterpret_cast<uintptr_t>(d_X) & 15) == 0, "tcgen05 kernel needs 16-byte aligned vectors");
  LSHX_REQUIRE(n < (1ll << 31), "hash batch of %lld rows exceeds one launch", (long long)n);
  CUtensorMap tm_x;
  int rc = make_map(&tm_x, d_X, (uint64_t)n, (uint64_t)s.dim, (uint64_t)s.dim * sizeof(float), TK, TM,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc != LSHX_OK) return rc;
  const uint32_t planes = (split == 2) ? 1u : 2u;   // FP16x3 keeps q_hi | q_lo in one 64 B row
  TcParams p;
  p.n = n;
  p.kc = s.dim_pad / TK;
  p.ncols_pass = plan->ncols_pass;
  {
    const uint32_t b_stage = planes * (uint32_t)plan->ncols_pass * TKB * 4u;
    const uint32_t b_all = (uint32_t)p.kc * (TK / TKB) * b_stage;  // the whole split matrix of one pass
    p.b_resident = (plan->npass == 1 && b_all <= SMEM_BYTES - 4 * X_STAGE_BYTES) ? 1 : 0;
    const uint32_t b_bytes = p.b_resident ? b_all : BS * b_stage;
    const uint32_t fit = (SMEM_BYTES - b_bytes) / X_STAGE_BYTES;
    p.xs = (int)(fit < (uint32_t)XS_MAX ? fit : (uint32_t)XS_MAX);
  }
  p.npass = plan->npass;
  p.repack = plan->repack;
  p.r = s.rows_per_band;
  p.bpb = s.bpb;
  p.bpp = plan->bpp;
  p.num_bands = s.num_bands;
  p.mtiles = (n + TM - 1) / TM;
  p.sig_bytes = s.sig_bytes;
  p.out_vec_ok = (s.sig_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0) ? 1 : 0;
  p.out = d_out;
  p.zero_flag = d_zero_flag;
  p.flags = plan->flags;
  p.redo_count = nullptr;
  p.redo_list = nullptr;
  p.dbg_acc = plan->dbg_acc;
  // FP16x3: a handle-owned scratch slot for the list of 128-row tiles to recompute in FP32 (a counter, then
  // up to four entries -- one per converter warp -- for every tile); no allocation on the launch path
  TcPlan::RedoSlot* slot = nullptr;
  if (split == 2) {
    LSHX_REQUIRE(plan->f16_ok, "a projection row has no representable FP16 scale: the FP16x3 arm cannot be "
                               "used with these projections (LSHX_KERNEL_AUTO takes the FP32 kernel)");
    slot = &plan->redo[plan->next_redo];
    plan->next_redo = (plan->next_redo + 1) % TcPlan::kRedoSlots;
    const size_t tiles = (size_t)p.mtiles;
    if (slot->done == nullptr) LSHX_CUDA(cudaEventCreateWithFlags(&slot->done, cudaEventDisableTiming));
    if (slot->tiles < tiles) {
      if (slot->used) LSHX_CUDA(cudaEventSynchronize(slot->done));   // its last launch still reads the old buffer
      if (slot->buf) cudaFree(slot->buf);
      slot->buf = nullptr;
      slot->tiles = 0;
      const size_t want = tiles < 4096 ? 4096 : tiles + tiles / 4;
      if (cudaMalloc(reinterpret_cast<void**>(&slot->buf), 16 + sizeof(int) * 4 * want) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc of %zu bytes for the FP32-recompute tile list failed", 16 + sizeof(int) * 4 * want);
        return LSHX_ERR_OOM;
      }
      LSHX_CUDA(cudaMemset(slot->buf, 0, 16));
      slot->tiles = want;
      slot->used = false;
    }
    if (slot->used && slot->last != stream) LSHX_CUDA(cudaStreamWaitEvent(stream, slot->done, 0));
    p.redo_count = slot->buf;
    p.redo_list = slot->buf + 4;
  }
  // 2-CTA kernel: streamed projections, byte-aligned bands, enough 256-row tiles to fill every SM pair
  const int npairs = plan->num_sms / 2;
  if ((p.flags & TC_FLAG_CG2) && plan->cg2_ok && !p.b_resident && npairs > 0 &&
      (((n + 2 * TM - 1) / (2 * TM)) * p.npass >= npairs || (p.flags & TC_FLAG_CG2_ALWAYS))) {
    p.mtiles = (n + 2 * TM - 1) / (2 * TM);
    const uint32_t b_bytes = BS2 * planes * (uint32_t)(plan->ncols_pass / 2) * TKB * 4u;   // per CTA: half the columns
    const uint32_t fit = (SMEM_BYTES - b_bytes) / X_STAGE_BYTES;
    p.xs = (int)(fit < (uint32_t)XS_MAX ? fit : (uint32_t)XS_MAX);
    const int64_t work2 = p.mtiles * p.npass;
    const unsigned grid2 = 2u * (unsigned)(work2 < npairs ? work2 : npairs);
    if (split == 0)
      hash_tc2_kernel<0><<<grid2, TC_THREADS, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm2_rhi, plan->tm2_rlo, p);
    else if (split == 1)
      hash_tc2_kernel<1><<<grid2, TC_THREADS, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm2_rhi, plan->tm2_rx, p);
    else
      hash_tc2_kernel<2><<<grid2, TC_THREADS, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm2_rh, plan->tm2_rh, p);
  } else {
    const int64_t work = p.mtiles * p.npass;
    const unsigned grid = (unsigned)(work < plan->num_sms ? work : plan->num_sms);
    // two converter warpgroups where a tile is a few short MMAs (resident projections: narrow, HBM-bound shapes)
    const bool conv2 = !(p.flags & TC_FLAG_CONV1) && (p.b_resident || (p.flags & TC_FLAG_CONV2));
    const unsigned threads = conv2 ? TC_THREADS_2CONV : TC_THREADS;
    if (split == 0)
      hash_tc_kernel<0><<<grid, threads, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm_rhi, plan->tm_rlo, p);
    else if (split == 1)
      hash_tc_kernel<1><<<grid, threads, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm_rhi, plan->tm_rx, p);
    else
      hash_tc_kernel<2><<<grid, threads, SMEM_BYTES + 1024, stream>>>(tm_x, plan->tm_rh, plan->tm_rh, p);
  }
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  if (split == 2) {
    // vectors outside the scaled FP16 range (normally none: the CTAs read a zero counter and return)
    // (the recompute kernel resets the slot's counter when its last CTA leaves)
    rc = launch_hash_ffma_tiles(s, d_X, n, plan->d_Rp, d_out, p.redo_list, p.redo_count, 2 * plan->num_sms, stream);
    if (rc != LSHX_OK) return rc;
    LSHX_CUDA(cudaEventRecord(slot->done, stream));
    slot->last = stream;
    slot->used = true;
  }
  return LSHX_OK;
}

}  // namespace lshx
