// extern "C" entry points of liblshx.so (declared in include/lshx.h) and the host-side
// staging pipelines around the kernels.

#include <emmintrin.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <thread>
#include <algorithm>
#include <vector>

#include <cuda_fp16.h>

#include "lshx_common.cuh"

namespace lshx {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device available (%s); liblshx has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    (void)cudaGetLastError();
    return LSHX_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    set_error("device %d out of range (have %d)", device, count);
    return LSHX_ERR_INVALID_ARG;
  }
  cudaDeviceProp prop;
  LSHX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; liblshx is built for sm_100a only", device, prop.major,
              prop.minor);
    return LSHX_ERR_NO_DEVICE;
  }
  return LSHX_OK;
}

// RAII device switch
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return LSHX_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    LSHX_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return LSHX_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace lshx

using namespace lshx;

// ---------------------------------------------------------------------------------------
// hasher
// ---------------------------------------------------------------------------------------

struct lshx_hasher {
  int device = 0;
  HashShape s{};
  float* d_Rp = nullptr;  // [ncols_pad][dim] fp32, zero rows for padding
  TcPlan* tc = nullptr;
  int kernel_pref = LSHX_KERNEL_AUTO;
  int last_kernel = 0;
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t ev_in = nullptr;
  DevBuf x_stage[2], out_stage[2], flag_stage[2];
  DevBuf raw_stage[2];         // typed host batches (fp16 / int8 / uint8) before the on-device cast
  // small-batch path (per-vector calls): pinned, device-mapped staging for up to small_rows rows
  int small_rows = 0;
  float* pin_x = nullptr;      // host, pinned + mapped (the fused latency kernel of the band index reads it in place)
  uint8_t* pin_out = nullptr;  // host, pinned + mapped: the kernel stores signatures / flags into it
  float* d_small_x = nullptr;
  // pageable host batches: pinned bounce buffers filled by several CPU threads (see hash_pageable)
  void* bounce_x[2] = {nullptr, nullptr};
  uint8_t* bounce_out[2] = {nullptr, nullptr};
  cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
  size_t bounce_rows = 0;
  std::mutex mu;
};

static int upload_projections(lshx_hasher* h, const float* R) {
  const HashShape& s = h->s;
  // build the padded column layout on the host: band b -> columns [b*8*bpb, b*8*bpb + r)
  std::vector<float> Rp((size_t)s.ncols_pad * s.dim, 0.f);
  for (int b = 0; b < s.num_bands; ++b)
    for (int j = 0; j < s.rows_per_band; ++j)
      std::memcpy(&Rp[((size_t)b * 8 * s.bpb + j) * s.dim],
                  &R[((size_t)b * s.rows_per_band + j) * s.dim], sizeof(float) * s.dim);
  LSHX_CUDA(cudaMemcpy(h->d_Rp, Rp.data(), Rp.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (h->tc) {
    tc_plan_destroy(h->tc);
    h->tc = nullptr;
  }
  if (tc_shape_supported(s)) {
    int rc = tc_plan_create(s, h->d_Rp, &h->tc);
    if (rc != LSHX_OK) return rc;
  }
  return LSHX_OK;
}

extern "C" int lshx_abi_version(void) { return LSHX_ABI_VERSION; }
extern "C" const char* lshx_last_error(void) { return g_err; }
extern "C" uint64_t lshx_launch_count(void) { return g_launches.load(); }

extern "C" int lshx_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < count; ++d) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
  }
  return ok;
}

extern "C" int lshx_hasher_create(int device, int dim, int num_bands, int rows_per_band,
                                  const float* projections_host, lshx_hasher** out) {
  LSHX_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  LSHX_REQUIRE(num_bands > 0, "num_bands must be > 0");
  LSHX_REQUIRE(rows_per_band > 0, "rows_per_band must be > 0");
  LSHX_REQUIRE(dim > 0, "dim must be > 0");
  LSHX_REQUIRE(projections_host != nullptr, "projections_host is null");
  const int64_t bpb = (rows_per_band + 7) / 8;
  LSHX_REQUIRE(bpb * num_bands <= (1 << 20), "signature of %lld bytes per vector is unsupported",
               (long long)(bpb * num_bands));
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);

  lshx_hasher* h = new lshx_hasher();
  h->device = device;
  HashShape& s = h->s;
  s.dim = dim;
  s.num_bands = num_bands;
  s.rows_per_band = rows_per_band;
  s.bpb = (int)bpb;
  s.sig_bytes = (int)(bpb * num_bands);
  s.ncols = s.sig_bytes * 8;
  s.ncols_pad = (s.ncols + 127) / 128 * 128;
  s.dim_pad = (dim + 31) / 32 * 32;
  auto fail = [&](int code) {
    lshx_hasher_destroy(h);
    return code;
  };
  if (cudaMalloc(&h->d_Rp, (size_t)s.ncols_pad * dim * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes for the projections failed",
              (size_t)s.ncols_pad * dim * sizeof(float));
    (void)cudaGetLastError();
    return fail(LSHX_ERR_OOM);
  }
  for (int i = 0; i < 2; ++i)
    if (cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      return fail(LSHX_ERR_CUDA);
    }
  if (cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess) {
    set_error("cudaEventCreate failed");
    return fail(LSHX_ERR_CUDA);
  }
  rc = upload_projections(h, projections_host);
  if (rc != LSHX_OK) return fail(rc);
  h->small_rows = hash_small_max_rows(s);
  if (h->small_rows > 0) {
    const size_t xb = (size_t)h->small_rows * dim * sizeof(float);
    const size_t ob = (size_t)h->small_rows * (s.sig_bytes + 1);
    if (cudaHostAlloc(&h->pin_x, xb, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostAlloc(&h->pin_out, ob, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaMalloc(&h->d_small_x, xb) != cudaSuccess) {
      (void)cudaGetLastError();
      h->small_rows = 0;  // the chunked path still works without the staging buffers
    }
  }
  *out = h;
  return LSHX_OK;
}

extern "C" int lshx_hasher_set_projections(lshx_hasher* h, const float* projections_host) {
  LSHX_REQUIRE(h != nullptr && projections_host != nullptr, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  LSHX_CUDA(cudaDeviceSynchronize());
  return upload_projections(h, projections_host);
}

extern "C" int lshx_hasher_set_kernel(lshx_hasher* h, int kernel) {
  LSHX_REQUIRE(h != nullptr, "null handle");
  const bool is_tc = kernel == LSHX_KERNEL_TCGEN05 || kernel == LSHX_KERNEL_TCGEN05_3XTF32 ||
                     kernel == LSHX_KERNEL_TCGEN05_TF32BF16;
  LSHX_REQUIRE(kernel == LSHX_KERNEL_AUTO || kernel == LSHX_KERNEL_FFMA || is_tc, "unknown kernel %d", kernel);
  std::lock_guard<std::mutex> lk(h->mu);   // h->tc is rebuilt by lshx_hasher_set_projections under the same lock
  LSHX_REQUIRE(!is_tc || h->tc != nullptr,
               "the tcgen05 kernel does not support this shape (dim %d, %d x %d)", h->s.dim,
               h->s.num_bands, h->s.rows_per_band);
  LSHX_REQUIRE(kernel != LSHX_KERNEL_TCGEN05 || tc_plan_f16_ok(h->tc),
               "a projection row has no representable FP16 scale (largest |r| below 2^-114, infinite or NaN): "
               "the scaled FP16x3 arm cannot hash with these projections");
  h->kernel_pref = kernel;
  return LSHX_OK;
}

extern "C" int lshx_hasher_last_kernel(const lshx_hasher* h) { return h ? h->last_kernel : 0; }
extern "C" int lshx_hasher_signature_bytes(const lshx_hasher* h) { return h ? h->s.sig_bytes : 0; }

// Copy into a pinned bounce buffer with NON-TEMPORAL stores.  A plain memcpy of a 1 MB piece leaves the lines
// dirty in the copying core's cache; the DMA engine that reads the buffer next then has to snoop them out of a
// dozen private caches, which runs at ~6.6 GB/s instead of PCIe's 55 (measured: a 12 MB H2D took 1.86 ms after the
// pool's memcpy, 0.24 ms from an untouched buffer -- tools/h2d_probe.py, LSHX_TRACE_PAGEABLE).  Streaming stores
// send the data to DRAM and keep the caches clean.
static void stream_copy(void* dst, const void* src, size_t bytes) {
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  if (bytes < 4096) {
    std::memcpy(d, s, bytes);
    return;
  }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
  std::memcpy(d, s, head);
  d += head; s += head; bytes -= head;
  const size_t blocks = bytes / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i a2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i a3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a0);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), a1);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), a2);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), a3);
    d += 64; s += 64;
  }
  std::memcpy(d, s, bytes - blocks * 64);
  _mm_sfence();
}

// memcpy split over the host cores: one core copies ~5-10 GB/s, PCIe Gen5 takes 55.  A small process-wide
// pool of sleeping workers (spawning threads per 64 MB chunk cost a third of the copy time).
namespace {
class CopyPool {
 public:
  explicit CopyPool(unsigned workers) {
    for (unsigned i = 0; i < workers; ++i) threads_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  unsigned workers() const { return (unsigned)threads_.size(); }
  // copy [src, src + bytes) to dst in `parts` pieces; the caller copies one piece itself
  void copy(void* dst, const void* src, size_t bytes, unsigned parts) {
    std::lock_guard<std::mutex> serial(call_mu_);   // one parallel copy at a time
    const size_t piece = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      piece_ = piece;
      next_ = 1;                       // piece 0 is the caller's
      total_ = (unsigned)((bytes + piece - 1) / piece);
      pending_ = total_ > 0 ? total_ - 1 : 0;
      ++generation_;
    }
    cv_.notify_all();
    stream_copy(dst, src, piece < bytes ? piece : bytes);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [&] { return stop_ || (generation_ != seen && next_ < total_); });
      if (stop_) return;
      while (next_ < total_) {
        const unsigned i = next_++;
        const size_t off = (size_t)i * piece_;
        const size_t len = (off + piece_ <= bytes_) ? piece_ : bytes_ - off;
        char* d = dst_ + off;
        const char* s = src_ + off;
        lk.unlock();
        stream_copy(d, s, len);
        lk.lock();
        if (--pending_ == 0) done_cv_.notify_all();
      }
      seen = generation_;
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  bool stop_ = false;
  uint64_t generation_ = 0;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, piece_ = 0;
  unsigned next_ = 0, total_ = 0, pending_ = 0;
};
}  // namespace

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  // 12 of 16 cores measured best for one process (48 GB/s).  Under torchrun every rank has its own pool:
  // share the cores between the LOCAL_WORLD_SIZE processes of this host (8 ranks x 12 threads on 32 cores
  // collapsed the pageable path to 2.5 M vectors/s per rank) and leave one for the thread that drives CUDA
  unsigned hw = std::thread::hardware_concurrency();
  unsigned local_ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 1024) local_ranks = (unsigned)v;
  }
  unsigned share = hw / local_ranks;
  unsigned nt = share >= 4 ? (share * 3 / 4 > 12 ? 12 : share * 3 / 4) : (share >= 2 ? share - 1 : 1);
  if (const char* e = getenv("LSHX_COPY_THREADS")) {   // tuning knob
    const int v = atoi(e);
    if (v >= 1 && v <= 64) nt = (unsigned)v;
  }
  if (bytes < (2u << 20) || nt <= 1) {
    stream_copy(dst, src, bytes);
    return;
  }
  static CopyPool* pool = new CopyPool(nt - 1);   // leaked on purpose: no join at process exit
  const unsigned parts = pool->workers() + 1 < nt ? pool->workers() + 1 : nt;
  pool->copy(dst, src, bytes, parts);
}

static bool is_pageable_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

// Host -> device upload of a buffer that may be ordinary pageable memory (numpy arrays: query vectors, candidate
// ids, signatures).  cudaMemcpyAsync from pageable memory runs at ~10 GB/s through the driver's own staging; a
// large pageable source instead goes through two pinned bounce buffers per device, filled with non-temporal
// stores (stream_copy, see above) while the previous piece is on the wire: ~45 GB/s.  The source is consumed
// when the call returns; the DMA is asynchronous on `st`.
namespace {
struct UploadRing {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ok = false;
};
constexpr size_t UPLOAD_PIECE = 8u << 20;
std::mutex g_upload_mu;
UploadRing g_upload_rings[64];
}  // namespace

static int upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return LSHX_OK;
  int dev = -1;
  if (bytes < (1u << 20) || !is_pageable_host(h_src) || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    LSHX_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    return LSHX_OK;
  }
  std::lock_guard<std::mutex> lk(g_upload_mu);
  UploadRing& ring = g_upload_rings[dev];
  if (!ring.ok) {
    for (int i = 0; i < 2; ++i) {
      if (cudaHostAlloc(&ring.buf[i], UPLOAD_PIECE, cudaHostAllocPortable) != cudaSuccess ||
          cudaEventCreateWithFlags(&ring.ev[i], cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();   // no bounce buffers: the plain copy still works
        LSHX_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return LSHX_OK;
      }
    }
    ring.ok = true;
  }
  const char* src = static_cast<const char*>(h_src);
  char* dst = static_cast<char*>(d_dst);
  int slot = 0;
  for (size_t off = 0; off < bytes; off += UPLOAD_PIECE, slot ^= 1) {
    const size_t len = bytes - off < UPLOAD_PIECE ? bytes - off : UPLOAD_PIECE;
    LSHX_CUDA(cudaEventSynchronize(ring.ev[slot]));   // the DMA that last read this slot is done (a fresh event is)
    parallel_memcpy(ring.buf[slot], src + off, len);
    LSHX_CUDA(cudaMemcpyAsync(dst + off, ring.buf[slot], len, cudaMemcpyHostToDevice, st));
    LSHX_CUDA(cudaEventRecord(ring.ev[slot], st));
  }
  return LSHX_OK;
}

static int launch_hash(lshx_hasher* h, const float* d_X, int64_t n, uint8_t* d_out,
                       uint8_t* d_flag, cudaStream_t st);

// ---- typed host batches: the cast to float32 that LSHHasher.hash_batch does with np.asarray(vectors,
// float32) (reference lshrs/hash/lsh.py:162) happens on the device, so fp16 embeddings cross PCIe at half
// and uint8 / int8 descriptors at a quarter of the float32 bytes.  Every one of these conversions is exact.
static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
static __device__ __forceinline__ float to_f32(uint8_t v) { return (float)v; }
static __device__ __forceinline__ float to_f32(int8_t v) { return (float)v; }

template <typename T>
__global__ void expand_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += stride) {
    if (i + 4 <= count) {   // both staging buffers are 256-byte aligned and i % 4 == 0
      T v[4];
      if (sizeof(T) == 1) *reinterpret_cast<uint32_t*>(v) = *reinterpret_cast<const uint32_t*>(in + i);
      else *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(in + i);
      *reinterpret_cast<float4*>(out + i) = make_float4(to_f32(v[0]), to_f32(v[1]), to_f32(v[2]), to_f32(v[3]));
    } else {
      for (int64_t j = i; j < count; ++j) out[j] = to_f32(in[j]);
    }
  }
}

static size_t dtype_size(int dtype) {
  switch (dtype) {
    case LSHX_DTYPE_F32: return 4;
    case LSHX_DTYPE_F16: return 2;
    case LSHX_DTYPE_U8:
    case LSHX_DTYPE_I8: return 1;
    default: return 0;
  }
}

static int expand_chunk(int dtype, const void* d_raw, float* d_x, int64_t count, cudaStream_t st) {
  const int threads = 256;
  int64_t want = (count / 4 + threads - 1) / threads;
  const int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  switch (dtype) {
    case LSHX_DTYPE_F16:
      expand_to_f32_kernel<<<blocks, threads, 0, st>>>(static_cast<const __half*>(d_raw), d_x, count);
      break;
    case LSHX_DTYPE_U8:
      expand_to_f32_kernel<<<blocks, threads, 0, st>>>(static_cast<const uint8_t*>(d_raw), d_x, count);
      break;
    case LSHX_DTYPE_I8:
      expand_to_f32_kernel<<<blocks, threads, 0, st>>>(static_cast<const int8_t*>(d_raw), d_x, count);
      break;
    default:
      set_error("unsupported dtype %d", dtype);
      return LSHX_ERR_INVALID_ARG;
  }
  count_launch();
  LSHX_CUDA(cudaGetLastError());
  return LSHX_OK;
}

// Large host batch that is not pinned float32: an ordinary (pageable) numpy array, and / or a typed one.
// cudaMemcpyAsync from pageable memory is staged by
// the driver on one thread (~11 GB/s measured).  Instead the rows go through two pinned bounce buffers
// filled by parallel_memcpy while the previous chunk is in flight; signatures and flags come back through
// pinned bounce buffers too and are copied out when their slot is reused.
static int hash_pageable(lshx_hasher* h, const void* Xv, int dtype, int64_t n, uint8_t* out, uint8_t* zero_flag) {
  const HashShape& s = h->s;
  const char* X = static_cast<const char*>(Xv);
  const size_t in_row = (size_t)s.dim * dtype_size(dtype);      // bytes of one row as the caller holds it
  const size_t row_bytes = (size_t)s.dim * sizeof(float);
  const size_t out_row = (size_t)s.sig_bytes + 1;  // signature bytes + 1 flag byte per row
  if (h->bounce_rows == 0) {
    size_t bounce_mb = 64;
    if (const char* e = getenv("LSHX_BOUNCE_MB")) {       // tuning knob
      const int v = atoi(e);
      if (v >= 1 && v <= 1024) bounce_mb = (size_t)v;
    }
    size_t rows = (bounce_mb << 20) / row_bytes;
    rows = rows / 128 * 128;
    if (rows < 128) rows = 128;
    for (int i = 0; i < 2; ++i) {
      if (cudaHostAlloc(&h->bounce_x[i], rows * row_bytes, cudaHostAllocDefault) != cudaSuccess ||
          // (4x the rows: a chunk of 1-byte elements holds four times as many rows in the same bounce_x bytes)
          cudaHostAlloc(reinterpret_cast<void**>(&h->bounce_out[i]), 4 * rows * out_row, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->bounce_ev[i], cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cannot allocate pinned bounce buffers");
        return LSHX_ERR_OOM;
      }
    }
    h->bounce_rows = rows;
  }
  // rows per chunk: what fills a bounce buffer in the caller's element type (float32: bounce_rows;
  // float16: twice, 1-byte types: four times as many -- the per-chunk fixed costs are per byte moved)
  int64_t chunk = (int64_t)(h->bounce_rows * row_bytes / in_row) / 128 * 128;
  // a batch that does not fill eight bounce buffers is cut in eight anyway (pieces of at least 2 MB), so that the
  // copy of one piece into its bounce buffer overlaps the DMA of the piece before it
  {
    int64_t quarter = ((n + 7) / 8 + 127) / 128 * 128;
    const int64_t min_rows = ((int64_t)((2u << 20) / in_row) + 127) / 128 * 128;
    if (quarter < min_rows) quarter = min_rows;
    if (quarter < chunk) chunk = quarter;
  }
  for (int i = 0; i < 2; ++i) {
    int rc;
    if ((rc = h->x_stage[i].reserve((size_t)chunk * row_bytes)) != LSHX_OK) return rc;
    if ((rc = h->out_stage[i].reserve((size_t)chunk * s.sig_bytes)) != LSHX_OK) return rc;
    if ((rc = h->flag_stage[i].reserve((size_t)chunk)) != LSHX_OK) return rc;
    if (dtype != LSHX_DTYPE_F32 && (rc = h->raw_stage[i].reserve((size_t)chunk * in_row)) != LSHX_OK) return rc;
  }
  struct Pending { int64_t r0 = 0, rows = 0; } pending[2];
  auto drain = [&](int slot) -> int {  // copy a finished slot's results to the caller's memory
    if (pending[slot].rows == 0) return LSHX_OK;
    LSHX_CUDA(cudaStreamSynchronize(h->streams[slot]));   // (the slot's stream carries nothing but this chunk)
    const Pending pd = pending[slot];
    std::memcpy(out + pd.r0 * s.sig_bytes, h->bounce_out[slot], (size_t)pd.rows * s.sig_bytes);
    if (zero_flag) std::memcpy(zero_flag + pd.r0, h->bounce_out[slot] + (size_t)chunk * s.sig_bytes, (size_t)pd.rows);
    pending[slot].rows = 0;
    return LSHX_OK;
  };
  int slot = 0;
  static const bool trace = getenv("LSHX_TRACE_PAGEABLE") != nullptr;   // bring-up: phase times on stderr
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = trace ? now_us() : 0.0;
  double t_copy = 0.0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk, slot ^= 1) {
    const int64_t rows = (n - r0 < chunk) ? (n - r0) : chunk;
    int rc = drain(slot);  // also guarantees the slot's H2D source is no longer being read
    if (rc != LSHX_OK) return rc;
    const double tc0 = trace ? now_us() : 0.0;
    parallel_memcpy(h->bounce_x[slot], X + r0 * in_row, (size_t)rows * in_row);
    if (trace) t_copy += now_us() - tc0;
    cudaStream_t st = h->streams[slot];
    if (dtype == LSHX_DTYPE_F32) {
      LSHX_CUDA(cudaMemcpyAsync(h->x_stage[slot].p, h->bounce_x[slot], (size_t)rows * row_bytes,
                                cudaMemcpyHostToDevice, st));
    } else {
      LSHX_CUDA(cudaMemcpyAsync(h->raw_stage[slot].p, h->bounce_x[slot], (size_t)rows * in_row,
                                cudaMemcpyHostToDevice, st));
      rc = expand_chunk(dtype, h->raw_stage[slot].p, static_cast<float*>(h->x_stage[slot].p), rows * s.dim, st);
      if (rc != LSHX_OK) return rc;
    }
    uint8_t* d_o = static_cast<uint8_t*>(h->out_stage[slot].p);
    uint8_t* d_f = zero_flag ? static_cast<uint8_t*>(h->flag_stage[slot].p) : nullptr;
    double t_a = 0, t_b = 0, t_c = 0;
    if (trace) { t_a = now_us(); cudaStreamSynchronize(st); t_b = now_us(); }
    rc = launch_hash(h, static_cast<const float*>(h->x_stage[slot].p), rows, d_o, d_f, st);
    if (rc != LSHX_OK) return rc;
    if (trace) {
      cudaStreamSynchronize(st);
      t_c = now_us();
      fprintf(stderr, "[lshx pageable]   rows %lld: H2D %.0f us, kernel %.0f us\n", (long long)rows, t_b - t_a, t_c - t_b);
    }
    LSHX_CUDA(cudaMemcpyAsync(h->bounce_out[slot], d_o, (size_t)rows * s.sig_bytes, cudaMemcpyDeviceToHost, st));
    if (zero_flag)
      LSHX_CUDA(cudaMemcpyAsync(h->bounce_out[slot] + (size_t)chunk * s.sig_bytes, d_f, (size_t)rows,
                                cudaMemcpyDeviceToHost, st));
    LSHX_CUDA(cudaEventRecord(h->bounce_ev[slot], st));
    pending[slot].r0 = r0;
    pending[slot].rows = rows;
  }
  const double t_issued = trace ? now_us() : 0.0;
  int rc = drain(slot);
  if (rc != LSHX_OK) return rc;
  rc = drain(slot ^ 1);
  if (trace)
    fprintf(stderr, "[lshx pageable] n=%lld chunk=%lld: total %.0f us, bounce memcpy %.0f us, issue %.0f us, drain %.0f us\n",
            (long long)n, (long long)chunk, now_us() - t_begin, t_copy, t_issued - t_begin - t_copy, now_us() - t_issued);
  return rc;
}

static int launch_hash(lshx_hasher* h, const float* d_X, int64_t n, uint8_t* d_out,
                       uint8_t* d_flag, cudaStream_t st) {
  // operand split of the tcgen05 kernel: < 0 = the plan's default (scaled FP16x3)
  const int split = h->kernel_pref == LSHX_KERNEL_TCGEN05_3XTF32 ? 0
                    : h->kernel_pref == LSHX_KERNEL_TCGEN05_TF32BF16 ? 1 : -1;
  // AUTO: tcgen05 (scaled FP16x3) for 16-byte aligned vectors, unless a projection row has no representable
  // FP16 scale -- then the FP32 kernel, which has numpy's semantics for any finite or non-finite plane
  const bool auto_tc = h->kernel_pref == LSHX_KERNEL_AUTO && (reinterpret_cast<uintptr_t>(d_X) & 15) == 0 &&
                       (tc_plan_f16_ok(h->tc) || tc_plan_default_split(h->tc) != 2);
  const bool use_tc = h->tc != nullptr && (h->kernel_pref == LSHX_KERNEL_TCGEN05 || split >= 0 || auto_tc);
  h->last_kernel = use_tc ? (split >= 0 ? h->kernel_pref : LSHX_KERNEL_TCGEN05) : LSHX_KERNEL_FFMA;
  // one launch takes < 2^31 rows (TMA coordinates / grid size are 32-bit): split larger batches
  const int64_t piece = 1ll << 30;
  for (int64_t r0 = 0; r0 < n; r0 += piece) {
    const int64_t rows = (n - r0 < piece) ? (n - r0) : piece;
    const float* x = d_X + r0 * h->s.dim;
    uint8_t* o = d_out + r0 * h->s.sig_bytes;
    uint8_t* f = d_flag ? d_flag + r0 : nullptr;
    const int rc = use_tc ? launch_hash_tc(h->s, h->tc, split, x, rows, o, f, st)
                          : launch_hash_ffma(h->s, x, rows, h->d_Rp, o, f, st);
    if (rc != LSHX_OK) return rc;
  }
  return LSHX_OK;
}

extern "C" int lshx_hash_batch(lshx_hasher* h, const float* X, int64_t n, int x_is_device,
                               uint8_t* out, int out_is_device, uint8_t* zero_flag,
                               void* stream) {
  LSHX_REQUIRE(h != nullptr, "null handle");
  LSHX_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return LSHX_OK;
  LSHX_REQUIRE(X != nullptr && out != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const HashShape& s = h->s;
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);

  // ---- everything on the device: one asynchronous launch ----------------------------
  if (x_is_device && out_is_device) {
    return launch_hash(h, X, n, out, zero_flag, user);  // NULL = the default stream
  }

  // ---- a few host rows (LSHRS.ingest / query hash ONE vector per call): latency path ------------
  if (!x_is_device && !out_is_device && n <= h->small_rows && h->kernel_pref == LSHX_KERNEL_AUTO) {
    cudaStream_t st = h->streams[0];
    const size_t xb = (size_t)n * s.dim * sizeof(float);
    uint8_t* d_out_map = nullptr;
    LSHX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_out_map), h->pin_out, 0));
    uint8_t* d_flag_map = zero_flag ? d_out_map + (size_t)h->small_rows * s.sig_bytes : nullptr;
    h->last_kernel = LSHX_KERNEL_SMALL;
    int rc;
    if (n == 1 && s.dim <= 1024) {   // the vector rides in the kernel's parameter block: no copy before the launch
      rc = launch_hash_small_one(s, X, h->d_Rp, d_out_map, d_flag_map, st);
    } else {
      std::memcpy(h->pin_x, X, xb);
      LSHX_CUDA(cudaMemcpyAsync(h->d_small_x, h->pin_x, xb, cudaMemcpyHostToDevice, st));
      rc = launch_hash_small(s, h->d_small_x, (int)n, h->d_Rp, d_out_map, d_flag_map, st);
    }
    if (rc != LSHX_OK) return rc;
    LSHX_CUDA(cudaStreamSynchronize(st));
    std::memcpy(out, h->pin_out, (size_t)n * s.sig_bytes);
    if (zero_flag) std::memcpy(zero_flag, h->pin_out + (size_t)h->small_rows * s.sig_bytes, (size_t)n);
    return LSHX_OK;
  }

  // ---- a large pageable host batch: pinned bounce buffers filled by several CPU threads -----------
  if (!x_is_device && !out_is_device && (size_t)n * s.dim * sizeof(float) >= (1u << 20) && is_pageable_host(X)) {
    LSHX_CUDA(cudaStreamSynchronize(user));
    return hash_pageable(h, X, LSHX_DTYPE_F32, n, out, zero_flag);
  }

  // ---- at least one side on the host: chunked, two streams, synchronous -----------
  const size_t row_bytes = (size_t)s.dim * sizeof(float);
  int64_t chunk = (int64_t)((256u << 20) / row_bytes);
  chunk = chunk / 128 * 128;
  if (chunk < 128) chunk = 128;
  if (chunk > n) chunk = n;
  for (int i = 0; i < 2; ++i) {
    int rc;
    if (!x_is_device && (rc = h->x_stage[i].reserve((size_t)chunk * row_bytes)) != LSHX_OK) return rc;
    if (!out_is_device) {
      if ((rc = h->out_stage[i].reserve((size_t)chunk * s.sig_bytes)) != LSHX_OK) return rc;
      if (zero_flag && (rc = h->flag_stage[i].reserve((size_t)chunk)) != LSHX_OK) return rc;
    }
  }
  if (x_is_device) {
    // inputs were produced on the caller's stream: order our streams after it
    LSHX_CUDA(cudaEventRecord(h->ev_in, user));
    for (int i = 0; i < 2; ++i) LSHX_CUDA(cudaStreamWaitEvent(h->streams[i], h->ev_in, 0));
  } else {
    LSHX_CUDA(cudaStreamSynchronize(user));  // host inputs: nothing to order, just drain the caller's stream
  }
  int slot = 0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk, slot ^= 1) {
    const int64_t rows = (n - r0 < chunk) ? (n - r0) : chunk;
    cudaStream_t st = h->streams[slot];
    const float* d_x;
    if (x_is_device) {
      d_x = X + r0 * s.dim;
    } else {
      LSHX_CUDA(cudaMemcpyAsync(h->x_stage[slot].p, X + r0 * s.dim, (size_t)rows * row_bytes,
                                cudaMemcpyHostToDevice, st));
      d_x = static_cast<const float*>(h->x_stage[slot].p);
    }
    uint8_t* d_o = out_is_device ? out + r0 * s.sig_bytes : static_cast<uint8_t*>(h->out_stage[slot].p);
    uint8_t* d_f = zero_flag ? (out_is_device ? zero_flag + r0
                                              : static_cast<uint8_t*>(h->flag_stage[slot].p))
                             : nullptr;
    int rc = launch_hash(h, d_x, rows, d_o, d_f, st);
    if (rc != LSHX_OK) return rc;
    if (!out_is_device) {
      LSHX_CUDA(cudaMemcpyAsync(out + r0 * s.sig_bytes, d_o, (size_t)rows * s.sig_bytes,
                                cudaMemcpyDeviceToHost, st));
      if (zero_flag)
        LSHX_CUDA(cudaMemcpyAsync(zero_flag + r0, d_f, (size_t)rows, cudaMemcpyDeviceToHost, st));
    }
  }
  LSHX_CUDA(cudaStreamSynchronize(h->streams[0]));
  LSHX_CUDA(cudaStreamSynchronize(h->streams[1]));
  return LSHX_OK;
}

extern "C" int lshx_hash_batch_typed(lshx_hasher* h, const void* X, int dtype, int64_t n, uint8_t* out,
                                     uint8_t* zero_flag) {
  LSHX_REQUIRE(h != nullptr, "null handle");
  LSHX_REQUIRE(n >= 0, "n must be >= 0");
  LSHX_REQUIRE(dtype_size(dtype) != 0, "unsupported dtype %d", dtype);
  if (n == 0) return LSHX_OK;
  LSHX_REQUIRE(X != nullptr && out != nullptr, "null buffer");
  if (dtype == LSHX_DTYPE_F32)
    return lshx_hash_batch(h, static_cast<const float*>(X), n, 0, out, 0, zero_flag, nullptr);
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  LSHX_CUDA(cudaStreamSynchronize(nullptr));
  return hash_pageable(h, X, dtype, n, out, zero_flag);
}

extern "C" int lshx_env_overrides(void) {
  static const char* const names[] = {"LSHX_TC_FLAGS", "LSHX_TC_SPLIT", "LSHX_COPY_THREADS", "LSHX_BOUNCE_MB",
                                      "LSHX_TRACE_PAGEABLE"};
  int mask = 0;
  for (int i = 0; i < 5; ++i) {
    const char* e = getenv(names[i]);
    if (e != nullptr && *e) mask |= 1 << i;
  }
  return mask;
}

extern "C" int lshx_hasher_debug_accumulators(lshx_hasher* h, const float* X_host, int64_t n, float* acc_out,
                                              int64_t acc_capacity, int* out_rows, int* out_cols) {
  LSHX_REQUIRE(h != nullptr && X_host != nullptr && acc_out != nullptr && out_rows && out_cols, "null argument");
  LSHX_REQUIRE(n > 0, "n must be > 0");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  LSHX_REQUIRE(h->tc != nullptr, "the tcgen05 kernel does not support this shape");
  const HashShape& s = h->s;
  int rc;
  if ((rc = h->x_stage[0].reserve((size_t)n * s.dim * sizeof(float))) != LSHX_OK) return rc;
  if ((rc = h->out_stage[0].reserve((size_t)n * s.sig_bytes)) != LSHX_OK) return rc;
  if ((rc = tc_plan_set_debug(h->tc, true)) != LSHX_OK) return rc;
  cudaStream_t st = h->streams[0];
  LSHX_CUDA(cudaMemcpyAsync(h->x_stage[0].p, X_host, (size_t)n * s.dim * sizeof(float), cudaMemcpyHostToDevice, st));
  const int pref = h->kernel_pref;
  if (pref == LSHX_KERNEL_AUTO || pref == LSHX_KERNEL_FFMA) h->kernel_pref = LSHX_KERNEL_TCGEN05;
  rc = launch_hash(h, static_cast<const float*>(h->x_stage[0].p), n, static_cast<uint8_t*>(h->out_stage[0].p),
                   nullptr, st);
  h->kernel_pref = pref;
  int cols = 0;
  const float* d_acc = tc_plan_debug_buffer(h->tc, &cols);
  // the 2-CTA variant works on 256-row tiles; report what the first tile held
  const int tile_rows = (n >= 256) ? 256 : 128;
  const int rows = (int)(n < tile_rows ? n : tile_rows);
  if (rc == LSHX_OK && (int64_t)rows * cols > acc_capacity) {
    set_error("acc_out holds %lld floats, %lld needed", (long long)acc_capacity, (long long)rows * cols);
    rc = LSHX_ERR_INVALID_ARG;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == LSHX_OK && e == cudaSuccess)
    e = cudaMemcpy(acc_out, d_acc, (size_t)rows * cols * sizeof(float), cudaMemcpyDeviceToHost);
  tc_plan_set_debug(h->tc, false);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(e);
  *out_rows = rows;
  *out_cols = cols;
  return LSHX_OK;
}

extern "C" int lshx_signatures_to_hex(const uint8_t* sig, int64_t n, int sig_bytes, char* hex_out) {
  LSHX_REQUIRE(sig != nullptr && hex_out != nullptr && n >= 0 && sig_bytes > 0, "bad argument");
  static const char digits[] = "0123456789abcdef";
  const int64_t total = n * sig_bytes;
  for (int64_t i = 0; i < total; ++i) {
    hex_out[2 * i] = digits[sig[i] >> 4];
    hex_out[2 * i + 1] = digits[sig[i] & 15];
  }
  return LSHX_OK;
}

extern "C" int lshx_hasher_destroy(lshx_hasher* h) {
  if (!h) return LSHX_OK;
  {
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    if (h->tc) tc_plan_destroy(h->tc);
    if (h->d_Rp) cudaFree(h->d_Rp);
    for (int i = 0; i < 2; ++i) {
      h->x_stage[i].release();
      h->raw_stage[i].release();
      h->out_stage[i].release();
      h->flag_stage[i].release();
      if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->pin_x) cudaFreeHost(h->pin_x);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    if (h->d_small_x) cudaFree(h->d_small_x);
    for (int i = 0; i < 2; ++i) {
      if (h->bounce_x[i]) cudaFreeHost(h->bounce_x[i]);
      if (h->bounce_out[i]) cudaFreeHost(h->bounce_out[i]);
      if (h->bounce_ev[i]) cudaEventDestroy(h->bounce_ev[i]);
    }
    (void)cudaGetLastError();
  }
  delete h;
  return LSHX_OK;
}

// ---------------------------------------------------------------------------------------
// reranker
// ---------------------------------------------------------------------------------------

struct lshx_reranker {
  int device = 0;
  int dim = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // host-buffer calls alternate query chunks on the two
  cudaEvent_t ev = nullptr;
  DevBuf q, vecs, offs, ids, pos, score, count, zero, all;
  DevBuf big;   // global sort scratch of oversized selections (rerank_scratch_bytes)
  std::mutex mu;
};

extern "C" int lshx_rerank_create(int device, int dim, lshx_reranker** out) {
  LSHX_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  LSHX_REQUIRE(dim > 0, "dim must be > 0");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);
  lshx_reranker* r = new lshx_reranker();
  r->device = device;
  r->dim = dim;
  if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&r->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->ev, cudaEventDisableTiming) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    (void)cudaGetLastError();
    lshx_rerank_destroy(r);
    return LSHX_ERR_CUDA;
  }
  *out = r;
  return LSHX_OK;
}

extern "C" int lshx_rerank_destroy(lshx_reranker* r) {
  if (!r) return LSHX_OK;
  {
    DeviceGuard g(r->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&r->q, &r->vecs, &r->offs, &r->ids, &r->pos, &r->score, &r->count, &r->zero, &r->all, &r->big})
      b->release();
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->stream2) cudaStreamDestroy(r->stream2);
    if (r->ev) cudaEventDestroy(r->ev);
    (void)cudaGetLastError();
  }
  delete r;
  return LSHX_OK;
}

// Shared body of lshx_rerank_topk / lshx_rerank_scores.
static int rerank_common(lshx_reranker* r, const float* Q, int64_t nq, const float* vectors,
                         int64_t n_vectors, const int64_t* cand_offsets, const int64_t* cand_ids,
                         int64_t max_candidates, int64_t total_candidates, bool select, int k,
                         double p, int out_stride, int32_t* out_pos, float* out_score,
                         int32_t* out_count, float* out_all, int32_t* out_zero, int on_device,
                         void* stream) {
  LSHX_REQUIRE(r != nullptr, "null handle");
  LSHX_REQUIRE(nq >= 0, "nq must be >= 0");
  LSHX_REQUIRE(on_device >= 0 && on_device <= 2, "on_device must be 0, 1 or 2");
  if (nq == 0) return LSHX_OK;
  LSHX_REQUIRE(Q != nullptr && cand_offsets != nullptr, "null buffer");
  if (select) {
    LSHX_REQUIRE(k > 0 || p > 0.0, "k must be > 0");
    LSHX_REQUIRE(!(p > 1.0), "top_p must be within the range (0, 1]");
    LSHX_REQUIRE(out_pos && out_score && out_count && out_stride > 0, "null output buffer");
  } else {
    LSHX_REQUIRE(out_all != nullptr, "null output buffer");
  }
  std::lock_guard<std::mutex> lk(r->mu);
  DeviceGuard g(r->device);
  const int dim = r->dim;
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);

  RerankArgs a{};
  a.nq = nq;
  a.n_vectors = n_vectors;
  a.dim = dim;
  a.k = k;
  a.p = p;
  a.out_stride = out_stride;
  a.select = select;

  if (on_device == 1) {
    LSHX_REQUIRE(vectors != nullptr || max_candidates == 0, "null vectors");
    a.Q = Q; a.V = vectors; a.offs = cand_offsets; a.ids = cand_ids;
    a.out_pos = out_pos; a.out_score = out_score; a.out_count = out_count; a.out_zero = out_zero;
    a.all_scores = out_all;
    a.max_cand = max_candidates;
    if (const size_t need = rerank_scratch_bytes(a, nullptr)) {
      // the scratch may still be in use by an earlier launch on another stream: drain before growing it
      if (need > r->big.cap) LSHX_CUDA(cudaDeviceSynchronize());
      int rc = r->big.reserve(need);
      if (rc != LSHX_OK) return rc;
      a.big_keys = static_cast<uint64_t*>(r->big.p);
      a.big_bytes = r->big.cap;
    }
    return launch_rerank(a, user);  // NULL = the default stream
  }

  // host-side offsets: validate and size
  int64_t total = cand_offsets[nq] - cand_offsets[0];
  int64_t maxc = 0;
  for (int64_t i = 0; i < nq; ++i) {
    const int64_t c = cand_offsets[i + 1] - cand_offsets[i];
    LSHX_REQUIRE(c >= 0, "cand_offsets must be non-decreasing");
    if (c > maxc) maxc = c;
  }
  LSHX_REQUIRE(cand_offsets[0] >= 0, "cand_offsets[0] must be >= 0");
  const int64_t end = cand_offsets[nq];
  LSHX_REQUIRE(total == 0 || (vectors != nullptr && n_vectors > 0), "no vectors to rank against");
  if (!cand_ids) LSHX_REQUIRE(end <= n_vectors, "cand_offsets run past the %lld packed vectors", (long long)n_vectors);
  if (!select) LSHX_REQUIRE(total_candidates >= end, "out_scores too small");
  a.max_cand = maxc;

  LSHX_CUDA(cudaStreamSynchronize(user));  // a device-resident corpus may still be being written there
  int rc;
  if ((rc = r->q.reserve((size_t)nq * dim * sizeof(float))) != LSHX_OK) return rc;
  if ((rc = r->offs.reserve((size_t)(nq + 1) * sizeof(int64_t))) != LSHX_OK) return rc;
  if (cand_ids && end > 0 && (rc = r->ids.reserve((size_t)end * sizeof(int64_t))) != LSHX_OK) return rc;
  const bool host_vectors = (on_device != 2) && total > 0;
  const int64_t vec_rows = cand_ids ? n_vectors : end;
  if (host_vectors && (rc = r->vecs.reserve((size_t)vec_rows * dim * sizeof(float))) != LSHX_OK) return rc;
  if (select) {
    if ((rc = r->pos.reserve((size_t)nq * out_stride * sizeof(int32_t))) != LSHX_OK) return rc;
    if ((rc = r->score.reserve((size_t)nq * out_stride * sizeof(float))) != LSHX_OK) return rc;
    if ((rc = r->count.reserve((size_t)nq * sizeof(int32_t))) != LSHX_OK) return rc;
  } else {
    if ((rc = r->all.reserve((size_t)(end > 0 ? end : 1) * sizeof(float))) != LSHX_OK) return rc;
  }
  if ((rc = r->zero.reserve((size_t)nq * sizeof(int32_t))) != LSHX_OK) return rc;
  const size_t big_need = rerank_scratch_bytes(a, nullptr);
  if (big_need) {
    if ((rc = r->big.reserve(big_need)) != LSHX_OK) return rc;
    a.big_keys = static_cast<uint64_t*>(r->big.p);
    a.big_bytes = r->big.cap;
  }

  float* d_q = static_cast<float*>(r->q.p);
  int64_t* d_offs = static_cast<int64_t*>(r->offs.p);
  int64_t* d_ids = cand_ids ? static_cast<int64_t*>(r->ids.p) : nullptr;
  float* d_vecs = static_cast<float*>(r->vecs.p);
  a.ids = (cand_ids && end > 0) ? d_ids : nullptr;
  a.V = host_vectors ? d_vecs : vectors;
  if (host_vectors) a.n_vectors = vec_rows;

  // everything every chunk needs first: the offsets, and a host corpus that ids gather from
  cudaStream_t streams[2] = {r->stream, r->stream2};
  LSHX_CUDA(cudaMemcpyAsync(d_offs, cand_offsets, (size_t)(nq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                            streams[0]));
  if (host_vectors && cand_ids)
    if ((rc = upload(d_vecs, vectors, (size_t)vec_rows * dim * sizeof(float), streams[0])) != LSHX_OK) return rc;
  LSHX_CUDA(cudaEventRecord(r->ev, streams[0]));
  LSHX_CUDA(cudaStreamWaitEvent(streams[1], r->ev, 0));

  // query chunks alternate between two streams so the H2D of one overlaps the kernel of the other
  // (an oversized selection shares one sort scratch: a single chunk on one stream)
  const int64_t qchunk = (nq >= 1024 && !big_need) ? (nq + 7) / 8 : nq;
  int slot = 0;
  for (int64_t c0 = 0; c0 < nq; c0 += qchunk, slot ^= 1) {
    const int64_t c1 = (c0 + qchunk < nq) ? c0 + qchunk : nq;
    const int64_t s0 = cand_offsets[c0], s1 = cand_offsets[c1];  // candidate slots of this chunk
    cudaStream_t st = streams[slot];
    if ((rc = upload(d_q + c0 * dim, Q + c0 * dim, (size_t)(c1 - c0) * dim * sizeof(float), st)) != LSHX_OK) return rc;
    if (a.ids && s1 > s0)
      if ((rc = upload(d_ids + s0, cand_ids + s0, (size_t)(s1 - s0) * sizeof(int64_t), st)) != LSHX_OK) return rc;
    if (host_vectors && !cand_ids && s1 > s0)  // packed candidates: the rows of this chunk's queries
      if ((rc = upload(d_vecs + s0 * dim, vectors + s0 * dim, (size_t)(s1 - s0) * dim * sizeof(float), st)) != LSHX_OK)
        return rc;
    RerankArgs c = a;
    c.nq = c1 - c0;
    c.Q = d_q + c0 * dim;
    c.offs = d_offs + c0;
    c.out_zero = static_cast<int32_t*>(r->zero.p) + c0;
    if (select) {
      c.out_pos = static_cast<int32_t*>(r->pos.p) + c0 * out_stride;
      c.out_score = static_cast<float*>(r->score.p) + c0 * out_stride;
      c.out_count = static_cast<int32_t*>(r->count.p) + c0;
    } else {
      c.all_scores = static_cast<float*>(r->all.p);
    }
    rc = launch_rerank(c, st);
    if (rc != LSHX_OK) return rc;
    if (select) {
      LSHX_CUDA(cudaMemcpyAsync(out_pos + c0 * out_stride, c.out_pos, (size_t)(c1 - c0) * out_stride * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
      LSHX_CUDA(cudaMemcpyAsync(out_score + c0 * out_stride, c.out_score,
                                (size_t)(c1 - c0) * out_stride * sizeof(float), cudaMemcpyDeviceToHost, st));
      LSHX_CUDA(cudaMemcpyAsync(out_count + c0, c.out_count, (size_t)(c1 - c0) * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
    } else if (s1 > s0) {
      LSHX_CUDA(cudaMemcpyAsync(out_all + s0, c.all_scores + s0, (size_t)(s1 - s0) * sizeof(float),
                                cudaMemcpyDeviceToHost, st));
    }
    if (out_zero)
      LSHX_CUDA(cudaMemcpyAsync(out_zero + c0, c.out_zero, (size_t)(c1 - c0) * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
  }
  LSHX_CUDA(cudaStreamSynchronize(streams[0]));
  LSHX_CUDA(cudaStreamSynchronize(streams[1]));
  return LSHX_OK;
}

extern "C" int lshx_rerank_topk(lshx_reranker* r, const float* Q, int64_t nq, const float* vectors,
                                int64_t n_vectors, const int64_t* cand_offsets,
                                const int64_t* cand_ids, int64_t max_candidates, int k, double p,
                                int out_stride, int32_t* out_pos, float* out_score,
                                int32_t* out_count, int32_t* out_zero, int on_device, void* stream) {
  return rerank_common(r, Q, nq, vectors, n_vectors, cand_offsets, cand_ids, max_candidates, 0,
                       true, k, p, out_stride, out_pos, out_score, out_count, nullptr, out_zero,
                       on_device, stream);
}

extern "C" int lshx_rerank_scores(lshx_reranker* r, const float* Q, int64_t nq, const float* vectors,
                                  int64_t n_vectors, const int64_t* cand_offsets,
                                  const int64_t* cand_ids, int64_t total_candidates,
                                  float* out_scores, int32_t* out_zero, int on_device, void* stream) {
  return rerank_common(r, Q, nq, vectors, n_vectors, cand_offsets, cand_ids, 0, total_candidates,
                       false, 0, 0.0, 0, nullptr, nullptr, nullptr, out_scores, out_zero, on_device,
                       stream);
}

extern "C" int lshx_l2_normalize(lshx_reranker* r, const float* X, int64_t n, float* out,
                                 int32_t* zero_rows, int on_device, void* stream) {
  LSHX_REQUIRE(r != nullptr, "null handle");
  LSHX_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return LSHX_OK;
  LSHX_REQUIRE(X != nullptr && out != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(r->mu);
  DeviceGuard g(r->device);
  cudaStream_t user = reinterpret_cast<cudaStream_t>(stream);
  if (on_device) return launch_l2_normalize(X, n, r->dim, out, zero_rows, user);
  const size_t bytes = (size_t)n * r->dim * sizeof(float);
  int rc;
  if ((rc = r->q.reserve(bytes)) != LSHX_OK) return rc;
  if ((rc = r->vecs.reserve(bytes)) != LSHX_OK) return rc;
  if ((rc = r->zero.reserve((size_t)n * sizeof(int32_t))) != LSHX_OK) return rc;
  cudaStream_t st = r->stream;
  LSHX_CUDA(cudaMemcpyAsync(r->q.p, X, bytes, cudaMemcpyHostToDevice, st));
  rc = launch_l2_normalize(static_cast<const float*>(r->q.p), n, r->dim, static_cast<float*>(r->vecs.p),
                           static_cast<int32_t*>(r->zero.p), st);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(out, r->vecs.p, bytes, cudaMemcpyDeviceToHost, st));
  if (zero_rows)
    LSHX_CUDA(cudaMemcpyAsync(zero_rows, r->zero.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  LSHX_CUDA(cudaStreamSynchronize(st));
  return LSHX_OK;
}


// ---------------------------------------------------------------------------------------
// device band index (index_join.cu)
// ---------------------------------------------------------------------------------------

struct lshx_index {
  int device = 0;
  int nb = 0, bpb = 0;
  int64_t n = 0;         // entries per band (tombstones included)
  // Two sorted runs per band segment: the main run [0, main_n) and the delta run [main_n, delta_n); entries
  // [delta_n, n) were appended since the last query.  A small add is sorted into the delta run on its own; the
  // segment is re-sorted as a whole only when the delta outgrows a quarter of the main run.
  int64_t main_n = 0, delta_n = 0;
  int64_t cap = 0;       // entries each band segment can hold
  uint64_t* keys[2] = {nullptr, nullptr};
  int64_t* ids[2] = {nullptr, nullptr};
  int cur = 0;
  unsigned long long* d_max_id = nullptr;
  int* d_bad = nullptr;
  cudaStream_t stream = nullptr;
  DevBuf hist, stage_sig, stage_ids, gone;
  // state of the last query (device buffers, valid until the next query on this handle)
  DevBuf q_sig, lo, cnt, raw_count, raw_off, ws_off, meta, ws, out_ids, out_coll, uniq, topk_ids, topk_cnt;
  DevBuf rr_pos, rr_score, rr_count, rr_zero, rr_ids, rr_q, q_flag;
  int64_t last_nq = -1, last_total = 0, last_max = 0;
  int64_t last_q_rows = -1;   // >= 0: rr_q holds the query VECTORS of the last result (lshx_index_query_host_vectors)
  int last_q_dim = 0;
  // latency path (lshx_index_query_vectors): pinned + mapped result block the kernel stores into, and the
  // ticket counter by which the fused hash + query kernel finds its last CTA
  uint8_t* pin_res = nullptr;
  unsigned* d_ticket = nullptr;
  unsigned long long* d_dbg = nullptr;   // diagnostics: phase stamps of the latency kernel (lshx_index_debug_timeline)
  std::mutex mu;
};

constexpr int IDX_SMALL_MAX_Q = 32;
constexpr int IDX_SMALL_MAX_CAP = 4096;

static int index_reserve(lshx_index* ix, int64_t want) {
  if (want <= ix->cap) return LSHX_OK;
  int64_t cap = ix->cap ? ix->cap : 4096;
  while (cap < want) cap += cap / 2 + 4096;
  uint64_t* nk[2] = {nullptr, nullptr};
  int64_t* ni[2] = {nullptr, nullptr};
  const size_t bytes = (size_t)ix->nb * cap * 8;
  for (int i = 0; i < 2; ++i) {
    if (cudaMalloc(reinterpret_cast<void**>(&nk[i]), bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&ni[i]), bytes) != cudaSuccess) {
      (void)cudaGetLastError();
      for (int j = 0; j < 2; ++j) {
        if (nk[j]) cudaFree(nk[j]);
        if (ni[j]) cudaFree(ni[j]);
      }
      set_error("cudaMalloc of %zu bytes for the band index failed", bytes);
      return LSHX_ERR_OOM;
    }
  }
  if (ix->n > 0) {   // only the current side carries data
    LSHX_CUDA(cudaMemcpy2DAsync(nk[0], (size_t)cap * 8, ix->keys[ix->cur], (size_t)ix->cap * 8, (size_t)ix->n * 8,
                                (size_t)ix->nb, cudaMemcpyDeviceToDevice, ix->stream));
    LSHX_CUDA(cudaMemcpy2DAsync(ni[0], (size_t)cap * 8, ix->ids[ix->cur], (size_t)ix->cap * 8, (size_t)ix->n * 8,
                                (size_t)ix->nb, cudaMemcpyDeviceToDevice, ix->stream));
    LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  }
  for (int i = 0; i < 2; ++i) {
    if (ix->keys[i]) cudaFree(ix->keys[i]);
    if (ix->ids[i]) cudaFree(ix->ids[i]);
    ix->keys[i] = nk[i];
    ix->ids[i] = ni[i];
  }
  ix->cur = 0;
  ix->cap = cap;
  return LSHX_OK;
}

extern "C" int lshx_index_create(int device, int num_bands, int bytes_per_band, lshx_index** out) {
  LSHX_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  LSHX_REQUIRE(num_bands > 0 && num_bands <= 255, "num_bands must be in [1, 255] for the device index");
  LSHX_REQUIRE(bytes_per_band > 0 && bytes_per_band <= 8,
               "the device index takes band keys of at most 8 bytes (rows_per_band <= 64)");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);
  lshx_index* ix = new lshx_index();
  ix->device = device;
  ix->nb = num_bands;
  ix->bpb = bytes_per_band;
  if (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ix->d_max_id), 8) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ix->d_bad), 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ix->d_ticket), 4) != cudaSuccess ||
      cudaMemset(ix->d_max_id, 0, 8) != cudaSuccess || cudaMemset(ix->d_bad, 0, 4) != cudaSuccess ||
      cudaMemset(ix->d_ticket, 0, 4) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cannot create the band index on device %d", device);
    lshx_index_destroy(ix);
    return LSHX_ERR_CUDA;
  }
  *out = ix;
  return LSHX_OK;
}

extern "C" int lshx_index_destroy(lshx_index* ix) {
  if (!ix) return LSHX_OK;
  {
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
      if (ix->keys[i]) cudaFree(ix->keys[i]);
      if (ix->ids[i]) cudaFree(ix->ids[i]);
    }
    if (ix->d_max_id) cudaFree(ix->d_max_id);
    if (ix->d_bad) cudaFree(ix->d_bad);
    if (ix->d_ticket) cudaFree(ix->d_ticket);
    if (ix->d_dbg) cudaFree(ix->d_dbg);
    if (ix->pin_res) cudaFreeHost(ix->pin_res);
    for (DevBuf* b : {&ix->hist, &ix->stage_sig, &ix->stage_ids, &ix->gone, &ix->q_sig, &ix->lo, &ix->cnt,
                      &ix->raw_count, &ix->raw_off, &ix->ws_off, &ix->meta, &ix->ws, &ix->out_ids, &ix->out_coll,
                      &ix->uniq, &ix->topk_ids, &ix->topk_cnt, &ix->rr_pos, &ix->rr_score, &ix->rr_count,
                      &ix->rr_zero, &ix->rr_ids, &ix->rr_q, &ix->q_flag})
      b->release();
    if (ix->stream) cudaStreamDestroy(ix->stream);
    (void)cudaGetLastError();
  }
  delete ix;
  return LSHX_OK;
}

extern "C" int64_t lshx_index_size(const lshx_index* ix) { return ix ? ix->n : 0; }

static int index_add_common(lshx_index* ix, const uint8_t* signatures, const int64_t* ids, int64_t n, int on_device,
                            void* stream, int per_band) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  LSHX_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return LSHX_OK;
  LSHX_REQUIRE(signatures != nullptr && ids != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  int rc = index_reserve(ix, ix->n + n);
  if (rc != LSHX_OK) return rc;
  const size_t sig_bytes = (size_t)n * ix->nb * ix->bpb;
  const size_t id_bytes = (size_t)n * 8 * (per_band ? ix->nb : 1);
  const uint8_t* d_sig = signatures;
  const int64_t* d_ids = ids;
  if (!on_device) {
    if ((rc = ix->stage_sig.reserve(sig_bytes)) != LSHX_OK) return rc;
    if ((rc = ix->stage_ids.reserve(id_bytes)) != LSHX_OK) return rc;
    if ((rc = upload(ix->stage_sig.p, signatures, sig_bytes, ix->stream)) != LSHX_OK) return rc;
    if ((rc = upload(ix->stage_ids.p, ids, id_bytes, ix->stream)) != LSHX_OK) return rc;
    d_sig = static_cast<const uint8_t*>(ix->stage_sig.p);
    d_ids = static_cast<const int64_t*>(ix->stage_ids.p);
  } else {
    // the signatures were produced on the caller's stream: order ours after it
    LSHX_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  }
  rc = index_append(d_sig, d_ids, n, ix->nb, ix->bpb, ix->keys[ix->cur], ix->ids[ix->cur], ix->cap, ix->n,
                    ix->d_max_id, ix->d_bad, per_band, ix->stream);
  if (rc != LSHX_OK) return rc;
  int bad = 0;
  LSHX_CUDA(cudaMemcpyAsync(&bad, ix->d_bad, 4, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  if (bad) {
    LSHX_CUDA(cudaMemset(ix->d_bad, 0, 4));
    set_error("vector ids must lie in [0, 2^56) for the device index");
    return LSHX_ERR_INVALID_ARG;   // the entries were not counted in: n is unchanged
  }
  ix->n += n;
  ix->last_nq = -1;
  return LSHX_OK;
}

extern "C" int lshx_index_add(lshx_index* ix, const uint8_t* signatures, const int64_t* ids, int64_t n,
                              int on_device, void* stream) {
  return index_add_common(ix, signatures, ids, n, on_device, stream, 0);
}

extern "C" int lshx_index_add_entries(lshx_index* ix, const uint8_t* signatures, const int64_t* ids_per_band, int64_t n) {
  return index_add_common(ix, signatures, ids_per_band, n, 0, nullptr, 1);
}

extern "C" int lshx_index_clear(lshx_index* ix) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  ix->n = ix->main_n = ix->delta_n = 0;
  ix->last_nq = -1;
  LSHX_CUDA(cudaMemset(ix->d_max_id, 0, 8));
  return LSHX_OK;
}

extern "C" int lshx_index_remove(lshx_index* ix, const int64_t* ids_host, int64_t n) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  LSHX_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0 || ix->n == 0) return LSHX_OK;
  LSHX_REQUIRE(ids_host != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  std::vector<int64_t> sorted(ids_host, ids_host + n);
  std::sort(sorted.begin(), sorted.end());
  int rc = ix->gone.reserve((size_t)n * 8);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(ix->gone.p, sorted.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ix->stream));
  rc = index_tombstone(ix->ids[ix->cur], ix->n, ix->cap, ix->nb, static_cast<const int64_t*>(ix->gone.p), n, ix->stream);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  ix->last_nq = -1;
  return LSHX_OK;
}

// sort what add() appended since the last query: into the delta run when that is small beside the main run,
// else the whole segment (which folds the delta into the main run).  force_full: a caller that reads ranges of
// ONE run (get_buckets, export).
constexpr int64_t IDX_MIN_MAIN = 32768;   // below this a full sort is as cheap as bookkeeping

static int index_make_sorted(lshx_index* ix, bool force_full = false) {
  if (ix->main_n == ix->n) return LSHX_OK;
  if (!force_full && ix->delta_n == ix->n) return LSHX_OK;
  unsigned long long max_id = 0;
  LSHX_CUDA(cudaMemcpyAsync(&max_id, ix->d_max_id, 8, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  // bytes of max_id + 1: the all-ones pattern of a tombstone's low bytes then never equals a live id's, so
  // tombstones sort after every live id of their bucket
  int id_bytes = 1;
  while (id_bytes < 8 && ((max_id + 1) >> (8 * id_bytes)) != 0) ++id_bytes;
  const int64_t tail = ix->n - ix->main_n;
  const bool full = force_full || ix->main_n < IDX_MIN_MAIN || tail * 4 > ix->main_n;
  const int64_t first = full ? 0 : ix->main_n, count = full ? ix->n : tail;
  const size_t hist_entries = index_sort_hist_entries(count, ix->nb);
  int rc = ix->hist.reserve(hist_entries * sizeof(unsigned));
  if (rc != LSHX_OK) return rc;
  int cur = ix->cur;
  rc = index_sort(ix->keys, ix->ids, &cur, first, count, ix->cap, ix->nb, ix->bpb, id_bytes,
                  static_cast<unsigned*>(ix->hist.p), hist_entries, ix->stream);
  if (rc != LSHX_OK) return rc;
  if (full) {
    ix->cur = cur;
    ix->main_n = ix->delta_n = ix->n;
    return LSHX_OK;
  }
  if (cur != ix->cur) {   // an odd number of passes left the sorted delta in the other buffer: bring it home
    const size_t pitch = (size_t)ix->cap * 8, width = (size_t)count * 8;
    LSHX_CUDA(cudaMemcpy2DAsync(ix->keys[ix->cur] + first, pitch, ix->keys[cur] + first, pitch, width, (size_t)ix->nb,
                                cudaMemcpyDeviceToDevice, ix->stream));
    LSHX_CUDA(cudaMemcpy2DAsync(ix->ids[ix->cur] + first, pitch, ix->ids[cur] + first, pitch, width, (size_t)ix->nb,
                                cudaMemcpyDeviceToDevice, ix->stream));
  }
  ix->delta_n = ix->n;
  return LSHX_OK;
}

// lookup + scan + join of nq signatures that are in device memory and ordered on ix->stream (caller holds ix->mu)
static int index_query_core(lshx_index* ix, const uint8_t* d_sig, int64_t nq, int64_t* total_candidates,
                            int64_t* max_candidates) {
  int rc;
  if ((rc = ix->lo.reserve((size_t)nq * ix->nb * 2 * 8)) != LSHX_OK) return rc;    // two runs per band
  if ((rc = ix->cnt.reserve((size_t)nq * ix->nb * 2 * 4)) != LSHX_OK) return rc;
  if ((rc = ix->raw_count.reserve((size_t)nq * 4)) != LSHX_OK) return rc;
  if ((rc = ix->raw_off.reserve((size_t)(nq + 1) * 8)) != LSHX_OK) return rc;
  if ((rc = ix->ws_off.reserve((size_t)(nq + 1) * 8)) != LSHX_OK) return rc;
  if ((rc = ix->meta.reserve(32)) != LSHX_OK) return rc;
  if ((rc = ix->uniq.reserve((size_t)nq * 4)) != LSHX_OK) return rc;
  const int nruns = ix->n > ix->main_n ? 2 : 1;
  rc = index_lookup_scan(d_sig, nq, ix->nb, ix->bpb, ix->keys[ix->cur], ix->main_n, ix->n, ix->cap,
                         static_cast<int64_t*>(ix->lo.p), static_cast<int*>(ix->cnt.p),
                         static_cast<int*>(ix->raw_count.p), static_cast<int64_t*>(ix->raw_off.p),
                         static_cast<int64_t*>(ix->ws_off.p), static_cast<int64_t*>(ix->meta.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  int64_t meta[3] = {0, 0, 0};
  LSHX_CUDA(cudaMemcpyAsync(meta, ix->meta.p, sizeof(meta), cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  const int64_t total = meta[0], maxc = meta[1], ws_total = meta[2];
  LSHX_REQUIRE(total < (1ll << 40), "query batch matches %lld bucket entries; split the batch", (long long)total);
  if ((rc = ix->out_ids.reserve((size_t)(total > 0 ? total : 1) * 8)) != LSHX_OK) return rc;
  if ((rc = ix->out_coll.reserve((size_t)(total > 0 ? total : 1) * 4)) != LSHX_OK) return rc;
  uint64_t* d_ws = nullptr;
  if (maxc > (int64_t)index_join_smem_cap()) {   // a query too large for the shared-memory sort: global workspace
    if ((rc = ix->ws.reserve((size_t)ws_total * 2 * 8)) != LSHX_OK) return rc;
    d_ws = static_cast<uint64_t*>(ix->ws.p);
  }
  rc = index_join(nq, ix->nb, nruns, ix->ids[ix->cur], ix->cap, static_cast<const int64_t*>(ix->lo.p),
                  static_cast<const int*>(ix->cnt.p), static_cast<const int*>(ix->raw_count.p),
                  static_cast<const int64_t*>(ix->raw_off.p), static_cast<const int64_t*>(ix->ws_off.p), d_ws,
                  ws_total, static_cast<int64_t*>(ix->out_ids.p), static_cast<int*>(ix->out_coll.p),
                  static_cast<int*>(ix->uniq.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  ix->last_nq = nq;
  ix->last_total = total;
  ix->last_max = maxc;
  if (total_candidates) *total_candidates = total;
  if (max_candidates) *max_candidates = maxc;
  return LSHX_OK;
}

extern "C" int lshx_index_query(lshx_index* ix, const uint8_t* signatures, int64_t nq, int on_device, void* stream,
                                int64_t* total_candidates, int64_t* max_candidates) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  LSHX_REQUIRE(nq >= 0, "nq must be >= 0");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  ix->last_nq = -1;
  ix->last_q_rows = -1;
  if (total_candidates) *total_candidates = 0;
  if (max_candidates) *max_candidates = 0;
  if (nq == 0) {
    ix->last_nq = 0;
    ix->last_total = ix->last_max = 0;
    return LSHX_OK;
  }
  LSHX_REQUIRE(signatures != nullptr, "null buffer");
  int rc = index_make_sorted(ix);
  if (rc != LSHX_OK) return rc;
  const size_t sig_bytes = (size_t)nq * ix->nb * ix->bpb;
  const uint8_t* d_sig = signatures;
  if (!on_device) {
    if ((rc = ix->q_sig.reserve(sig_bytes)) != LSHX_OK) return rc;
    if ((rc = upload(ix->q_sig.p, signatures, sig_bytes, ix->stream)) != LSHX_OK) return rc;
    d_sig = static_cast<const uint8_t*>(ix->q_sig.p);
  } else {
    LSHX_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  }
  return index_query_core(ix, d_sig, nq, total_candidates, max_candidates);
}

// query_batch on HOST vectors in one pass over PCIe: upload the nq vectors once, hash them on the device, join --
// the signatures never visit the host, and the vectors stay in the handle for lshx_index_rerank (Q = NULL).
extern "C" int lshx_index_query_host_vectors(lshx_index* ix, lshx_hasher* h, const float* X, int64_t nq,
                                             uint8_t* zero_flag, int64_t* total_candidates,
                                             int64_t* max_candidates) {
  LSHX_REQUIRE(ix != nullptr && h != nullptr, "null handle");
  LSHX_REQUIRE(ix->device == h->device, "index and hasher live on different devices");
  LSHX_REQUIRE(nq >= 0, "nq must be >= 0");
  std::lock_guard<std::mutex> lk(ix->mu);
  std::lock_guard<std::mutex> lk2(h->mu);
  DeviceGuard g(ix->device);
  const HashShape& s = h->s;
  LSHX_REQUIRE(s.num_bands == ix->nb && s.sig_bytes == ix->nb * ix->bpb, "hasher and index shapes differ");
  ix->last_nq = -1;
  ix->last_q_rows = -1;
  if (total_candidates) *total_candidates = 0;
  if (max_candidates) *max_candidates = 0;
  if (nq == 0) {
    ix->last_nq = 0;
    ix->last_total = ix->last_max = 0;
    return LSHX_OK;
  }
  LSHX_REQUIRE(X != nullptr, "null buffer");
  int rc = index_make_sorted(ix);
  if (rc != LSHX_OK) return rc;
  if ((rc = ix->rr_q.reserve((size_t)nq * s.dim * sizeof(float))) != LSHX_OK) return rc;
  if ((rc = ix->q_sig.reserve((size_t)nq * s.sig_bytes)) != LSHX_OK) return rc;
  if ((rc = ix->q_flag.reserve((size_t)nq)) != LSHX_OK) return rc;
  if ((rc = upload(ix->rr_q.p, X, (size_t)nq * s.dim * sizeof(float), ix->stream)) != LSHX_OK) return rc;
  rc = launch_hash(h, static_cast<const float*>(ix->rr_q.p), nq, static_cast<uint8_t*>(ix->q_sig.p),
                   zero_flag ? static_cast<uint8_t*>(ix->q_flag.p) : nullptr, ix->stream);
  if (rc != LSHX_OK) return rc;
  if (zero_flag)
    LSHX_CUDA(cudaMemcpyAsync(zero_flag, ix->q_flag.p, (size_t)nq, cudaMemcpyDeviceToHost, ix->stream));
  rc = index_query_core(ix, static_cast<const uint8_t*>(ix->q_sig.p), nq, total_candidates, max_candidates);
  if (rc != LSHX_OK) return rc;
  ix->last_q_rows = nq;       // rr_q holds the queries of this result
  ix->last_q_dim = s.dim;
  return LSHX_OK;
}

extern "C" int lshx_index_get_buckets(lshx_index* ix, const int32_t* band_ids, const uint8_t* keys, int64_t m,
                                      int64_t* offsets, int64_t* ids_out, int64_t ids_capacity, int64_t* needed) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  LSHX_REQUIRE(m >= 0, "m must be >= 0");
  LSHX_REQUIRE(offsets != nullptr, "null buffer");
  if (needed) *needed = 0;
  offsets[0] = 0;
  if (m == 0) return LSHX_OK;
  LSHX_REQUIRE(band_ids != nullptr && keys != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  std::vector<uint64_t> want((size_t)m);
  for (int64_t t = 0; t < m; ++t) {
    LSHX_REQUIRE(band_ids[t] >= 0 && band_ids[t] < ix->nb, "band id %d outside [0, %d)", band_ids[t], ix->nb);
    uint64_t k = 0;
    for (int j = 0; j < ix->bpb; ++j) k |= (uint64_t)keys[t * ix->bpb + j] << (8 * j);
    want[(size_t)t] = k;
  }
  if (ix->n == 0) {
    for (int64_t t = 0; t < m; ++t) offsets[t + 1] = 0;
    return LSHX_OK;
  }
  int rc = index_make_sorted(ix, /*force_full=*/true);   // a bucket is read as ONE id-ascending range
  if (rc != LSHX_OK) return rc;
  // scratch in the query buffers (their contents die with last_nq below)
  ix->last_nq = -1;
  if ((rc = ix->q_sig.reserve((size_t)m * 8)) != LSHX_OK) return rc;
  if ((rc = ix->cnt.reserve((size_t)m * 4)) != LSHX_OK) return rc;
  if ((rc = ix->lo.reserve((size_t)m * 8)) != LSHX_OK) return rc;
  if ((rc = ix->raw_off.reserve((size_t)(m + 1) * 8)) != LSHX_OK) return rc;
  if ((rc = ix->ws_off.reserve((size_t)m * 8)) != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(ix->q_sig.p, want.data(), (size_t)m * 8, cudaMemcpyHostToDevice, ix->stream));
  LSHX_CUDA(cudaMemcpyAsync(ix->cnt.p, band_ids, (size_t)m * 4, cudaMemcpyHostToDevice, ix->stream));
  rc = index_bucket_lookup(static_cast<const int*>(ix->cnt.p), static_cast<const uint64_t*>(ix->q_sig.p), m,
                           ix->keys[ix->cur], ix->n, ix->cap, static_cast<int64_t*>(ix->lo.p),
                           static_cast<int64_t*>(ix->ws_off.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  std::vector<int64_t> raw((size_t)m + 1);
  LSHX_CUDA(cudaMemcpyAsync(raw.data() + 1, ix->ws_off.p, (size_t)m * 8, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  raw[0] = 0;
  for (int64_t t = 0; t < m; ++t) raw[(size_t)t + 1] += raw[(size_t)t];
  const int64_t total = raw[(size_t)m];
  if (needed) *needed = total;
  if (ids_out == nullptr || ids_capacity < total) {
    // sizes only: an upper bound per bucket (stored entries, removed ones and repeats included)
    for (int64_t t = 0; t <= m; ++t) offsets[t] = raw[(size_t)t];
    if (ids_out != nullptr) {
      set_error("ids_out holds %lld entries, the buckets may need %lld", (long long)ids_capacity, (long long)total);
      return LSHX_ERR_INVALID_ARG;
    }
    return LSHX_OK;
  }
  if (total == 0) {
    for (int64_t t = 0; t < m; ++t) offsets[t + 1] = 0;
    return LSHX_OK;
  }
  if ((rc = ix->out_ids.reserve((size_t)total * 8)) != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(ix->raw_off.p, raw.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, ix->stream));
  rc = index_bucket_gather(static_cast<const int*>(ix->cnt.p), static_cast<const int64_t*>(ix->lo.p),
                           static_cast<const int64_t*>(ix->raw_off.p), m, ix->ids[ix->cur], ix->cap,
                           static_cast<int64_t*>(ix->out_ids.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  std::vector<int64_t> stored((size_t)total);
  LSHX_CUDA(cudaMemcpyAsync(stored.data(), ix->out_ids.p, (size_t)total * 8, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  // SET members: live ids, each once (they are stored in ascending order, removed ones as -1 at the end)
  int64_t w = 0;
  for (int64_t t = 0; t < m; ++t) {
    const int64_t begin = w;
    for (int64_t e = raw[(size_t)t]; e < raw[(size_t)t + 1]; ++e) {
      const int64_t id = stored[(size_t)e];
      if (id < 0 || (w > begin && ids_out[w - 1] == id)) continue;
      ids_out[w++] = id;
    }
    offsets[t + 1] = w;
  }
  return LSHX_OK;
}

extern "C" int lshx_index_export(lshx_index* ix, uint8_t* keys_out, int64_t* ids_out, int64_t capacity, int64_t* n_out) {
  LSHX_REQUIRE(ix != nullptr && n_out != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  *n_out = ix->n;
  if (keys_out == nullptr && ids_out == nullptr) return LSHX_OK;   // size only
  LSHX_REQUIRE(keys_out != nullptr && ids_out != nullptr, "null buffer");
  LSHX_REQUIRE(capacity >= ix->n, "buffers hold %lld entries per band, the index has %lld", (long long)capacity,
               (long long)ix->n);
  if (ix->n == 0) return LSHX_OK;
  int rc = index_make_sorted(ix, /*force_full=*/true);
  if (rc != LSHX_OK) return rc;
  const int64_t n = ix->n;
  std::vector<uint64_t> k((size_t)ix->nb * n);
  LSHX_CUDA(cudaMemcpy2DAsync(k.data(), (size_t)n * 8, ix->keys[ix->cur], (size_t)ix->cap * 8, (size_t)n * 8,
                              (size_t)ix->nb, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaMemcpy2DAsync(ids_out, (size_t)n * 8, ix->ids[ix->cur], (size_t)ix->cap * 8, (size_t)n * 8,
                              (size_t)ix->nb, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  for (size_t e = 0; e < k.size(); ++e)
    for (int j = 0; j < ix->bpb; ++j) keys_out[e * ix->bpb + j] = (uint8_t)(k[e] >> (8 * j));
  return LSHX_OK;
}

extern "C" int lshx_index_query_vectors(lshx_index* ix, lshx_hasher* h, const float* X, int nq, int capacity,
                                        int64_t* out_ids, int32_t* out_collisions, int32_t* out_count,
                                        uint8_t* zero_flag) {
  LSHX_REQUIRE(ix != nullptr && h != nullptr, "null handle");
  LSHX_REQUIRE(ix->device == h->device, "index and hasher live on different devices");
  LSHX_REQUIRE(nq >= 0 && nq <= IDX_SMALL_MAX_Q, "lshx_index_query_vectors takes at most %d vectors", IDX_SMALL_MAX_Q);
  LSHX_REQUIRE(capacity > 0 && capacity <= IDX_SMALL_MAX_CAP, "capacity must be in [1, %d]", IDX_SMALL_MAX_CAP);
  if (nq == 0) return LSHX_OK;
  LSHX_REQUIRE(X != nullptr && out_ids != nullptr && out_collisions != nullptr && out_count != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  std::lock_guard<std::mutex> lk2(h->mu);
  DeviceGuard g(ix->device);
  const HashShape& s = h->s;
  LSHX_REQUIRE(s.num_bands == ix->nb && s.sig_bytes == ix->nb * ix->bpb, "hasher and index shapes differ");
  LSHX_REQUIRE(nq <= h->small_rows, "the hasher's latency path takes at most %d rows of this dimension", h->small_rows);
  LSHX_REQUIRE(h->kernel_pref == LSHX_KERNEL_AUTO, "a hasher pinned to one kernel does not take the latency path");
  if (!ix->pin_res) {
    const size_t bytes = (size_t)IDX_SMALL_MAX_Q * IDX_SMALL_MAX_CAP * 12 + IDX_SMALL_MAX_Q * 8;
    if (cudaHostAlloc(reinterpret_cast<void**>(&ix->pin_res), bytes, cudaHostAllocMapped) != cudaSuccess) {
      (void)cudaGetLastError();
      ix->pin_res = nullptr;
      set_error("cannot allocate the pinned result block of the latency path");
      return LSHX_ERR_OOM;
    }
  }
  int rc = index_make_sorted(ix);
  if (rc != LSHX_OK) return rc;
  ix->last_nq = -1;
  cudaStream_t st = ix->stream;
  // result block: ids [nq][capacity] | collisions [nq][capacity] | count [nq] | zero flag [nq]
  uint8_t* res = ix->pin_res;
  int64_t* h_ids = reinterpret_cast<int64_t*>(res);
  int32_t* h_coll = reinterpret_cast<int32_t*>(res + (size_t)nq * capacity * 8);
  int32_t* h_count = h_coll + (size_t)nq * capacity;
  uint8_t* h_flag = reinterpret_cast<uint8_t*>(h_count + nq);
  uint8_t* d_res = nullptr;
  LSHX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_res), res, 0));
  const size_t off_coll = (size_t)nq * capacity * 8, off_count = off_coll + (size_t)nq * capacity * 4,
               off_flag = off_count + (size_t)nq * 4;
  if (ix->n == 0) {   // nothing indexed: only the zero-vector test has work to do
    for (int q = 0; q < nq; ++q) {
      out_count[q] = 0;
      if (zero_flag) {
        bool viol = false;
        for (int k = 0; k < s.dim; ++k) viol |= !(std::fabs(X[(size_t)q * s.dim + k]) <= 1e-8f);
        zero_flag[q] = viol ? 0 : 1;
      }
    }
    return LSHX_OK;
  }
  if ((rc = ix->q_sig.reserve((size_t)IDX_SMALL_MAX_Q * s.sig_bytes)) != LSHX_OK) return rc;
  const size_t xb = (size_t)nq * s.dim * sizeof(float);
  std::memcpy(h->pin_x, X, xb);
  float* d_x_map = nullptr;                  // the kernel reads the vectors from the pinned block in place
  LSHX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_x_map), h->pin_x, 0));
  h->last_kernel = LSHX_KERNEL_SMALL;
  rc = index_hash_query_small(d_x_map, h->pin_x, nq, s.dim, h->d_Rp, static_cast<uint8_t*>(ix->q_sig.p), s.sig_bytes,
                              zero_flag ? d_res + off_flag : nullptr, ix->d_ticket, ix->nb, ix->bpb,
                              ix->keys[ix->cur], ix->ids[ix->cur], ix->main_n, ix->n, ix->cap, capacity, 0,
                              reinterpret_cast<int64_t*>(d_res), reinterpret_cast<int*>(d_res + off_coll),
                              reinterpret_cast<int*>(d_res + off_count), nullptr, nullptr, ix->d_dbg, st);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaStreamSynchronize(st));
  for (int q = 0; q < nq; ++q) {
    const int c = h_count[q];
    out_count[q] = c;
    const int take = c < capacity ? (c > 0 ? c : 0) : capacity;
    std::memcpy(out_ids + (size_t)q * capacity, h_ids + (size_t)q * capacity, (size_t)take * 8);
    std::memcpy(out_collisions + (size_t)q * capacity, h_coll + (size_t)q * capacity, (size_t)take * 4);
  }
  if (zero_flag) std::memcpy(zero_flag, h_flag, (size_t)nq);
  return LSHX_OK;
}

constexpr int IDX_SMALL_RERANK_CAP = 1024;   // candidates per query the fused latency path reranks

extern "C" int lshx_index_query_rerank_vectors(lshx_index* ix, lshx_hasher* h, lshx_reranker* r, const float* X,
                                               int nq, const float* corpus_device, int64_t n_vectors, int k,
                                               double p, int out_stride, int64_t* out_ids, float* out_score,
                                               int32_t* out_count, int32_t* out_zero, int32_t* out_candidates,
                                               uint8_t* zero_flag) {
  LSHX_REQUIRE(ix != nullptr && h != nullptr && r != nullptr, "null handle");
  LSHX_REQUIRE(ix->device == h->device && ix->device == r->device, "index, hasher and reranker live on different devices");
  LSHX_REQUIRE(nq >= 0 && nq <= IDX_SMALL_MAX_Q, "the latency path takes at most %d vectors", IDX_SMALL_MAX_Q);
  LSHX_REQUIRE(k > 0 || p > 0.0, "k must be > 0");
  LSHX_REQUIRE(!(p > 1.0), "top_p must be within the range (0, 1]");
  LSHX_REQUIRE(out_stride > 0 && out_stride <= IDX_SMALL_RERANK_CAP, "out_stride must be in [1, %d]", IDX_SMALL_RERANK_CAP);
  if (nq == 0) return LSHX_OK;
  LSHX_REQUIRE(X && corpus_device && n_vectors > 0 && out_ids && out_score && out_count && out_candidates, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  std::lock_guard<std::mutex> lk2(h->mu);
  std::lock_guard<std::mutex> lk3(r->mu);
  DeviceGuard g(ix->device);
  const HashShape& s = h->s;
  LSHX_REQUIRE(s.num_bands == ix->nb && s.sig_bytes == ix->nb * ix->bpb, "hasher and index shapes differ");
  LSHX_REQUIRE(s.dim == r->dim, "hasher and reranker dimensions differ");
  LSHX_REQUIRE(nq <= h->small_rows, "the hasher's latency path takes at most %d rows of this dimension", h->small_rows);
  LSHX_REQUIRE(h->kernel_pref == LSHX_KERNEL_AUTO, "a hasher pinned to one kernel does not take the latency path");
  if (!ix->pin_res) {
    const size_t bytes = (size_t)IDX_SMALL_MAX_Q * IDX_SMALL_MAX_CAP * 12 + IDX_SMALL_MAX_Q * 8;
    if (cudaHostAlloc(reinterpret_cast<void**>(&ix->pin_res), bytes, cudaHostAllocMapped) != cudaSuccess) {
      (void)cudaGetLastError();
      ix->pin_res = nullptr;
      set_error("cannot allocate the pinned result block of the latency path");
      return LSHX_ERR_OOM;
    }
  }
  if (ix->n == 0) {
    for (int q = 0; q < nq; ++q) {
      out_count[q] = 0;
      out_candidates[q] = 0;
      if (out_zero) out_zero[q] = 0;
      if (zero_flag) {
        bool viol = false;
        for (int c = 0; c < s.dim; ++c) viol |= !(std::fabs(X[(size_t)q * s.dim + c]) <= 1e-8f);
        zero_flag[q] = viol ? 0 : 1;
      }
    }
    return LSHX_OK;
  }
  int rc = index_make_sorted(ix);
  if (rc != LSHX_OK) return rc;
  ix->last_nq = -1;
  cudaStream_t st = ix->stream;
  const int RC = IDX_SMALL_RERANK_CAP;
  // mapped result block: ids [nq][stride] | score [nq][stride] | count [nq] | zero-norm [nq] | candidates [nq] | flag [nq]
  uint8_t* res = ix->pin_res;
  uint8_t* d_res = nullptr;
  LSHX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_res), res, 0));
  const size_t off_score = (size_t)nq * out_stride * 8, off_count = off_score + (size_t)nq * out_stride * 4,
               off_rzero = off_count + (size_t)nq * 4, off_cand = off_rzero + (size_t)nq * 4,
               off_flag = off_cand + (size_t)nq * 4;
  if ((rc = ix->q_sig.reserve((size_t)IDX_SMALL_MAX_Q * s.sig_bytes)) != LSHX_OK) return rc;
  if ((rc = ix->out_ids.reserve((size_t)IDX_SMALL_MAX_Q * RC * 8)) != LSHX_OK) return rc;
  if ((rc = ix->uniq.reserve((size_t)IDX_SMALL_MAX_Q * 4)) != LSHX_OK) return rc;
  if ((rc = ix->raw_off.reserve((size_t)(IDX_SMALL_MAX_Q + 1) * 8)) != LSHX_OK) return rc;
  if ((rc = ix->rr_pos.reserve((size_t)IDX_SMALL_MAX_Q * RC * 4)) != LSHX_OK) return rc;
  const size_t xb = (size_t)nq * s.dim * sizeof(float);
  std::memcpy(h->pin_x, X, xb);
  float* d_x_map = nullptr;
  LSHX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_x_map), h->pin_x, 0));
  h->last_kernel = LSHX_KERNEL_SMALL;
  rc = index_hash_query_small(d_x_map, h->pin_x, nq, s.dim, h->d_Rp, static_cast<uint8_t*>(ix->q_sig.p), s.sig_bytes,
                              zero_flag ? d_res + off_flag : nullptr, ix->d_ticket, ix->nb, ix->bpb,
                              ix->keys[ix->cur], ix->ids[ix->cur], ix->main_n, ix->n, ix->cap, RC, RC,
                              static_cast<int64_t*>(ix->out_ids.p), nullptr, reinterpret_cast<int*>(d_res + off_cand),
                              static_cast<int*>(ix->uniq.p), static_cast<int64_t*>(ix->raw_off.p), nullptr, st);
  if (rc != LSHX_OK) return rc;
  RerankArgs a{};
  a.Q = d_x_map;                            // the rerank kernel stages each query in shared memory once
  a.nq = nq;
  a.V = corpus_device;
  a.n_vectors = n_vectors;
  a.offs = static_cast<const int64_t*>(ix->raw_off.p);
  a.ids = static_cast<const int64_t*>(ix->out_ids.p);
  a.cand_counts = static_cast<const int32_t*>(ix->uniq.p);
  a.dim = s.dim;
  a.k = k;
  a.p = p;
  a.out_stride = out_stride;
  a.out_pos = static_cast<int32_t*>(ix->rr_pos.p);
  a.out_score = reinterpret_cast<float*>(d_res + off_score);
  a.out_count = reinterpret_cast<int32_t*>(d_res + off_count);
  a.out_zero = reinterpret_cast<int32_t*>(d_res + off_rzero);
  a.max_cand = RC;
  a.select = true;
  if ((rc = launch_rerank(a, st)) != LSHX_OK) return rc;
  rc = index_pos_to_id(static_cast<const int64_t*>(ix->out_ids.p), static_cast<const int64_t*>(ix->raw_off.p), a.out_pos,
                       a.out_count, nq, out_stride, reinterpret_cast<int64_t*>(d_res), st);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaStreamSynchronize(st));
  const int32_t* h_count = reinterpret_cast<const int32_t*>(res + off_count);
  for (int q = 0; q < nq; ++q) {
    const int c = h_count[q];
    out_count[q] = c;
    std::memcpy(out_ids + (size_t)q * out_stride, res + (size_t)q * out_stride * 8, (size_t)(c > 0 ? c : 0) * 8);
    std::memcpy(out_score + (size_t)q * out_stride, res + off_score + (size_t)q * out_stride * 4, (size_t)(c > 0 ? c : 0) * 4);
  }
  std::memcpy(out_candidates, res + off_cand, (size_t)nq * 4);
  if (out_zero) std::memcpy(out_zero, res + off_rzero, (size_t)nq * 4);
  if (zero_flag) std::memcpy(zero_flag, res + off_flag, (size_t)nq);
  return LSHX_OK;
}

extern "C" int lshx_index_debug_timeline(lshx_index* ix, int enable, uint64_t* stamps_out /* 9 */) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (enable && !ix->d_dbg) {
    LSHX_CUDA(cudaMalloc(reinterpret_cast<void**>(&ix->d_dbg), 9 * 8));
    LSHX_CUDA(cudaMemset(ix->d_dbg, 0, 9 * 8));
  }
  if (stamps_out && ix->d_dbg) {
    LSHX_CUDA(cudaStreamSynchronize(ix->stream));
    LSHX_CUDA(cudaMemcpy(stamps_out, ix->d_dbg, 9 * 8, cudaMemcpyDeviceToHost));
  }
  if (!enable && ix->d_dbg) {
    cudaFree(ix->d_dbg);
    ix->d_dbg = nullptr;
  }
  return LSHX_OK;
}

extern "C" int lshx_index_fetch(lshx_index* ix, int64_t* offsets, int32_t* counts, int64_t* ids, int32_t* collisions) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  LSHX_REQUIRE(ix->last_nq >= 0, "no query result on this handle (lshx_index_query first; add / remove drop it)");
  const int64_t nq = ix->last_nq;
  if (nq == 0) {
    if (offsets) offsets[0] = 0;
    return LSHX_OK;
  }
  if (offsets)
    LSHX_CUDA(cudaMemcpyAsync(offsets, ix->raw_off.p, (size_t)(nq + 1) * 8, cudaMemcpyDeviceToHost, ix->stream));
  if (counts) LSHX_CUDA(cudaMemcpyAsync(counts, ix->uniq.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ix->stream));
  if (ids && ix->last_total > 0)
    LSHX_CUDA(cudaMemcpyAsync(ids, ix->out_ids.p, (size_t)ix->last_total * 8, cudaMemcpyDeviceToHost, ix->stream));
  if (collisions && ix->last_total > 0)
    LSHX_CUDA(cudaMemcpyAsync(collisions, ix->out_coll.p, (size_t)ix->last_total * 4, cudaMemcpyDeviceToHost,
                              ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  return LSHX_OK;
}

extern "C" int lshx_index_topk(lshx_index* ix, int top_k, int64_t* out_ids, int32_t* out_count) {
  LSHX_REQUIRE(ix != nullptr, "null handle");
  LSHX_REQUIRE(top_k > 0, "top_k must be greater than zero when provided");
  LSHX_REQUIRE(out_ids != nullptr && out_count != nullptr, "null buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  LSHX_REQUIRE(ix->last_nq >= 0, "no query result on this handle (lshx_index_query first)");
  const int64_t nq = ix->last_nq;
  if (nq == 0) return LSHX_OK;
  int rc;
  if ((rc = ix->topk_ids.reserve((size_t)nq * top_k * 8)) != LSHX_OK) return rc;
  if ((rc = ix->topk_cnt.reserve((size_t)nq * 4)) != LSHX_OK) return rc;
  rc = index_topk(static_cast<const int64_t*>(ix->out_ids.p), static_cast<const int64_t*>(ix->raw_off.p),
                  static_cast<const int*>(ix->uniq.p), nq, top_k, static_cast<int64_t*>(ix->topk_ids.p),
                  static_cast<int*>(ix->topk_cnt.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(out_ids, ix->topk_ids.p, (size_t)nq * top_k * 8, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaMemcpyAsync(out_count, ix->topk_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  return LSHX_OK;
}

extern "C" int lshx_index_rerank(lshx_index* ix, lshx_reranker* r, const float* Q, int q_on_device,
                                 const float* corpus_device, int64_t n_vectors, int k, double p, int out_stride,
                                 int64_t* out_ids, float* out_score, int32_t* out_count, int32_t* out_zero) {
  LSHX_REQUIRE(ix != nullptr && r != nullptr, "null handle");
  LSHX_REQUIRE(ix->device == r->device, "index and reranker live on different devices");
  LSHX_REQUIRE(k > 0 || p > 0.0, "k must be > 0");
  LSHX_REQUIRE(!(p > 1.0), "top_p must be within the range (0, 1]");
  LSHX_REQUIRE(out_stride > 0 && out_ids && out_score && out_count, "null output buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  std::lock_guard<std::mutex> lk2(r->mu);
  DeviceGuard g(ix->device);
  LSHX_REQUIRE(ix->last_nq >= 0, "no query result on this handle (lshx_index_query first)");
  const int64_t nq = ix->last_nq;
  if (nq == 0) return LSHX_OK;
  LSHX_REQUIRE(corpus_device != nullptr && n_vectors > 0, "null buffer");
  const int dim = r->dim;
  int rc;
  const float* d_q = Q;
  if (Q == nullptr) {   // the vectors lshx_index_query_host_vectors uploaded for this very result
    LSHX_REQUIRE(ix->last_q_rows == nq && ix->last_q_dim == dim,
                 "Q is NULL but the handle holds no query vectors for this result (lshx_index_query_host_vectors)");
    d_q = static_cast<const float*>(ix->rr_q.p);
  } else if (!q_on_device) {
    ix->last_q_rows = -1;     // rr_q is about to hold these vectors instead
    if ((rc = ix->rr_q.reserve((size_t)nq * dim * sizeof(float))) != LSHX_OK) return rc;
    if ((rc = upload(ix->rr_q.p, Q, (size_t)nq * dim * sizeof(float), ix->stream)) != LSHX_OK) return rc;
    d_q = static_cast<const float*>(ix->rr_q.p);
  }
  if ((rc = ix->rr_pos.reserve((size_t)nq * out_stride * 4)) != LSHX_OK) return rc;
  if ((rc = ix->rr_score.reserve((size_t)nq * out_stride * 4)) != LSHX_OK) return rc;
  if ((rc = ix->rr_count.reserve((size_t)nq * 4)) != LSHX_OK) return rc;
  if ((rc = ix->rr_zero.reserve((size_t)nq * 4)) != LSHX_OK) return rc;
  if ((rc = ix->rr_ids.reserve((size_t)nq * out_stride * 8)) != LSHX_OK) return rc;
  RerankArgs a{};
  a.Q = d_q;
  a.nq = nq;
  a.V = corpus_device;
  a.n_vectors = n_vectors;
  a.offs = static_cast<const int64_t*>(ix->raw_off.p);
  a.ids = static_cast<const int64_t*>(ix->out_ids.p);        // candidate id = row of the resident corpus
  a.cand_counts = static_cast<const int32_t*>(ix->uniq.p);  // the lists do not fill their slot ranges
  a.dim = dim;
  a.k = k;
  a.p = p;
  a.out_stride = out_stride;
  a.out_pos = static_cast<int32_t*>(ix->rr_pos.p);
  a.out_score = static_cast<float*>(ix->rr_score.p);
  a.out_count = static_cast<int32_t*>(ix->rr_count.p);
  a.out_zero = static_cast<int32_t*>(ix->rr_zero.p);
  a.max_cand = ix->last_max;
  a.select = true;
  if (const size_t need = rerank_scratch_bytes(a, nullptr)) {
    if ((rc = r->big.reserve(need)) != LSHX_OK) return rc;
    a.big_keys = static_cast<uint64_t*>(r->big.p);
    a.big_bytes = r->big.cap;
  }
  if ((rc = launch_rerank(a, ix->stream)) != LSHX_OK) return rc;
  rc = index_pos_to_id(static_cast<const int64_t*>(ix->out_ids.p), static_cast<const int64_t*>(ix->raw_off.p),
                       a.out_pos, a.out_count, nq, out_stride, static_cast<int64_t*>(ix->rr_ids.p), ix->stream);
  if (rc != LSHX_OK) return rc;
  LSHX_CUDA(cudaMemcpyAsync(out_ids, ix->rr_ids.p, (size_t)nq * out_stride * 8, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaMemcpyAsync(out_score, ix->rr_score.p, (size_t)nq * out_stride * 4, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaMemcpyAsync(out_count, ix->rr_count.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ix->stream));
  if (out_zero)
    LSHX_CUDA(cudaMemcpyAsync(out_zero, ix->rr_zero.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ix->stream));
  LSHX_CUDA(cudaStreamSynchronize(ix->stream));
  return LSHX_OK;
}

// ---------------------------------------------------------------------------------------
// multi-process plumbing for the signature relay (lshrs_b200/fabric.py): CUDA IPC memory and events,
// peer / host copies on a caller's stream.  No arithmetic; one rank per GPU under torchrun cannot share
// device memory or order its streams against another process's without these.
// ---------------------------------------------------------------------------------------

static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(cudaIpcEventHandle_t) == 64, "IPC handles are 64 bytes");

extern "C" int lshx_ipc_mem_alloc(int device, size_t bytes, void** dptr, unsigned char* handle_out) {
  LSHX_REQUIRE(dptr != nullptr && handle_out != nullptr && bytes > 0, "bad argument");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);
  *dptr = nullptr;
  LSHX_CUDA(cudaMalloc(dptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
  if (e != cudaSuccess) {
    cudaFree(*dptr);
    *dptr = nullptr;
    LSHX_CUDA(e);
  }
  std::memcpy(handle_out, &h, sizeof(h));
  return LSHX_OK;
}

extern "C" int lshx_ipc_mem_open(int device, const unsigned char* handle, void** dptr) {
  LSHX_REQUIRE(dptr != nullptr && handle != nullptr, "bad argument");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);   // the importing GPU: peer access to the exporting one is enabled lazily
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  LSHX_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return LSHX_OK;
}

extern "C" int lshx_ipc_mem_close(void* dptr) {
  if (dptr) LSHX_CUDA(cudaIpcCloseMemHandle(dptr));
  return LSHX_OK;
}

extern "C" int lshx_ipc_mem_free(int device, void* dptr) {
  if (!dptr) return LSHX_OK;
  DeviceGuard g(device);
  LSHX_CUDA(cudaFree(dptr));
  return LSHX_OK;
}

extern "C" int lshx_ipc_event_create(int device, void** event, unsigned char* handle_out) {
  LSHX_REQUIRE(event != nullptr && handle_out != nullptr, "bad argument");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);
  cudaEvent_t ev = nullptr;
  LSHX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
  cudaIpcEventHandle_t h;
  cudaError_t e = cudaIpcGetEventHandle(&h, ev);
  if (e != cudaSuccess) {
    cudaEventDestroy(ev);
    LSHX_CUDA(e);
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *event = ev;
  return LSHX_OK;
}

extern "C" int lshx_ipc_event_open(int device, const unsigned char* handle, void** event) {
  LSHX_REQUIRE(event != nullptr && handle != nullptr, "bad argument");
  int rc = check_device(device);
  if (rc != LSHX_OK) return rc;
  DeviceGuard g(device);
  cudaIpcEventHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  cudaEvent_t ev = nullptr;
  LSHX_CUDA(cudaIpcOpenEventHandle(&ev, h));
  *event = ev;
  return LSHX_OK;
}

extern "C" int lshx_ipc_event_record(int device, void* event, void* stream) {
  LSHX_REQUIRE(event != nullptr, "null event");
  DeviceGuard g(device);
  LSHX_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
  return LSHX_OK;
}

extern "C" int lshx_ipc_event_wait(int device, void* event, void* stream) {
  LSHX_REQUIRE(event != nullptr, "null event");
  DeviceGuard g(device);
  LSHX_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(event), 0));
  return LSHX_OK;
}

extern "C" int lshx_ipc_event_destroy(void* event) {
  if (event) LSHX_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return LSHX_OK;
}

extern "C" int lshx_memcpy_async(int device, void* dst, const void* src, size_t bytes, void* stream) {
  LSHX_REQUIRE(dst != nullptr && src != nullptr, "null buffer");
  DeviceGuard g(device);
  LSHX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return LSHX_OK;
}
