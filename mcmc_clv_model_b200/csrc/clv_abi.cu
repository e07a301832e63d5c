// clv_abi.cu — host side of libclv_b200.so: the C-ABI declared in include/clv_b200.h.
// Owns device memory, orders the kernels of one Gibbs sweep, streams kept draws back to the host,
// and (customer-sharded mode) all-reduces the int64 level-2 statistics over NCCL each sweep.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <emmintrin.h>
#include <math_constants.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/clv_b200.h"
#include "clv_kernels.cuh"
#include "clv_forecast.cuh"
#include "clv_cbs.cuh"

using namespace clv;

namespace {

thread_local std::string g_last_error;

struct NcclUid { char internal[128]; };
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, NcclUid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    // prefer a copy already mapped into the process (torch's bundled NCCL), else the system one
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl lacks required symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;
constexpr int NCCL_INT64 = 4, NCCL_SUM = 0;
// communicators are created once per (device, rank, world) and shared by every later handle of the process:
// ncclCommInitRank costs seconds, a sampler call should not pay it twice
struct CachedComm { int device, rank, world; nccl_comm_t comm; };
std::vector<CachedComm> g_comms;
// peer mailboxes of the fused all-reduce are likewise created and connected once per (device, rank, world, chains)
struct CachedMailbox {
  int device, rank, world, chains;
  void* base; size_t bytes;
  bool connected;
  void* peer_base[P2P_MAX_WORLD];
};
std::vector<CachedMailbox> g_mailboxes;
unsigned long long g_epoch = 0;     // process-wide: mailbox flags never repeat across handles
std::mutex g_cache_mutex;           // guards g_comms / g_mailboxes / g_epoch (handles may live on different host threads)

}  // namespace

struct clv_sampler {
  clv_config cfg{};
  std::string err;
  int D = 2, K = 1, S = 20, chains = 1, ncol = 4, P = 5;
  long long N = 0;
  bool have_data = false, have_hyper = false, inited = false;
  int sm_count = 148;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  // hyper (host)
  std::vector<double> beta0, A0, gamma0;
  double nu0 = 0;
  // device
  ModelConst* d_mc = nullptr;
  ModelConst h_mc{};
  ChainParams* d_params = nullptr;
  int* d_x = nullptr;
  double *d_tx = nullptr, *d_T = nullptr, *d_Xc = nullptr, *d_logs = nullptr;
  double *d_ll = nullptr, *d_lm = nullptr, *d_le = nullptr, *d_z = nullptr, *d_tau = nullptr;
  unsigned long long* d_acc = nullptr;
  int* d_err = nullptr;
  // per-run buffers
  long long* d_loglik = nullptr; long long loglik_cap = 0;
  double* d_level2 = nullptr; long long level2_cap = 0;
  double* d_draws[2] = {nullptr, nullptr}; long long draws_cap_bytes[2] = {0, 0};
  cudaEvent_t ev_chunk_ready[2] = {nullptr, nullptr}, ev_copy_done[2] = {nullptr, nullptr};
  // resident draws of the last run (single-chunk runs only)
  long long resident_draws = 0;
  // injected staging
  double* d_inj = nullptr; long long inj_cap = 0;
  // persistent mode
  unsigned long long* d_acc3 = nullptr;
  unsigned int* d_barrier = nullptr;
  size_t stats_smem = 0;           // dynamic shared memory of the sweep kernels: (nstat + 1) int64 columns x 128 threads
  int persist_grid_x = 0;          // 0: cooperative launch not possible for this problem
  bool persist_fits = false;       // every tile has its own co-resident block (small problems: AUTO picks persistent)
  // bookkeeping
  long long sweeps_done = 0, launches = 0;
  int grid_x = 1;
  int cpt = 1, grid2_x = 1;        // customers per thread of the sweep kernel (CLV_SWEEP_CPT) and the grid of the 2-customer variant
  unsigned int* d_tile_ctr = nullptr; long long n_big = 0, n_small = 0; int grid2_dyn_x = 0;   // dynamic tiles of k_sweep2 (0: static)
  // comm
  nccl_comm_t comm = nullptr; int world = 1, rank = 0;
  // peer mailboxes (P2P all-reduce fused into k_level2)
  void* d_mailbox = nullptr; size_t mailbox_bytes = 0;
  bool p2p = false;
  unsigned long long* peer_mail[P2P_MAX_WORLD] = {nullptr};
  bool use_pdl = true;                   // programmatic dependent launch of the two kernels of a sweep (CLV_NO_PDL=1 disables)
  // When a kernel lets its successor in the stream become resident (CLV_PDL_MODE overrides): 1 = both kernels at once
  // (best when the sweep grid fills the GPU: the next sweep's blocks move in as this sweep's blocks retire); 2 = k_sweep at
  // once, k_level2 only after its own wait (small grids: otherwise the blocks of several future sweeps pile up beside the
  // running one and slow it down -- measured 33 vs 18 us per sweep at 23 570 customers, profiles/r02_kernel_ab.txt);
  // 3 = k_sweep after its tiles.  0 = chosen from the grid size at create.
  int pdl_mode = 0;
  unsigned long long init_epoch = 0;     // bumped by every clv_init_state: mailbox flags never repeat
  // fused forecast of the kept draws (clv_set_fused_forecast)
  bool fc_enabled = false; double fc_T_star = 39.0; uint64_t fc_seed = 0;
  FusedForecast* d_fc = nullptr; unsigned long long* d_fc_sx = nullptr; unsigned int* d_fc_sz = nullptr;
  long long fc_draws = 0;                // draws per chain accumulated by the last run
  unsigned long long fc_deferred_last = 0;   // cells the last resident forecast handed to its second pass
  // statistics computed by clv_init_state(h, NULL)
  clv_init_stats last_stats{};
  std::vector<double> last_xtx;
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_sweep, ev_l2;
  double t_sweep_ms = 0, t_l2_ms = 0; long long t_n = 0;
};

namespace {

int fail(clv_sampler* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  g_last_error = buf;
  return code;
}

#define CK(h, call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(h, CLV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Device memory comes from the device's default stream-ordered pool with an unlimited release threshold: what a
// handle frees stays cached in the process, so a sampler call does not pay cudaMalloc/cudaFree (measured: up to 1.5 s of
// driver time per call for the GB-sized buffers of a 10 M-customer run) every time.  (The peer mailboxes stay on
// cudaMalloc: CUDA IPC needs it.)
thread_local cudaStream_t t_alloc_stream = nullptr;   // stream the calling API function allocates / frees on

std::mutex g_once_mutex;            // guards the per-device "done" flags below (handles may live on several host threads)

void ensure_pool(int dev) {
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lock(g_once_mutex);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}

// bytes a pool allocation can still obtain: free device memory plus what the pool holds but does not use
size_t available_bytes(int dev) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t)(reserved - used);
  }
  return free_b;
}

cudaError_t pool_malloc(void** p, size_t bytes) { return cudaMallocAsync(p, std::max<size_t>(bytes, 8), t_alloc_stream); }
cudaError_t dfree(void* p) { return p ? cudaFreeAsync(p, t_alloc_stream) : cudaSuccess; }

template <typename T>
cudaError_t dmalloc(T** p, size_t n) { return pool_malloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

// ---- small dense host linear algebra (K <= 16) ------------------------------------------------
bool chol_host(const double* A, double* L, int n) {
  std::fill(L, L + n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int m = 0; m < j; ++m) s -= L[i * n + m] * L[j * n + m];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  return true;
}

// inverse of an SPD matrix through its Cholesky factor
bool spd_inverse(const double* A, double* Ainv, int n) {
  std::vector<double> L(n * n), Li(n * n, 0.0);
  if (!chol_host(A, L.data(), n)) return false;
  for (int c = 0; c < n; ++c) {               // Li = L^-1 (lower)
    for (int r = c; r < n; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int m = c; m < r; ++m) s -= L[r * n + m] * Li[m * n + c];
      Li[r * n + c] = s / L[r * n + r];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int m = std::max(i, j); m < n; ++m) s += Li[m * n + i] * Li[m * n + j];
      Ainv[i * n + j] = s;
    }
  return true;
}

int ceil_log2(double v) {
  int e = 0;
  double p = 1.0;
  while (p < v && e < 1000) { p *= 2.0; ++e; }
  return e;
}

int fx_bits_host(double max_abs) {              // hostmath.fx_bits
  if (!(max_abs > 1.0)) return 51;
  int ex;
  double f = std::frexp(max_abs, &ex);
  return 51 - (f == 0.5 ? ex - 1 : ex);
}

// total * 2^-bits, correctly rounded (hostmath.from_fx)
double i128_scaled(__int128 t, int bits) {
  if (t == 0) return 0.0;
  const bool neg = t < 0;
  unsigned __int128 a = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  int msb = 127;
  while (!((a >> msb) & 1)) --msb;
  double m;
  int sh = 0;
  if (msb <= 52) m = (double)(uint64_t)a;
  else {
    sh = msb - 52;
    unsigned __int128 q = a >> sh, rem = a & ((((unsigned __int128)1) << sh) - 1), half = ((unsigned __int128)1) << (sh - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    m = (double)(uint64_t)q;
  }
  return std::ldexp(neg ? -m : m, sh - bits);
}

// ---- host transfer engine -------------------------------------------------------------------------
// A device->host (or host->device) copy whose host side is ordinary pageable memory runs at 3-5 GB/s through the
// driver's own staging: one thread copies, and a freshly allocated destination faults its pages in on that thread.
// Here the copy is cut into pieces that travel through a small ring of page-locked buffers (DMA at PCIe speed), and a
// few worker threads move each piece between the ring and the caller's array in parallel (first-touch page faults
// included).  The level-1 draws of the reference's full-data run are 12 GB; this is what bounds that run.
// memcpy with non-temporal stores: the destination of a staged transfer (the caller's array, or a ring slot the copy
// engine reads next) is not read again by this core, so its lines need not be fetched before they are overwritten
// (write-allocate makes a plain memcpy move 3 bytes per byte copied) nor displace the cache.  CLV_COPY_NT=0: plain memcpy.
static void copy_stream_stores(char* dst, const char* src, size_t n) {
  static const bool plain = [] { const char* e = getenv("CLV_COPY_NT"); return e && atoi(e) == 0; }();
  if (plain || n < 4096) { memcpy(dst, src, n); return; }
  const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i, src += 64, dst += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i*)src), b = _mm_loadu_si128((const __m128i*)(src + 16));
    const __m128i c = _mm_loadu_si128((const __m128i*)(src + 32)), d = _mm_loadu_si128((const __m128i*)(src + 48));
    _mm_stream_si128((__m128i*)dst, a); _mm_stream_si128((__m128i*)(dst + 16), b);
    _mm_stream_si128((__m128i*)(dst + 32), c); _mm_stream_si128((__m128i*)(dst + 48), d);
  }
  _mm_sfence();
  memcpy(dst, src, n - blocks * 64);
}

class HostCopyPool {
 public:
  static HostCopyPool& get() { static HostCopyPool p; return p; }
  int threads() const { return (int)workers_.size() + 1; }
  // memcpy split over the workers and the calling thread; returns when all parts are done
  void copy(void* dst, const void* src, size_t bytes) {
    const int parts = bytes < (2u << 20) ? 1 : threads();
    if (parts == 1) { copy_stream_stores((char*)dst, (const char*)src, bytes); return; }
    const size_t per = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> g(m_);
      for (int t = 1; t < parts; ++t) {
        const size_t off = std::min(bytes, per * t), len = std::min(per, bytes - off);
        if (len) { jobs_.push_back({(char*)dst + off, (const char*)src + off, len}); ++pending_; }
      }
    }
    cv_.notify_all();
    copy_stream_stores((char*)dst, (const char*)src, std::min(per, bytes));
    std::unique_lock<std::mutex> g(m_);
    done_cv_.wait(g, [&] { return pending_ == 0; });
  }

 private:
  struct Job { char* dst; const char* src; size_t len; };
  HostCopyPool() {
    int n = std::min((int)std::thread::hardware_concurrency() / 2, 8);
    if (const char* env = getenv("CLV_COPY_THREADS")) n = atoi(env);
    n = std::max(1, std::min(n, 32));
    for (int t = 1; t < n; ++t) workers_.emplace_back([this] { run(); });
  }
  ~HostCopyPool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || !jobs_.empty(); });
        if (jobs_.empty()) return;
        j = jobs_.back();
        jobs_.pop_back();
      }
      copy_stream_stores(j.dst, j.src, j.len);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Job> jobs_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  int pending_ = 0;
  bool stop_ = false;
};

// bytes per ring slot (page-locking costs ~1.5 ms per MB: keep the ring small); CLV_STAGE_PIECE_MB overrides
static const size_t STAGE_PIECE = [] { const char* e = getenv("CLV_STAGE_PIECE_MB"); return (size_t)std::max(1, e ? atoi(e) : 16) << 20; }();
constexpr int STAGE_SLOTS = 3;
struct StagingRing {
  char* slot[STAGE_SLOTS] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[STAGE_SLOTS] = {nullptr, nullptr, nullptr};
  bool ok = false;
};
// Rings are leased for the duration of one copy and returned to a process-wide free list (the chains-over-devices
// driver runs one sampler per short-lived thread: a ring per thread would leak page-locked memory).  At most one ring
// per concurrently copying thread ever exists; they live until the process ends.
class RingLease {
 public:
  RingLease() {
    cudaGetDevice(&dev_);                       // the ring's events belong to the device that was current at creation
    {
      std::lock_guard<std::mutex> g(mutex());
      auto& fl = free_list(dev_);
      if (!fl.empty()) { r_ = fl.back(); fl.pop_back(); }
    }
    if (!r_) {
      r_ = new StagingRing();
      bool good = true;
      for (int i = 0; i < STAGE_SLOTS && good; ++i)
        good = cudaHostAlloc((void**)&r_->slot[i], STAGE_PIECE, cudaHostAllocPortable) == cudaSuccess &&
               cudaEventCreateWithFlags(&r_->ev[i], cudaEventDisableTiming) == cudaSuccess;
      if (!good) cudaGetLastError();
      r_->ok = good;
    }
  }
  ~RingLease() {
    std::lock_guard<std::mutex> g(mutex());
    free_list(dev_).push_back(r_);
  }
  StagingRing& ring() { return *r_; }

 private:
  static std::mutex& mutex() { static std::mutex m; return m; }
  static std::vector<StagingRing*>& free_list(int dev) {          // call with mutex() held
    static std::vector<std::vector<StagingRing*>>* v = new std::vector<std::vector<StagingRing*>>();
    if ((int)v->size() <= dev) v->resize(dev + 1);
    return (*v)[std::max(dev, 0)];
  }
  StagingRing* r_ = nullptr;
  int dev_ = 0;
};
size_t staging_min_bytes() {
  const char* e = getenv("CLV_STAGING_MIN_BYTES");      // smaller copies take the driver's own pageable path
  return e ? (size_t)atoll(e) : (size_t)(64u << 20);
}

// page-locked (cudaHostAlloc / cudaHostRegister) memory goes straight to the copy engine
bool is_page_locked(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// First touch of a caller's output array, ahead of the copies.  The level-1 draws of a run land in a FRESH array
// (np.empty; NumPy advises huge pages for it): the first write to every page makes the kernel allocate and ZERO it, in the
// thread that writes -- here the copy threads, in the middle of the transfer.  Measured on the C2 shape x 4 chains (12 GB of
// draws, tools/d2h_probe.py): 1.15 s into a fresh array, 0.57 s into the same array again, 0.38 s without level-1 output:
// zeroing 12 GB costs more than sampling.  While the GPU runs the burn-in the host has nothing to do, so a few threads
// touch the pages in the order the run will fill them.  The touch is `lock or byte, 0`: it takes the write fault but
// leaves the byte as it is, atomically -- a copy that has already delivered data there (the toucher fell behind) loses
// nothing.  The threads stop when the run ends.
class FirstToucher {
 public:
  FirstToucher() = default;
  FirstToucher(const FirstToucher&) = delete;
  FirstToucher& operator=(const FirstToucher&) = delete;
  ~FirstToucher() { stop(); }
  // ranges are taken in the given order by n_threads threads
  void start(std::vector<std::pair<char*, size_t>> ranges, int n_threads) {
    ranges_ = std::move(ranges);
    if (ranges_.empty() || n_threads < 1) return;
    try {
      for (int t = 0; t < n_threads; ++t)
        th_.emplace_back([this] {
          for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= ranges_.size()) return;
            char* p = ranges_[i].first;
            char* const end = p + ranges_[i].second;
            for (; p < end; p += 4096) {
              if (stop_.load(std::memory_order_relaxed)) return;
              __atomic_fetch_or((unsigned char*)p, (unsigned char)0, __ATOMIC_RELAXED);
            }
          }
        });
    } catch (...) {
      // no more threads to be had: the ones that started carry on, the copies take the remaining faults themselves
    }
  }
  void stop() {
    stop_.store(true);
    finish();
  }
  void finish() {                       // wait until every range has been touched (or stop() was called)
    for (auto& t : th_) t.join();
    th_.clear();
  }

 private:
  std::vector<std::pair<char*, size_t>> ranges_;
  std::vector<std::thread> th_;
  std::atomic<size_t> next_{0};
  std::atomic<bool> stop_{false};
};

// Half of this process's share of the cores, at most 8 (measured on a 16-core box: 6 / 8 / 12 / 16 threads give the same
// times).  Under torchrun the ranks of a node share its cores (LOCAL_WORLD_SIZE): the thread that launches the sweeps must
// keep a core to itself, or the burn-in it is supposed to overlap with slows down.
static int first_touch_threads() {
  int ranks = 1;
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(lw));
  int nt = std::max(1, std::min(8, (int)std::thread::hardware_concurrency() / (2 * ranks)));
  if (const char* te = getenv("CLV_FIRST_TOUCH_THREADS")) nt = std::max(1, std::min(32, atoi(te)));
  return nt;
}
static bool first_touch_wanted(const void* p, size_t bytes) {
  const char* fe = getenv("CLV_FIRST_TOUCH");
  return !(fe && atoi(fe) == 0) && p && bytes >= staging_min_bytes() && !is_page_locked(p);
}

// device -> pageable host, ordered after everything already in `stream`; returns when the data is in `dst`
cudaError_t copy_to_host_staged(void* dst, const void* src_dev, size_t bytes, cudaStream_t stream) {
  if (bytes < staging_min_bytes() || is_page_locked(dst)) return cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
  RingLease lease;
  StagingRing& r = lease.ring();
  if (!r.ok) return cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
  HostCopyPool& pool = HostCopyPool::get();
  const size_t n = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
  auto issue = [&](size_t i) -> cudaError_t {
    const size_t off = i * STAGE_PIECE, len = std::min(STAGE_PIECE, bytes - off);
    cudaError_t e = cudaMemcpyAsync(r.slot[i % STAGE_SLOTS], (const char*)src_dev + off, len, cudaMemcpyDeviceToHost, stream);
    return e != cudaSuccess ? e : cudaEventRecord(r.ev[i % STAGE_SLOTS], stream);
  };
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < std::min<size_t>(n, STAGE_SLOTS - 1) && e == cudaSuccess; ++i) e = issue(i);
  for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
    if (i + STAGE_SLOTS - 1 < n) e = issue(i + STAGE_SLOTS - 1);      // its slot was emptied in iteration i - 1
    if (e == cudaSuccess) e = cudaEventSynchronize(r.ev[i % STAGE_SLOTS]);
    const size_t off = i * STAGE_PIECE, len = std::min(STAGE_PIECE, bytes - off);
    if (e == cudaSuccess) pool.copy((char*)dst + off, r.slot[i % STAGE_SLOTS], len);
  }
  return e;
}

// pageable host -> device; returns when the last piece has been handed to the copy engine (stream ordered after that)
cudaError_t copy_to_device_staged(void* dst_dev, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes < staging_min_bytes() || is_page_locked(src)) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, stream);
  RingLease lease;
  StagingRing& r = lease.ring();
  if (!r.ok) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, stream);
  HostCopyPool& pool = HostCopyPool::get();
  const size_t n = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
    const int sl = (int)(i % STAGE_SLOTS);
    if (i >= STAGE_SLOTS) e = cudaEventSynchronize(r.ev[sl]);          // the DMA that last read this slot
    const size_t off = i * STAGE_PIECE, len = std::min(STAGE_PIECE, bytes - off);
    if (e == cudaSuccess) pool.copy(r.slot[sl], (const char*)src + off, len);
    if (e == cudaSuccess) e = cudaMemcpyAsync((char*)dst_dev + off, r.slot[sl], len, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaEventRecord(r.ev[sl], stream);
  }
  // the ring is reused by the next call: its slots must have been read
  for (int sl = 0; sl < STAGE_SLOTS && e == cudaSuccess; ++sl)
    if ((size_t)sl < n) e = cudaEventSynchronize(r.ev[sl]);
  return e;
}

SweepArgs base_args(clv_sampler* h) {
  SweepArgs a{};
  a.mc = h->d_mc;
  a.params = h->d_params;
  a.x = h->d_x; a.t_x = h->d_tx; a.T_cal = h->d_T; a.Xc = h->d_Xc; a.log_s = h->d_logs;
  a.ll = h->d_ll; a.lm = h->d_lm; a.le = h->d_le; a.z = h->d_z; a.tau = h->d_tau;
  a.acc = h->d_acc;
  a.loglik_acc = h->d_loglik;
  a.loglik_stride = 1;
  a.draws = nullptr; a.chunk_cap = 1; a.slot = -1; a.draw_index = 0;
  a.sweep = 0; a.chain_offset = (uint32_t)h->cfg.chain_offset; a.seed = h->cfg.seed;
  a.rk = round_keys(h->cfg.seed);
  a.store_zt = 0;
  a.error_flag = h->p2p ? h->d_err : nullptr;     // only sharded runs can be told to stop by a peer
  a.pdl_early = (h->pdl_mode == 3) ? 0 : 1;
  a.fc = nullptr;                                  // set by the run driver for sweeps that may keep a draw
  a.tile_counter = (h->cpt == 2 && h->grid2_dyn_x > 0) ? h->d_tile_ctr : nullptr;
  a.n_big = h->n_big; a.n_small = h->n_small;
  return a;
}

cudaEvent_t pool_event(clv_sampler* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}

// Launch with (or without) the programmatic-stream-serialization attribute: the kernel may then become resident while
// its predecessor in the stream still runs and execute everything it has before griddepcontrol.wait.
template <typename Arg>
cudaError_t launch_kernel(void (*kern)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Arg& arg) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, arg);
}

template <int D>
cudaError_t launch_sweep_kernel(clv_sampler* h, const SweepArgs& a, int mode, bool pdl) {
  dim3 grid(h->grid_x, h->chains), block(SWEEP_THREADS);
  const size_t sm = h->stats_smem;
  if (a.fc) {      // fused forecast of the kept draws: its own instantiations
    if (mode == MODE_STRICT) return launch_kernel(k_sweep<D, MODE_STRICT, true>, grid, block, sm, h->stream, pdl, a);
    return launch_kernel(k_sweep<D, MODE_FAST, true>, grid, block, sm, h->stream, pdl, a);
  }
  if (h->cpt == 2 && mode != MODE_INJECT) {      // two customers per thread: tiles of 256, its own grid
    dim3 grid2(a.tile_counter ? h->grid2_dyn_x : h->grid2_x, h->chains);
    if (mode == MODE_STRICT) return launch_kernel(k_sweep2<D, MODE_STRICT>, grid2, block, sm, h->stream, pdl, a);
    return launch_kernel(k_sweep2<D, MODE_FAST>, grid2, block, sm, h->stream, pdl, a);
  }
  if (mode == MODE_FAST) return launch_kernel(k_sweep<D, MODE_FAST>, grid, block, sm, h->stream, pdl, a);
  if (mode == MODE_STRICT) return launch_kernel(k_sweep<D, MODE_STRICT>, grid, block, sm, h->stream, pdl, a);
  return launch_kernel(k_sweep<D, MODE_INJECT>, grid, block, sm, h->stream, pdl, a);
}

int allreduce_acc(clv_sampler* h) {
  if (!h->comm || h->p2p) return 0;    // p2p: the level-2 kernel reduces over the peer mailboxes itself
  int r = g_nccl.AllReduce(h->d_acc, h->d_acc, (size_t)h->chains * NSTAT_MAX, NCCL_INT64, NCCL_SUM, h->comm, h->stream);
  if (r != 0) return fail(h, CLV_ERR_COMM, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return 0;
}

// One Gibbs sweep in the reference's block order (bi:387-399 / tri:512-536).
int enqueue_sweep(clv_sampler* h, SweepArgs a, Level2Args l2, int mode) {
  // tag of this sweep's mailbox words: never 0, differs from the tag of sweep - 2 (same parity slot) and of earlier inits
  l2.tag = (uint32_t)(((h->init_epoch % 255ull) + 1ull) << 24) | (l2.sweep & 0xffffffu);
  // injected variates are uploaded between the kernels, and per-kernel timing records events there: plain launches
  const bool pdl = h->use_pdl && mode != MODE_INJECT && !h->timing;
  cudaError_t le = cudaSuccess;
  auto do_l2 = [&]() {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) { e0 = pool_event(h); e1 = pool_event(h); cudaEventRecord(e0, h->stream); }
    cudaError_t e = (h->D == 2) ? launch_kernel(k_level2<2>, dim3(h->chains), dim3(32), 0, h->stream, pdl, l2)
                                : launch_kernel(k_level2<3>, dim3(h->chains), dim3(32), 0, h->stream, pdl, l2);
    if (le == cudaSuccess) le = e;
    if (h->timing) { cudaEventRecord(e1, h->stream); h->ev_l2.push_back({e0, e1}); }
    h->launches++;
  };
  auto do_sweep = [&]() {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) { e0 = pool_event(h); e1 = pool_event(h); cudaEventRecord(e0, h->stream); }
    cudaError_t e = (h->D == 2) ? launch_sweep_kernel<2>(h, a, mode, pdl) : launch_sweep_kernel<3>(h, a, mode, pdl);
    if (le == cudaSuccess) le = e;
    if (h->timing) { cudaEventRecord(e1, h->stream); h->ev_sweep.push_back({e0, e1}); }
    h->launches++;
  };
  if (h->D == 2) {
    if (int r = allreduce_acc(h)) return r;
    do_l2();
    do_sweep();
  } else {
    do_sweep();
    if (int r = allreduce_acc(h)) return r;
    do_l2();
  }
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) return fail(h, CLV_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(le));
  h->sweeps_done++;
  return 0;
}

// How long a rank waits for its peers' mailbox words before it gives up (CLV_ERR_COMM).  Ranks are separate processes:
// start-up skew, a rank blocked in a pageable flush or a progress callback are all legitimate, so the default is long.
long long p2p_timeout_ns() {
  static const long long v = [] {
    const char* e = getenv("CLV_P2P_TIMEOUT_S");
    const double sec = e ? atof(e) : 120.0;
    return (long long)(std::max(0.001, sec) * 1e9);
  }();
  return v;
}

Level2Args base_l2(clv_sampler* h) {
  Level2Args l{};
  l.mc = h->d_mc; l.params = h->d_params; l.acc = h->d_acc;
  l.level2_draws = h->d_level2; l.n_draws = 1; l.draw_index = -1;
  l.sweep = 0; l.chain_offset = (uint32_t)h->cfg.chain_offset; l.seed = h->cfg.seed;
  l.injected = 0; l.iw_norm = l.iw_chi2 = l.beta_norm = nullptr;
  l.error_flag = h->d_err;
  l.world = h->p2p ? h->world : 0; l.rank = h->rank; l.n_chains = h->chains;
  l.tag = 1u;
  l.pdl_early = (h->pdl_mode == 2) ? 0 : 1;
  l.timeout_ns = p2p_timeout_ns();
  for (int r = 0; r < P2P_MAX_WORLD; ++r) l.peer_mail[r] = h->peer_mail[r];
  return l;
}

int collect_timing(clv_sampler* h) {
  for (auto& p : h->ev_sweep) { float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second); h->t_sweep_ms += ms; h->t_n++; }
  for (auto& p : h->ev_l2) { float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second); h->t_l2_ms += ms; }
  h->ev_sweep.clear(); h->ev_l2.clear(); h->ev_used = 0;
  return 0;
}

int check_device_error(clv_sampler* h) {
  int flag = 0;
  CK(h, cudaMemcpyAsync(&flag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  if (h->timing) collect_timing(h);
  if (flag == 2) return fail(h, CLV_ERR_COMM, "peer mailbox all-reduce timed out after CLV_P2P_TIMEOUT_S (a rank stopped?) at sweep <= %lld; the remaining sweeps were skipped", h->sweeps_done);
  if (flag) return fail(h, CLV_ERR_NUMERIC, "level-2 scale matrix not positive definite or non-finite (sweep <= %lld)", h->sweeps_done);
  return 0;
}

int recompute_stats(clv_sampler* h) {
  CK(h, cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * h->chains * NSTAT_MAX, h->stream));
  if (h->D == 2) {   // bivariate: the next sweep starts with a level-2 draw from the current state (bi:393)
    SweepArgs a = base_args(h);
    dim3 grid(h->grid_x, h->chains);
    k_stats_only<2><<<grid, SWEEP_THREADS, h->stats_smem, h->stream>>>(a);
    h->launches++;
    CK(h, cudaGetLastError());
  }
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int clv_abi_version(void) { return CLV_ABI_VERSION; }

const char* clv_last_error(const clv_sampler* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

static int upload_rk();   // reciprocal tables of the Poisson inversions (constant memory, once per device)

int clv_create(clv_sampler** out, const clv_config* cfg) {
  if (!out || !cfg) return fail(nullptr, CLV_ERR_ARG, "clv_create: null argument");
  *out = nullptr;
  if (cfg->model_dim != 2 && cfg->model_dim != 3) return fail(nullptr, CLV_ERR_ARG, "model_dim must be 2 or 3");
  if (cfg->n_cov < 1 || cfg->n_cov > CLV_MAX_K) return fail(nullptr, CLV_ERR_ARG, "n_cov must be in [1, %d]", CLV_MAX_K);
  if (cfg->n_chains < 1 || cfg->n_chains > 65535) return fail(nullptr, CLV_ERR_ARG, "n_chains must be in [1, 65535]");
  if (cfg->n_mh_steps < 0 || cfg->n_mh_steps > 30000) return fail(nullptr, CLV_ERR_ARG, "n_mh_steps out of range");
  if (cfg->n_local < 1 || cfg->n_global < cfg->n_local || cfg->gid_offset < 0 ||
      cfg->gid_offset + cfg->n_local > cfg->n_global || cfg->n_global > 0xFFFFFFFFll)
    return fail(nullptr, CLV_ERR_ARG, "bad n_local/n_global/gid_offset");
  if (cfg->rng_mode < 0 || cfg->rng_mode > 2) return fail(nullptr, CLV_ERR_ARG, "bad rng_mode");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CLV_ERR_ARG, "device %d out of range (have %d)", cfg->device, ndev);
  clv_sampler* h = new clv_sampler();
  h->cfg = *cfg;
  h->D = cfg->model_dim; h->K = cfg->n_cov; h->S = cfg->n_mh_steps; h->chains = cfg->n_chains;
  h->use_pdl = getenv("CLV_NO_PDL") == nullptr;
  if (const char* e = getenv("CLV_PDL_MODE")) h->pdl_mode = atoi(e);
  h->N = cfg->n_local; h->ncol = h->D == 2 ? 4 : 5; h->P = h->D * h->K + h->D * (h->D + 1) / 2;
  auto bail = [&](int code) { std::string m = h->err; clv_destroy(h); g_last_error = m; return code; };
#define CKC(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { fail(h, CLV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2)); return bail(CLV_ERR_CUDA); } } while (0)
  CKC(cudaSetDevice(cfg->device));
  const bool trace = getenv("CLV_TRACE_CREATE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto t_begin = now();
  CKC(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));   // not cudaGetDeviceProperties: that one costs milliseconds
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  t_alloc_stream = h->stream;
  ensure_pool(cfg->device);
  for (int b = 0; b < 2; ++b) {
    CKC(cudaEventCreateWithFlags(&h->ev_chunk_ready[b], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->ev_copy_done[b], cudaEventDisableTiming));
  }
  const size_t N = (size_t)h->N, C = (size_t)h->chains;
  CKC(dmalloc(&h->d_mc, 1));
  CKC(dmalloc(&h->d_params, C));
  CKC(dmalloc(&h->d_x, N));
  CKC(dmalloc(&h->d_tx, N));
  CKC(dmalloc(&h->d_T, N));
  CKC(dmalloc(&h->d_Xc, N * (size_t)std::max(h->K - 1, 1)));
  CKC(dmalloc(&h->d_logs, h->D == 3 ? N : 1));
  CKC(dmalloc(&h->d_ll, C * N));
  CKC(dmalloc(&h->d_lm, C * N));
  CKC(dmalloc(&h->d_le, h->D == 3 ? C * N : 1));
  CKC(dmalloc(&h->d_z, C * N));
  CKC(dmalloc(&h->d_tau, C * N));
  CKC(dmalloc(&h->d_acc, C * NSTAT_MAX));
  CKC(dmalloc(&h->d_err, 1));
  if (trace) fprintf(stderr, "[clv_create] streams/events + allocation calls %.2f ms\n", ms_since(t_begin));
  CKC(cudaStreamSynchronize(h->stream));      // pool allocations are stream ordered; the memsets below use the legacy stream
  if (trace) fprintf(stderr, "[clv_create] ... allocations complete %.2f ms\n", ms_since(t_begin));
  CKC(cudaMemset(h->d_err, 0, sizeof(int)));
  CKC(cudaMemset(h->d_acc, 0, sizeof(unsigned long long) * C * NSTAT_MAX));
  CKC(cudaMemset(h->d_params, 0, sizeof(ChainParams) * C));
  {
    double et[EXP_N];
    for (int j = 0; j < EXP_N; ++j) et[j] = (double)exp2l((long double)j / (long double)EXP_N);
    CKC(cudaMemcpyToSymbol(c_exptab, et, sizeof et));
  }
  if (upload_rk()) { h->err = g_last_error; return bail(CLV_ERR_CUDA); }
  if (trace) fprintf(stderr, "[clv_create] ... memsets + constants %.2f ms\n", ms_since(t_begin));
  // grid: a few resident waves of 128-thread blocks, grid-stride over customer tiles
  long long ntiles = (h->N + SWEEP_THREADS - 1) / SWEEP_THREADS;
  long long per_sm_blocks = 24;       // 3 waves of 8 resident blocks (measured best at 1.25 M customers per GPU)
  if (const char* env = getenv("CLV_SWEEP_BLOCKS_PER_SM")) per_sm_blocks = std::max(1ll, atoll(env));   // tuning knob
  long long want = ((long long)h->sm_count * per_sm_blocks + h->chains - 1) / h->chains;
  h->grid_x = (int)std::max<long long>(1, std::min(ntiles, want));
  // Customers per thread of the sweep kernel: two when the problem fills the GPU (the two independent instruction streams
  // per warp raise the issue rate: 1.514 vs 1.617 ms per 10 M-customer sweep), one for small problems, where the sweep is
  // bound by the latency of one customer's 20 dependent steps (13.7 vs 19.4 us per sweep at 4 x 2 357 customers).
  h->cpt = ((long long)h->N * h->chains >= 100000) ? 2 : 1;
  if (const char* env = getenv("CLV_SWEEP_CPT")) h->cpt = atoi(env) == 2 ? 2 : 1;
  {
    const long long ntiles2 = (h->N + CPT * SWEEP_THREADS - 1) / (CPT * SWEEP_THREADS);
    // grid of the two-customer kernel (5 resident blocks per SM): 8 waves when that still leaves >= 4 tiles per block, else 4
    // (1.25 M customers: 12 / 16 / 20 / 24 blocks per SM -> 204.0 / 206.0 / 202.2 / 205.2 us; 10 M: 24 / 30 / 40 / 60 -> 1521 / 1513 / 1508 / 1512 us)
    long long per2 = (ntiles2 * h->chains >= 40ll * h->sm_count * 4) ? 40 : 20;
    if (const char* env = getenv("CLV_SWEEP_BLOCKS_PER_SM2")) per2 = std::max(1ll, atoll(env));
    h->grid2_x = (int)std::max<long long>(1, std::min(ntiles2, ((long long)h->sm_count * per2 + h->chains - 1) / h->chains));
  }
  {
    // dynamic tiles for k_sweep2 (CLV_SWEEP_DYNAMIC=0 switches to the static grid-stride split): a one-wave grid, and the last
    // rounds' worth of customers cut into 128-customer tiles
    const char* env = getenv("CLV_SWEEP_DYNAMIC");
    const bool dyn = CPT >= 2 && !(env && atoi(env) == 0);
    const long long resident = std::max<long long>(1, (long long)h->sm_count * CLV_MINBLOCKS2 / h->chains);
    long long small_rounds = 2;          // E32 kernel, 1 vs 2 rounds: 1.25 M 173.3 / 171.3 us, 2.5 M 335.0 / 331.7, 10 M 1268.8 / 1260.8 (profiles/r02_kernel_ab.txt)
    if (const char* e2 = getenv("CLV_SWEEP_SMALL_ROUNDS")) small_rounds = std::max(0ll, atoll(e2));
    const long long small_cust = std::min<long long>(h->N / 4, resident * small_rounds * SWEEP_THREADS);   // at most a quarter of the shard
    h->n_big = (h->N - small_cust) / (CPT * SWEEP_THREADS);
    h->n_small = (h->N - h->n_big * CPT * SWEEP_THREADS + SWEEP_THREADS - 1) / SWEEP_THREADS;
    // only when a block has at least two tiles to draw: with one tile per block (many chains of a small data set) the counter
    // and the finer tail only cost (56 x 2 357 customers: 37.2 vs 31.4 us per sweep)
    h->grid2_dyn_x = (dyn && h->n_big + h->n_small >= 2 * resident) ? (int)resident : 0;
    CKC(dmalloc(&h->d_tile_ctr, (size_t)2 * C));
    CKC(cudaMemsetAsync(h->d_tile_ctr, 0, sizeof(unsigned int) * 2 * C, h->stream));
  }
  if (h->pdl_mode == 0) {
    const long long blocks = (long long)(h->cpt == 2 ? (h->grid2_dyn_x > 0 ? h->grid2_dyn_x : h->grid2_x) : h->grid_x) * h->chains, resident = (long long)(h->cpt == 2 ? CLV_MINBLOCKS2 : CLV_MINBLOCKS) * h->sm_count;
    h->pdl_mode = (blocks >= resident) ? 1 : 2;
  }
  h->stats_smem = (size_t)(h->K * h->D + h->D * (h->D + 1) / 2 + 1) * SWEEP_THREADS * sizeof(long long);
  if (h->stats_smem > 48 * 1024) {
    const int bytes = (int)h->stats_smem;
    cudaFuncSetAttribute(k_sweep<2, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<2, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<2, MODE_INJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<3, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<3, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<3, MODE_INJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep2<2, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep2<2, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep2<3, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep2<3, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<2, MODE_FAST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<2, MODE_STRICT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<3, MODE_FAST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_sweep<3, MODE_STRICT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_stats_only<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_persistent<2, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_persistent<2, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_persistent<3, MODE_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_persistent<3, MODE_STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  }
  // persistent cooperative mode: all blocks must be co-resident
  CKC(dmalloc(&h->d_acc3, 3 * C * NSTAT_MAX));
  CKC(dmalloc(&h->d_barrier, 2));
  CKC(cudaStreamSynchronize(h->stream));
  CKC(cudaMemset(h->d_barrier, 0, 2 * sizeof(unsigned int)));
  {
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
    // occupancy of the instantiation this handle will launch (STRICT may need more registers than FAST)
    const bool strict = cfg->rng_mode == CLV_RNG_PHILOX_STRICT;
    cudaError_t eo;
    if (h->D == 2) eo = strict ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent<2, MODE_STRICT>, SWEEP_THREADS, h->stats_smem)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent<2, MODE_FAST>, SWEEP_THREADS, h->stats_smem);
    else eo = strict ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent<3, MODE_STRICT>, SWEEP_THREADS, h->stats_smem)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent<3, MODE_FAST>, SWEEP_THREADS, h->stats_smem);
    long long maxb = (eo == cudaSuccess && coop) ? (long long)per_sm * h->sm_count : 0;
    if (maxb >= h->chains) {
      h->persist_grid_x = (int)std::min<long long>(ntiles, maxb / h->chains);
      h->persist_fits = (long long)h->persist_grid_x == ntiles;
    }
  }
  if (trace) fprintf(stderr, "[clv_create] ... attributes + occupancy %.2f ms\n", ms_since(t_begin));
#undef CKC
  *out = h;
  return CLV_OK;
}

void clv_destroy(clv_sampler* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  t_alloc_stream = h->stream;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->d_acc3) dfree(h->d_acc3);
  if (h->d_barrier) dfree(h->d_barrier);
  if (h->d_tile_ctr) dfree(h->d_tile_ctr);
  if (h->d_fc) dfree(h->d_fc);
  if (h->d_fc_sx) dfree(h->d_fc_sx);
  if (h->d_fc_sz) dfree(h->d_fc_sz);
  void* ptrs[] = {h->d_mc, h->d_params, h->d_x, h->d_tx, h->d_T, h->d_Xc, h->d_logs, h->d_ll, h->d_lm, h->d_le,
                  h->d_z, h->d_tau, h->d_acc, h->d_err, h->d_loglik, h->d_level2, h->d_draws[0], h->d_draws[1], h->d_inj};
  for (void* p : ptrs) if (p) dfree(p);
  for (int b = 0; b < 2; ++b) {
    if (h->ev_chunk_ready[b]) cudaEventDestroy(h->ev_chunk_ready[b]);
    if (h->ev_copy_done[b]) cudaEventDestroy(h->ev_copy_done[b]);
  }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->stream) cudaStreamSynchronize(h->stream);   // the frees above are stream ordered
  t_alloc_stream = nullptr;
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
}

// the columns every model needs (x, t_x, T_cal, log_s): enqueued on the handle's stream, not synchronised
static int upload_cbs_columns(clv_sampler* h, const int32_t* x, const double* t_x, const double* T_cal, const double* log_s) {
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const size_t N = (size_t)h->N;
  CK(h, cudaMemcpyAsync(h->d_x, x, N * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(h, copy_to_device_staged(h->d_tx, t_x, N * sizeof(double), h->stream));
  CK(h, copy_to_device_staged(h->d_T, T_cal, N * sizeof(double), h->stream));
  if (h->D == 3) CK(h, cudaMemcpyAsync(h->d_logs, log_s, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return CLV_OK;
}

// The frame column by column (a DataFrame's own layout): covariate k lands in its SoA column directly -- no row-major
// matrix on the host, no intercept column over PCIe (8 B per customer), no split kernel.
int clv_set_data_columns(clv_sampler* h, const int32_t* x, const double* t_x, const double* T_cal, const double* const* cov,
                         const double* log_s) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!x || !t_x || !T_cal || (h->K > 1 && !cov)) return fail(h, CLV_ERR_ARG, "clv_set_data_columns: null column");
  for (int k = 0; k + 1 < h->K; ++k)
    if (!cov[k]) return fail(h, CLV_ERR_ARG, "clv_set_data_columns: covariate column %d is null", k);
  if (h->D == 3 && !log_s) return fail(h, CLV_ERR_ARG, "clv_set_data_columns: log_s is required for the trivariate model");
  if (int rc = upload_cbs_columns(h, x, t_x, T_cal, log_s)) return rc;
  const size_t N = (size_t)h->N;
  for (int k = 0; k + 1 < h->K; ++k)
    CK(h, copy_to_device_staged(h->d_Xc + (size_t)k * N, cov[k], N * sizeof(double), h->stream));
  cudaError_t es = cudaStreamSynchronize(h->stream);
  if (es != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_set_data_columns failed: %s", cudaGetErrorString(es));
  h->have_data = true;
  h->inited = false;
  return CLV_OK;
}

int clv_set_data(clv_sampler* h, const int32_t* x, const double* t_x, const double* T_cal, const double* X,
                 const double* log_s) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!x || !t_x || !T_cal || !X) return fail(h, CLV_ERR_ARG, "clv_set_data: null column");
  if (h->D == 3 && !log_s) return fail(h, CLV_ERR_ARG, "clv_set_data: log_s is required for the trivariate model");
  if (int rc = upload_cbs_columns(h, x, t_x, T_cal, log_s)) return rc;
  const size_t N = (size_t)h->N;
  double* d_rows = nullptr;
  int bad_intercept = 0;
  if (h->K > 1) {
    // row-major N x K (column 0 = intercept): one contiguous copy, then split into SoA columns on the device
    CK(h, dmalloc(&d_rows, N * (size_t)h->K));
    cudaError_t e = cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
    if (e == cudaSuccess) e = copy_to_device_staged(d_rows, X, N * (size_t)h->K * sizeof(double), h->stream);
    if (e == cudaSuccess) {
      k_split_columns<<<h->sm_count * 8, 256, 0, h->stream>>>(d_rows, (long long)N, h->K, h->d_Xc, h->d_err);
      h->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad_intercept, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
    if (e != cudaSuccess) { dfree(d_rows); return fail(h, CLV_ERR_CUDA, "design-matrix upload failed: %s", cudaGetErrorString(e)); }
  } else {
    for (size_t i = 0; i < N; ++i)
      if (X[i] != 1.0) { bad_intercept = 1; break; }
  }
  cudaError_t es = cudaStreamSynchronize(h->stream);
  if (d_rows) dfree(d_rows);
  if (es != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_set_data failed: %s", cudaGetErrorString(es));
  if (bad_intercept) return fail(h, CLV_ERR_ARG, "column 0 of X must be the intercept (all ones)");
  h->have_data = true;
  h->inited = false;
  return CLV_OK;
}

int clv_set_hyper(clv_sampler* h, const double* beta0, const double* A0, double nu0, const double* gamma0) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!beta0 || !A0 || !gamma0) return fail(h, CLV_ERR_ARG, "clv_set_hyper: null argument");
  h->beta0.assign(beta0, beta0 + h->K * h->D);
  h->A0.assign(A0, A0 + h->K * h->K);
  h->gamma0.assign(gamma0, gamma0 + h->D * h->D);
  h->nu0 = nu0;
  h->have_hyper = true;
  h->inited = false;
  return CLV_OK;
}

// Exact, shard-independent initialisation statistics on the device (+ NCCL when sharded); see k_init_quantities.
static int device_init_stats(clv_sampler* h, clv_init_stats* out, std::vector<double>& xtx) {
  const int K = h->K, D = h->D;
  const int nqA = 3 + K * (K + 1) / 2;
  unsigned long long* d_max = nullptr;
  long long* d_sum = nullptr;
  CK(h, dmalloc(&d_max, (size_t)NQ_MAX));
  CK(h, dmalloc(&d_sum, (size_t)NQ_MAX * 2));
  int rc = 0;
  InitQArgs a{};
  a.x = h->d_x; a.t_x = h->d_tx; a.T_cal = h->d_T; a.Xc = h->d_Xc; a.log_s = h->d_logs;
  a.N = h->N; a.K = K; a.D = D; a.out_max = d_max; a.out_sum = d_sum;
  const double n = (double)h->cfg.n_global;
  std::vector<double> tot(NQ_MAX, 0.0);
  double max_abs_x = 1.0;
  // K <= 5: the register-accumulating instantiation (same totals bit for bit); CLV_INIT_GENERIC=1 forces the general kernel
  const char* genv = getenv("CLV_INIT_GENERIC");
  const bool generic_only = genv && atoi(genv) != 0;
  auto launch_q = [&]() {
    const dim3 grid(h->sm_count * 8), block(256);
    switch (generic_only ? 0 : K) {
      case 1: k_init_quantities_t<1><<<grid, block, 0, h->stream>>>(a); break;
      case 2: k_init_quantities_t<2><<<grid, block, 0, h->stream>>>(a); break;
      case 3: k_init_quantities_t<3><<<grid, block, 0, h->stream>>>(a); break;
      case 4: k_init_quantities_t<4><<<grid, block, 0, h->stream>>>(a); break;
      case 5: k_init_quantities_t<5><<<grid, block, 0, h->stream>>>(a); break;
      default: k_init_quantities<<<grid, block, 0, h->stream>>>(a);
    }
    h->launches++;
  };
  auto run_phase = [&](int phase, int nq) -> int {
    std::vector<unsigned long long> hmax(nq);
    std::vector<long long> hsum(2 * nq);
    a.phase = phase;
    a.mode = 0;
    CK(h, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long) * NQ_MAX, h->stream));
    CK(h, cudaMemsetAsync(d_sum, 0, sizeof(long long) * NQ_MAX * 2, h->stream));
    launch_q();
    CK(h, cudaGetLastError());
    if (h->comm) {
      int r = g_nccl.AllReduce(d_max, d_max, (size_t)nq, 5 /*ncclUint64*/, 2 /*ncclMax*/, h->comm, h->stream);
      if (r != 0) return fail(h, CLV_ERR_COMM, "ncclAllReduce(max) failed (%d)", r);
    }
    CK(h, cudaMemcpyAsync(hmax.data(), d_max, sizeof(unsigned long long) * nq, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    std::vector<int> bits(nq);
    for (int q = 0; q < nq; ++q) {
      double m;
      std::memcpy(&m, &hmax[q], sizeof m);
      if (!std::isfinite(m)) return fail(h, CLV_ERR_NUMERIC, "non-finite value in an initialisation statistic");
      bits[q] = fx_bits_host(m);
      a.scale[q] = std::ldexp(1.0, bits[q]);
      if (phase == 0 && q >= 3) max_abs_x = std::max(max_abs_x, std::sqrt(m));   // diagonal pairs give max |X_k|^2
    }
    a.mode = 1;
    launch_q();
    CK(h, cudaGetLastError());
    if (h->comm) {
      int r = g_nccl.AllReduce(d_sum, d_sum, (size_t)nq * 2, NCCL_INT64, NCCL_SUM, h->comm, h->stream);
      if (r != 0) return fail(h, CLV_ERR_COMM, "ncclAllReduce(sum) failed (%d)", r);
    }
    CK(h, cudaMemcpyAsync(hsum.data(), d_sum, sizeof(long long) * nq * 2, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < nq; ++q) {
      __int128 t = ((__int128)hsum[2 * q + 1] << 32) + (__int128)hsum[2 * q];
      tot[q] = i128_scaled(t, bits[q]);
    }
    return 0;
  };
  rc = run_phase(0, nqA);
  if (!rc) {
    const double mean_x = tot[0] / n, mean_t = tot[1] / n;
    out->lam_init = mean_x / mean_t;
    out->mean_log_s = (D == 3) ? tot[2] / n : 0.0;
    xtx.assign((size_t)K * K, 0.0);
    int q = 3;
    for (int i = 0; i < K; ++i)
      for (int j = i; j < K; ++j, ++q) xtx[i * K + j] = xtx[j * K + i] = tot[q];
    out->max_abs_x = max_abs_x;
    if (!(out->lam_init > 0.0) || !std::isfinite(out->lam_init))
      rc = fail(h, CLV_ERR_NUMERIC, "lam_init = mean(x)/mean(t) is not positive and finite (no repeat purchases?)");
  }
  if (!rc) {
    a.lam_init = out->lam_init;
    a.mean_log_s = out->mean_log_s;
    rc = run_phase(1, 2);
  }
  if (!rc) {
    out->mean_mu_init = tot[0] / n;
    out->omega2 = (D == 3) ? tot[1] / (n - 1.0) : 1.0;
    out->xtx = xtx.data();
  }
  dfree(d_max);
  dfree(d_sum);
  return rc;
}

int clv_init_state(clv_sampler* h, const clv_init_stats* st) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->have_data || !h->have_hyper) return fail(h, CLV_ERR_STATE, "clv_init_state: call clv_set_data and clv_set_hyper first");
  clv_init_stats computed{};
  std::vector<double> xtx_buf;
  if (!st) {
    CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
    if (int r = device_init_stats(h, &computed, xtx_buf)) return r;
    st = &computed;
    h->last_stats = computed;
    h->last_xtx = xtx_buf;
  }
  if (!st->xtx) return fail(h, CLV_ERR_ARG, "clv_init_state: stats without xtx");
  if (!(st->lam_init > 0.0) || !std::isfinite(st->lam_init) || !(st->mean_mu_init > 0.0))
    return fail(h, CLV_ERR_NUMERIC, "clv_init_state: lam_init / mean_mu_init must be positive and finite");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const int K = h->K, D = h->D;
  ModelConst& mc = h->h_mc;
  std::memset(&mc, 0, sizeof mc);
  mc.D = D; mc.K = K; mc.S = h->S; mc.compat = h->cfg.compat;
  mc.N = h->N; mc.N_global = h->cfg.n_global; mc.gid_offset = h->cfg.gid_offset;
  // prior mean with the data-dependent intercept row (bi:373-374, tri:497-499)
  std::vector<double> B0(h->beta0);
  B0[0] = std::log(st->lam_init);
  B0[1] = std::log(st->mean_mu_init);
  if (D == 3) B0[2] = st->mean_log_s;
  for (int d = 0; d < D; ++d) mc.center[d] = B0[d];
  // fixed point: |y - c| <= 140 (clip at +-70, bi:323-324), |x_k| <= max_abs_x
  double mx = std::max(1.0, st->max_abs_x);
  double bound = 140.0 * std::max(140.0, mx);
  int bits = 62 - ceil_log2((double)h->cfg.n_global + 1.0) - ceil_log2(bound);
  bits = std::max(4, std::min(bits, 50));
  mc.fx_scale = std::ldexp(1.0, bits); mc.fx_inv = std::ldexp(1.0, -bits);
  int lbits = std::max(4, std::min(62 - ceil_log2((double)h->cfg.n_global + 1.0) - 20, 40));
  mc.ll_scale = std::ldexp(1.0, lbits); mc.ll_inv = std::ldexp(1.0, -lbits);
  // V = (X'X + A0)^-1, chol(V)
  std::vector<double> G(K * K);
  for (int t = 0; t < K * K; ++t) G[t] = st->xtx[t] + h->A0[t];
  if (!spd_inverse(G.data(), mc.V, K)) return fail(h, CLV_ERR_NUMERIC, "X'X + A0 is not positive definite");
  if (!chol_host(mc.V, mc.LV, K)) return fail(h, CLV_ERR_NUMERIC, "(X'X + A0)^-1 is not positive definite");
  // centred prior mean B0c = B0 - e0 c'
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < D; ++d) mc.B0c[k * D + d] = B0[k * D + d] - (k == 0 ? mc.center[d] : 0.0);
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < D; ++d) {
      double s = 0.0;
      for (int m = 0; m < K; ++m) s += h->A0[k * K + m] * mc.B0c[m * D + d];
      mc.A0B0c[k * D + d] = s;
    }
  for (int d = 0; d < D; ++d)
    for (int e = 0; e < D; ++e) {
      double s = h->gamma0[d * D + e];
      for (int k = 0; k < K; ++k) s += mc.B0c[k * D + d] * mc.A0B0c[k * D + e];
      mc.Q0[d * D + e] = s;
    }
  mc.nu_n = h->nu0 + (double)h->cfg.n_global;
  mc.omega2 = (D == 3) ? st->omega2 : 1.0;
  if (D == 3 && !(mc.omega2 > 0.0)) return fail(h, CLV_ERR_NUMERIC, "omega2 = var(log_s) must be positive");
  CK(h, cudaMemcpyAsync(h->d_mc, &mc, sizeof mc, cudaMemcpyHostToDevice, h->stream));
  // initial level-1 state (bi:368-370, tri:489-493) and level-2 placeholders (bi:379, tri:504)
  const size_t N = (size_t)h->N;
  std::vector<ChainParams> cps(h->chains);
  for (int c = 0; c < h->chains; ++c) {
    std::memset(&cps[c], 0, sizeof(ChainParams));
    for (int t = 0; t < K * D; ++t) cps[c].beta[t] = B0[t];
    for (int t = 0; t < D * D; ++t) cps[c].Sigma[t] = h->gamma0[t];
  }
  k_init_state<<<h->sm_count * 8, 256, 0, h->stream>>>(h->d_tx, (long long)N, h->chains, D, st->lam_init, h->d_ll, h->d_lm, h->d_le);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaMemsetAsync(h->d_z, 0, sizeof(double) * N * h->chains, h->stream));
  CK(h, cudaMemsetAsync(h->d_tau, 0, sizeof(double) * N * h->chains, h->stream));
  CK(h, cudaMemcpyAsync(h->d_params, cps.data(), sizeof(ChainParams) * h->chains, cudaMemcpyHostToDevice, h->stream));
  if (D == 2) k_derive_params<2><<<(h->chains + 63) / 64, 64, 0, h->stream>>>(h->d_mc, h->d_params, h->chains);
  else k_derive_params<3><<<(h->chains + 63) / 64, 64, 0, h->stream>>>(h->d_mc, h->d_params, h->chains);
  h->launches++;
  CK(h, cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream));
  h->inited = true;
  { std::lock_guard<std::mutex> lock(g_cache_mutex); h->init_epoch = ++g_epoch; }
  if (int r = recompute_stats(h)) return r;
  CK(h, cudaStreamSynchronize(h->stream));
  h->sweeps_done = 0;
  h->resident_draws = 0;
  return CLV_OK;
}

int clv_get_init_stats(clv_sampler* h, clv_init_stats* out, double* xtx_out) {
  if (!h || !out) return fail(h, CLV_ERR_ARG, "null argument");
  if (h->last_xtx.empty()) return fail(h, CLV_ERR_STATE, "no device-computed statistics: call clv_init_state(h, NULL) first");
  *out = h->last_stats;
  out->xtx = nullptr;
  if (xtx_out) std::memcpy(xtx_out, h->last_xtx.data(), sizeof(double) * h->K * h->K);
  return CLV_OK;
}

int clv_comm_unique_id(void* out128) {
  if (!out128) return fail(nullptr, CLV_ERR_ARG, "null argument");
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, CLV_ERR_COMM, "%s", err.c_str());
  NcclUid id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(nullptr, CLV_ERR_COMM, "ncclGetUniqueId failed (%d)", r);
  std::memcpy(out128, &id, sizeof id);
  return CLV_OK;
}

int clv_comm_init(clv_sampler* h, const void* unique_id128, int rank, int world) {
  if (!h || !unique_id128) return fail(h, CLV_ERR_ARG, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(h, CLV_ERR_ARG, "bad rank/world");
  if (h->cfg.sweep_mode == CLV_SWEEP_PERSISTENT) return fail(h, CLV_ERR_ARG, "persistent sweep mode is single-shard only");
  std::string err;
  if (!g_nccl.load(err)) return fail(h, CLV_ERR_COMM, "%s", err.c_str());
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  h->world = world; h->rank = rank;
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (auto& c : g_comms)
    if (c.device == h->cfg.device && c.rank == rank && c.world == world) { h->comm = c.comm; return CLV_OK; }
  NcclUid id;
  std::memcpy(&id, unique_id128, sizeof id);
  int r = g_nccl.CommInitRank(&h->comm, world, id, rank);
  if (r != 0) return fail(h, CLV_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  g_comms.push_back({h->cfg.device, rank, world, h->comm});
  return CLV_OK;
}

static CachedMailbox* find_mailbox(const clv_sampler* h, int rank, int world) {
  for (auto& m : g_mailboxes)
    if (m.device == h->cfg.device && m.chains == h->chains && (world == 0 || (m.rank == rank && m.world == world))) return &m;
  return nullptr;
}

int clv_p2p_export(clv_sampler* h, void* handle64) {
  if (!h || !handle64) return fail(h, CLV_ERR_ARG, "null argument");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CachedMailbox* m = find_mailbox(h, 0, 0);
  if (!m) {
    CachedMailbox nm{};
    nm.device = h->cfg.device; nm.chains = h->chains; nm.rank = -1; nm.world = 0; nm.connected = false;
    // [2 parities][P2P_MAX_WORLD senders][chains][2 * NSTAT_MAX] 64-bit words (32 bits of payload + 32-bit tag each)
    nm.bytes = sizeof(unsigned long long) * 2 * P2P_MAX_WORLD * (size_t)h->chains * 2 * NSTAT_MAX;
    CK(h, cudaMalloc(&nm.base, nm.bytes));
    CK(h, cudaMemset(nm.base, 0, nm.bytes));
    g_mailboxes.push_back(nm);
    m = &g_mailboxes.back();
  }
  h->d_mailbox = m->base; h->mailbox_bytes = m->bytes;
  cudaIpcMemHandle_t hd;
  CK(h, cudaIpcGetMemHandle(&hd, h->d_mailbox));
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &hd, 64);
  return CLV_OK;
}

/* 1 when this process already holds a connected mailbox set for (device, rank, world, chains): clv_p2p_connect may then
 * be called with handles == NULL (no handle exchange needed). */
int clv_p2p_is_cached(clv_sampler* h, int rank, int world) {
  if (!h) return 0;
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CachedMailbox* m = find_mailbox(h, rank, world);
  return (m && m->connected) ? 1 : 0;
}

int clv_p2p_connect(clv_sampler* h, const void* handles, int rank, int world) {
  if (!h) return fail(h, CLV_ERR_ARG, "null argument");
  if (world < 2 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return fail(h, CLV_ERR_ARG, "p2p: world must be in [2, %d]", P2P_MAX_WORLD);
  if (h->cfg.sweep_mode == CLV_SWEEP_PERSISTENT) return fail(h, CLV_ERR_ARG, "persistent sweep mode is single-shard only");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CachedMailbox* m = find_mailbox(h, rank, world);
  if (!(m && m->connected)) {
    m = find_mailbox(h, 0, 0);
    if (!m || !handles) return fail(h, CLV_ERR_STATE, "clv_p2p_connect: call clv_p2p_export first and pass the gathered handles");
    for (int r = 0; r < world; ++r) {
      void* base = m->base;
      if (r != rank) {
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, (const char*)handles + 64 * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(h, CLV_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      }
      m->peer_base[r] = base;
    }
    m->rank = rank; m->world = world; m->connected = true;
  }
  // a fresh start for this handle's tags; the caller synchronises the ranks (host barrier) between this call and the
  // first sweep, so no peer word can arrive before the memset
  CK(h, cudaMemset(m->base, 0, m->bytes));
  CK(h, cudaDeviceSynchronize());
  h->rank = rank; h->world = world;
  h->d_mailbox = m->base; h->mailbox_bytes = m->bytes;
  for (int r = 0; r < world; ++r) h->peer_mail[r] = (unsigned long long*)m->peer_base[r];
  h->p2p = true;
  return CLV_OK;
}

int64_t clv_sweeps_done(const clv_sampler* h) { return h ? h->sweeps_done : -1; }
int64_t clv_kernel_launches(const clv_sampler* h) { return h ? h->launches : -1; }

int clv_set_sweeps_done(clv_sampler* h, int64_t n) {
  if (!h || n < 0 || n > 0xFFFFFFF0ll) return fail(h, CLV_ERR_ARG, "clv_set_sweeps_done: bad argument");
  h->sweeps_done = n;
  return CLV_OK;
}

int clv_set_timing(clv_sampler* h, int on) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  CK(h, cudaStreamSynchronize(h->stream));
  collect_timing(h);
  h->timing = on != 0;
  h->t_sweep_ms = h->t_l2_ms = 0; h->t_n = 0;
  return CLV_OK;
}

int clv_kernel_time_ms(clv_sampler* h, double* sweep_ms, double* l2_ms, int64_t* n) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  CK(h, cudaStreamSynchronize(h->stream));
  collect_timing(h);
  if (sweep_ms) *sweep_ms = h->t_sweep_ms;
  if (l2_ms) *l2_ms = h->t_l2_ms;
  if (n) *n = h->t_n;
  h->t_sweep_ms = h->t_l2_ms = 0; h->t_n = 0;
  return CLV_OK;
}


namespace {

bool want_persistent(const clv_sampler* h) {
  if (h->comm || h->p2p || h->timing || h->fc_enabled || h->persist_grid_x <= 0 || h->cfg.rng_mode == CLV_RNG_INJECTED) return false;
  // AUTO = stream: with programmatic dependent launch the two-kernel path is at least as fast as the cooperative kernel
  // at every size measured (14.2 vs 16.4 us per sweep at 4 x 2 357 customers, 18.2 vs 20.4 at 2 x 23 570, 213 vs 261 at
  // 1.25 M; profiles/r02_kernel_ab.txt): the grid barrier costs more than the launch boundary that PDL hides.
  return h->cfg.sweep_mode == CLV_SWEEP_PERSISTENT;
}

struct RunCtx {                 // draw bookkeeping of the current clv_run (zeros for clv_advance)
  long long burnin = 0, thin = 1, n_draws = 1, chunk_base = 0, cap = 1, step0 = 1;
  double* draws = nullptr;
  bool keep_any = false;
};

int run_stream_segment(clv_sampler* h, const RunCtx& rc, long long n, bool store_zt_last);

// n sweeps fused in one cooperative launch (grid barrier per sweep).
int launch_persistent(clv_sampler* h, const RunCtx& rc, long long n, bool store_zt_last) {
  const size_t slot_bytes = sizeof(unsigned long long) * h->chains * NSTAT_MAX;
  const uint32_t first = (uint32_t)(h->sweeps_done + 1), last = (uint32_t)(h->sweeps_done + n);
  CK(h, cudaMemsetAsync(h->d_acc3, 0, 3 * slot_bytes, h->stream));
  if (h->D == 2)   // statistics of the current state feed the first level-2 draw (bi:393)
    CK(h, cudaMemcpyAsync(h->d_acc3 + (size_t)((first + 2u) % 3u) * h->chains * NSTAT_MAX, h->d_acc, slot_bytes,
                          cudaMemcpyDeviceToDevice, h->stream));
  PersistArgs pa{};
  pa.sw = base_args(h);
  pa.sw.loglik_stride = rc.n_draws;
  pa.sw.draws = rc.draws;
  pa.sw.chunk_cap = rc.cap;
  pa.params = h->d_params;
  pa.acc3 = h->d_acc3;
  pa.level2_draws = h->d_level2;
  pa.n_draws = rc.n_draws;
  pa.first_sweep = first;
  pa.n_sweeps = (uint32_t)n;
  pa.run_step0 = rc.step0;
  pa.burnin = rc.keep_any ? rc.burnin : (1ll << 62);
  pa.thin = rc.thin;
  pa.chunk_base = rc.chunk_base;
  pa.store_zt_last = store_zt_last ? 1 : 0;
  pa.error_flag = h->d_err;
  pa.barrier = h->d_barrier;
  dim3 grid(h->persist_grid_x, h->chains), block(SWEEP_THREADS);
  void* args[] = {&pa};
  const void* fn;
  const bool strict = h->cfg.rng_mode == CLV_RNG_PHILOX_STRICT;
  if (h->D == 2) fn = strict ? (const void*)k_persistent<2, MODE_STRICT> : (const void*)k_persistent<2, MODE_FAST>;
  else fn = strict ? (const void*)k_persistent<3, MODE_STRICT> : (const void*)k_persistent<3, MODE_FAST>;
  {
    cudaError_t ec = cudaLaunchCooperativeKernel(fn, grid, block, args, h->stats_smem, h->stream);
    if (ec == cudaErrorCooperativeLaunchTooLarge) {       // occupancy changed under us: the two-kernel path gives the same chain
      cudaGetLastError();
      h->persist_grid_x = 0;
      return run_stream_segment(h, rc, n, store_zt_last);
    }
    CK(h, ec);
  }
  h->launches++;
  if (h->D == 2)
    CK(h, cudaMemcpyAsync(h->d_acc, h->d_acc3 + (size_t)(last % 3u) * h->chains * NSTAT_MAX, slot_bytes,
                          cudaMemcpyDeviceToDevice, h->stream));
  else
    CK(h, cudaMemsetAsync(h->d_acc, 0, slot_bytes, h->stream));
  h->sweeps_done += n;
  return 0;
}

// Sweeps [rc.step0, rc.step0 + n) of the current run, stream mode: two kernels per sweep.
int run_stream_segment(clv_sampler* h, const RunCtx& rc, long long n, bool store_zt_last) {
  for (long long it = 0; it < n; ++it) {
    const long long step = rc.step0 + it;
    SweepArgs a = base_args(h);
    Level2Args l2 = base_l2(h);
    a.sweep = l2.sweep = (uint32_t)(h->sweeps_done + 1);
    l2.n_draws = rc.n_draws;
    a.loglik_stride = rc.n_draws;
    a.fc = (rc.keep_any && h->fc_enabled) ? h->d_fc : nullptr;
    const bool kept = rc.keep_any && step > rc.burnin && (step - 1 - rc.burnin) % rc.thin == 0;      // bi:402
    if (kept) {
      const long long draw = (step - 1 - rc.burnin) / rc.thin;
      a.draw_index = l2.draw_index = draw;
      a.slot = 0;
      if (rc.draws) {
        a.draws = rc.draws;
        a.chunk_cap = rc.cap;
        a.slot = draw - rc.chunk_base;
      }
    }
    if (store_zt_last && it + 1 == n) a.store_zt = 1;
    if (int r = enqueue_sweep(h, a, l2, h->cfg.rng_mode)) return r;
    if (h->timing && (it & 1023) == 1023) { CK(h, cudaStreamSynchronize(h->stream)); collect_timing(h); }
  }
  return 0;
}

int run_segment(clv_sampler* h, const RunCtx& rc, long long n, bool store_zt_last) {
  if (n <= 0) return 0;
  return want_persistent(h) ? launch_persistent(h, rc, n, store_zt_last) : run_stream_segment(h, rc, n, store_zt_last);
}

}  // namespace

int clv_advance(clv_sampler* h, int64_t n_sweeps, int sync) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "clv_advance: call clv_init_state first");
  if (h->cfg.rng_mode == CLV_RNG_INJECTED) return fail(h, CLV_ERR_ARG, "handle is in injected-RNG mode; use clv_sweep_injected");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  RunCtx rc;
  // cooperative launches are bounded so that a runaway kernel cannot outlive the watchdog of a shared box
  for (int64_t done = 0; done < n_sweeps;) {
    const long long n = std::min<long long>(n_sweeps - done, 20000);
    // a synchronous call leaves z / tau of its last sweep in the state arrays (clv_get_state)
    if (int r = run_segment(h, rc, n, sync != 0 && done + n == n_sweeps)) return r;
    done += n;
  }
  if (sync) return check_device_error(h);
  return CLV_OK;
}

int clv_advance_timed(clv_sampler* h, int64_t n_sweeps, double* elapsed_ms) {
  if (!h || !elapsed_ms) return fail(h, CLV_ERR_ARG, "null argument");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "clv_advance_timed: call clv_init_state first");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  cudaEvent_t e0, e1;
  CK(h, cudaEventCreate(&e0));
  CK(h, cudaEventCreate(&e1));
  CK(h, cudaStreamSynchronize(h->stream));
  CK(h, cudaEventRecord(e0, h->stream));
  int r = clv_advance(h, n_sweeps, 0);
  if (r == 0) {
    cudaEventRecord(e1, h->stream);
    r = check_device_error(h);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *elapsed_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return r;
}

static int ensure_run_buffers(clv_sampler* h, long long n_draws, bool want_level1, long long* chunk_cap) {
  const long long C = h->chains;
  if (h->loglik_cap < C * n_draws) {
    if (h->d_loglik) dfree(h->d_loglik);
    h->d_loglik = nullptr;
    CK(h, dmalloc(&h->d_loglik, (size_t)(C * n_draws)));
    h->loglik_cap = C * n_draws;
  }
  if (h->level2_cap < C * n_draws * h->P) {
    if (h->d_level2) dfree(h->d_level2);
    h->d_level2 = nullptr;
    CK(h, dmalloc(&h->d_level2, (size_t)(C * n_draws * h->P)));
    h->level2_cap = C * n_draws * h->P;
  }
  CK(h, cudaMemsetAsync(h->d_loglik, 0, sizeof(long long) * C * n_draws, h->stream));
  *chunk_cap = 0;
  if (!want_level1) return 0;
  const long long per_draw = C * h->N * h->ncol * (long long)sizeof(double);
  const size_t free_b = available_bytes(h->cfg.device);
  long long have = h->draws_cap_bytes[0] + h->draws_cap_bytes[1];
  long long budget = (long long)((double)(free_b + have) * 0.70);
  if (const char* env = getenv("CLV_DRAW_BUFFER_BYTES")) budget = std::min(budget, atoll(env));
  long long cap = n_draws;
  int nbuf = 1;
  if (per_draw * n_draws > budget) {
    nbuf = 2;
    cap = std::max<long long>(1, budget / (2 * per_draw));
  }
  for (int b = 0; b < 2; ++b) {
    long long need = (b < nbuf) ? cap * per_draw : 0;
    if (h->draws_cap_bytes[b] < need) {
      if (h->d_draws[b]) dfree(h->d_draws[b]);
      h->d_draws[b] = nullptr; h->draws_cap_bytes[b] = 0;
      cudaError_t e = pool_malloc((void**)&h->d_draws[b], (size_t)need);
      if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "cannot allocate %lld bytes for the draw buffer: %s", need, cudaGetErrorString(e));
      h->draws_cap_bytes[b] = need;
    }
  }
  *chunk_cap = cap;
  return 0;
}

static int run_impl(clv_sampler* h, int64_t burnin, int64_t mcmc, int64_t thin, double* level1, bool resident_only,
                    double* level2, double* loglik, clv_progress_cb cb, void* user, int64_t trace) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "clv_run: call clv_init_state first");
  if (h->cfg.rng_mode == CLV_RNG_INJECTED) return fail(h, CLV_ERR_ARG, "handle is in injected-RNG mode; use clv_sweep_injected");
  if (burnin < 0 || mcmc < 1 || thin < 1) return fail(h, CLV_ERR_ARG, "clv_run: need burnin >= 0, mcmc >= 1, thin >= 1");
  if (!level2) return fail(h, CLV_ERR_ARG, "clv_run: level2 output is required");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const long long n_draws = (mcmc - 1) / thin + 1;      // bi:360
  const long long C = h->chains, N = h->N, nc = h->ncol;
  long long cap = 0;
  const bool store = level1 != nullptr || resident_only;
  if (int r = ensure_run_buffers(h, n_draws, store, &cap)) return r;
  if (resident_only && cap < n_draws)
    return fail(h, CLV_ERR_ARG, "clv_run_resident: %lld draws do not fit the device draw buffer (%lld fit)", n_draws, cap);
  if (h->fc_enabled) {        // fused forecast: fresh per-(chain, customer) sums for this run
    CK(h, cudaMemsetAsync(h->d_fc_sx, 0, sizeof(unsigned long long) * C * N, h->stream));
    CK(h, cudaMemsetAsync(h->d_fc_sz, 0, sizeof(unsigned int) * C * N, h->stream));
    h->fc_draws = n_draws;
  }
  const long long total = burnin + mcmc;
  // the caller's level-1 array gets its first touch ahead of the copies, while the sweeps run (CLV_FIRST_TOUCH=0: off).
  // Only when there is time for it: a burn-in before the first kept draw, or an array of a GB and more (a short run without
  // burn-in -- the bench's end-to-end call: 320 MB kept at the first sweep -- is 1 % faster without the extra threads).
  FirstToucher toucher;
  if (level1) {
    const size_t chain_bytes = (size_t)n_draws * N * nc * sizeof(double);
    if ((burnin > 0 || chain_bytes * (size_t)C >= ((size_t)1 << 30)) && first_touch_wanted(level1, chain_bytes * (size_t)C)) {
      const size_t piece = 8u << 20;
      std::vector<std::pair<char*, size_t>> ranges;      // in the order of the flushes: piece by piece, every chain
      for (size_t off = 0; off < chain_bytes; off += piece)
        for (long long c = 0; c < C; ++c)
          ranges.emplace_back((char*)level1 + (size_t)c * chain_bytes + off, std::min(piece, chain_bytes - off));
      toucher.start(std::move(ranges), first_touch_threads());
    }
  }
  int buf = 0;
  long long chunk_base = 0;
  bool used[2] = {false, false};
  // Kept draws go to the host in pieces of about FLUSH_BYTES (all chains), not chunk by chunk: a piece is recorded
  // "ready" right after the segment that completed it, and its device->host copies are issued only after the NEXT
  // segment has been enqueued -- the copy keeps the calling thread busy (staging ring + parallel memcpy into the
  // caller's pageable array), and this order lets the GPU sample meanwhile.  When the run does not fit one device
  // chunk, two chunk buffers alternate; a buffer is rewritten only after the copies of its last piece.
  const long long per_draw_bytes = C * N * nc * (long long)sizeof(double);
  long long flush_bytes = 512ll << 20;
  if (const char* env = getenv("CLV_FLUSH_BYTES")) flush_bytes = std::max(1ll, atoll(env));
  const long long flush_every = std::max<long long>(1, flush_bytes / std::max<long long>(1, per_draw_bytes));
  long long flushed = 0;                        // draws already handed to the copy path
  struct PendingPiece { bool active = false; int buf = 0; long long first = 0, count = 0, chunk_base = 0; } pend;
  auto issue_copies = [&]() -> int {
    if (!pend.active) return 0;
    pend.active = false;
    // draws [first, first+count) of every chain -> host, on the copy stream
    CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_chunk_ready[pend.buf], 0));
    for (long long c = 0; c < C; ++c)
      CK(h, copy_to_host_staged(level1 + ((size_t)c * n_draws + pend.first) * N * nc,
                                h->d_draws[pend.buf] + ((size_t)c * cap + (pend.first - pend.chunk_base)) * N * nc,
                                (size_t)pend.count * N * nc * sizeof(double), h->copy_stream));
    CK(h, cudaEventRecord(h->ev_copy_done[pend.buf], h->copy_stream));
    return 0;
  };
  // The run is cut into segments that end where the host has to act: a piece of draws to flush, a full chunk, the
  // last kept draw (so that its copy overlaps the trailing non-kept sweeps), a progress callback, or the end.  A
  // segment is one cooperative launch (persistent mode) or 2 launches per sweep.
  long long step = 1;
  while (step <= total) {
    long long seg_end = total;
    if (store && flushed < n_draws) {
      long long last_draw = std::min(n_draws - 1, chunk_base + cap - 1);
      if (level1) last_draw = std::min(last_draw, flushed + flush_every - 1);
      seg_end = std::min(seg_end, burnin + 1 + last_draw * thin);
      if (!level1 && last_draw == n_draws - 1) seg_end = total;      // resident draws: nothing to copy, one segment
    }
    if (trace > 0) seg_end = std::min(seg_end, ((step + trace - 1) / trace) * trace);
    seg_end = std::min(seg_end, step + 19999);
    RunCtx rc;
    rc.burnin = burnin; rc.thin = thin; rc.n_draws = n_draws; rc.chunk_base = chunk_base; rc.cap = store ? cap : 1;
    rc.step0 = step; rc.draws = store ? h->d_draws[buf] : nullptr; rc.keep_any = true;
    if (int r = run_segment(h, rc, seg_end - step + 1, seg_end == total)) return r;
    step = seg_end + 1;
    if (int r = issue_copies()) return r;          // the previous piece, now overlapping the segment just enqueued
    if (level1 && flushed < n_draws) {
      // draws kept so far: those with burnin + 1 + d*thin <= seg_end
      const long long kept_upto = seg_end > burnin ? std::min(n_draws, (seg_end - burnin - 1) / thin + 1) : 0;
      const bool chunk_full = kept_upto - chunk_base == cap;
      if (kept_upto > flushed && (kept_upto - flushed >= flush_every || chunk_full || kept_upto == n_draws)) {
        CK(h, cudaEventRecord(h->ev_chunk_ready[buf], h->stream));
        pend.active = true; pend.buf = buf; pend.first = flushed; pend.count = kept_upto - flushed; pend.chunk_base = chunk_base;
        used[buf] = true;
        flushed = kept_upto;
        if (chunk_full && kept_upto != n_draws) {
          // the other buffer is written next: wait for ITS last copies (issued at the latest right after the
          // previous segment was enqueued, so the event is recorded)
          chunk_base = kept_upto;
          buf ^= 1;
          if (used[buf]) CK(h, cudaStreamWaitEvent(h->stream, h->ev_copy_done[buf], 0));
        }
      }
    }
    if (trace > 0 && seg_end % trace == 0) {                                   // bi:384-385
      if (int r = check_device_error(h)) return r;
      if (cb) cb(user, seg_end, total);
    }
  }
  if (int r = issue_copies()) return r;
  if (int r = check_device_error(h)) return r;
  CK(h, cudaMemcpyAsync(level2, h->d_level2, sizeof(double) * C * n_draws * h->P, cudaMemcpyDeviceToHost, h->stream));
  std::vector<long long> ll((size_t)(C * n_draws));
  CK(h, cudaMemcpyAsync(ll.data(), h->d_loglik, sizeof(long long) * C * n_draws, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  CK(h, cudaStreamSynchronize(h->copy_stream));
  if (loglik)
    for (size_t t = 0; t < ll.size(); ++t) loglik[t] = (double)ll[t] * h->h_mc.ll_inv;
  h->resident_draws = (store && cap >= n_draws) ? n_draws : 0;
  return CLV_OK;
}

int clv_run(clv_sampler* h, int64_t burnin, int64_t mcmc, int64_t thin, double* level1, double* level2,
            double* loglik, clv_progress_cb cb, void* user, int64_t trace) {
  return run_impl(h, burnin, mcmc, thin, level1, false, level2, loglik, cb, user, trace);
}

int clv_run_resident(clv_sampler* h, int64_t burnin, int64_t mcmc, int64_t thin, double* level2, double* loglik,
                     clv_progress_cb cb, void* user, int64_t trace) {
  return run_impl(h, burnin, mcmc, thin, nullptr, true, level2, loglik, cb, user, trace);
}

int clv_set_fused_forecast(clv_sampler* h, int enable, double T_star, uint64_t seed) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  if (int r = upload_rk()) return r;
  h->fc_enabled = enable != 0;
  h->fc_draws = 0;
  if (!h->fc_enabled) return CLV_OK;
  if (!(T_star >= 0.0)) return fail(h, CLV_ERR_ARG, "clv_set_fused_forecast: T_star must be >= 0");
  const size_t cn = (size_t)h->chains * (size_t)h->N;
  if (!h->d_fc) { CK(h, dmalloc(&h->d_fc, 1)); CK(h, dmalloc(&h->d_fc_sx, cn)); CK(h, dmalloc(&h->d_fc_sz, cn)); }
  h->fc_T_star = T_star; h->fc_seed = seed;
  FusedForecast f{};
  f.sum_x = h->d_fc_sx; f.sum_z = h->d_fc_sz; f.T_star = T_star; f.seed = seed; f.rk = round_keys(seed);
  CK(h, cudaMemcpyAsync(h->d_fc, &f, sizeof f, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(h->d_fc_sx, 0, sizeof(unsigned long long) * cn, h->stream));
  CK(h, cudaMemsetAsync(h->d_fc_sz, 0, sizeof(unsigned int) * cn, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));          // `f` is on this function's stack
  return CLV_OK;
}

int clv_fused_forecast_result(clv_sampler* h, double* mean_x_star, double* p_alive, int64_t* n_draws_total) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->fc_enabled || h->fc_draws <= 0) return fail(h, CLV_ERR_STATE, "no fused forecast: call clv_set_fused_forecast, then clv_run / clv_run_resident");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const long long C = h->chains, N = h->N;
  double *d_mx = nullptr, *d_pa = nullptr;
  CK(h, dmalloc(&d_mx, (size_t)N));
  CK(h, dmalloc(&d_pa, (size_t)N));
  k_fused_forecast_fold<<<h->sm_count * 4, 256, 0, h->stream>>>(h->d_fc_sx, h->d_fc_sz, N, (int)C, 1.0 / (double)(C * h->fc_draws), d_mx, d_pa);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && mean_x_star) e = cudaMemcpyAsync(mean_x_star, d_mx, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && p_alive) e = cudaMemcpyAsync(p_alive, d_pa, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dfree(d_mx); dfree(d_pa);
  if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_fused_forecast_result failed: %s", cudaGetErrorString(e));
  if (n_draws_total) *n_draws_total = C * h->fc_draws;
  return CLV_OK;
}

int clv_resident_draws(clv_sampler* h, const double** level1_dev, int64_t* n_draws) {
  if (!h || !level1_dev || !n_draws) return fail(h, CLV_ERR_ARG, "null argument");
  *level1_dev = h->resident_draws > 0 ? h->d_draws[0] : nullptr;
  *n_draws = h->resident_draws;
  return CLV_OK;
}

int clv_get_state(clv_sampler* h, int chain, double* ll, double* lm, double* le, double* z, double* tau,
                  double* beta, double* Sigma) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "state not initialised");
  if (chain < 0 || chain >= h->chains) return fail(h, CLV_ERR_ARG, "chain out of range");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  CK(h, cudaStreamSynchronize(h->stream));
  const size_t N = (size_t)h->N, o = (size_t)chain * N;
  if (ll) CK(h, cudaMemcpy(ll, h->d_ll + o, N * sizeof(double), cudaMemcpyDeviceToHost));
  if (lm) CK(h, cudaMemcpy(lm, h->d_lm + o, N * sizeof(double), cudaMemcpyDeviceToHost));
  if (le && h->D == 3) CK(h, cudaMemcpy(le, h->d_le + o, N * sizeof(double), cudaMemcpyDeviceToHost));
  if (z) CK(h, cudaMemcpy(z, h->d_z + o, N * sizeof(double), cudaMemcpyDeviceToHost));
  if (tau) CK(h, cudaMemcpy(tau, h->d_tau + o, N * sizeof(double), cudaMemcpyDeviceToHost));
  if (beta || Sigma) {
    ChainParams cp;
    CK(h, cudaMemcpy(&cp, h->d_params + chain, sizeof cp, cudaMemcpyDeviceToHost));
    if (beta) std::memcpy(beta, cp.beta, sizeof(double) * h->K * h->D);
    if (Sigma) std::memcpy(Sigma, cp.Sigma, sizeof(double) * h->D * h->D);
  }
  return CLV_OK;
}

int clv_set_state(clv_sampler* h, int chain, const double* ll, const double* lm, const double* le,
                  const double* beta, const double* Sigma) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "state not initialised");
  if (chain < 0 || chain >= h->chains) return fail(h, CLV_ERR_ARG, "chain out of range");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  CK(h, cudaStreamSynchronize(h->stream));
  const size_t N = (size_t)h->N, o = (size_t)chain * N;
  // the sweep kernel's exp is exact on the clip range of bi:323-324 and well defined up to +-700: reject anything else
  for (const double* v : {ll, lm})
    if (v)
      for (size_t i = 0; i < N; ++i)
        if (!(std::fabs(v[i]) <= 700.0)) return fail(h, CLV_ERR_ARG, "clv_set_state: log lambda / log mu must be finite and within +-700");
  if (ll) CK(h, cudaMemcpy(h->d_ll + o, ll, N * sizeof(double), cudaMemcpyHostToDevice));
  if (lm) CK(h, cudaMemcpy(h->d_lm + o, lm, N * sizeof(double), cudaMemcpyHostToDevice));
  if (le && h->D == 3) CK(h, cudaMemcpy(h->d_le + o, le, N * sizeof(double), cudaMemcpyHostToDevice));
  if (beta || Sigma) {
    ChainParams cp;
    CK(h, cudaMemcpy(&cp, h->d_params + chain, sizeof cp, cudaMemcpyDeviceToHost));
    if (beta) std::memcpy(cp.beta, beta, sizeof(double) * h->K * h->D);
    if (Sigma) std::memcpy(cp.Sigma, Sigma, sizeof(double) * h->D * h->D);
    CK(h, cudaMemcpy(h->d_params + chain, &cp, sizeof cp, cudaMemcpyHostToDevice));
    if (h->D == 2) k_derive_params<2><<<(h->chains + 63) / 64, 64, 0, h->stream>>>(h->d_mc, h->d_params, h->chains);
    else k_derive_params<3><<<(h->chains + 63) / 64, 64, 0, h->stream>>>(h->d_mc, h->d_params, h->chains);
    h->launches++;
  }
  if (int r = recompute_stats(h)) return r;
  CK(h, cudaStreamSynchronize(h->stream));
  return CLV_OK;
}

int clv_sweep_injected(clv_sampler* h, const clv_injected* v, int keep, double* level1, double* level2,
                       double* loglik) {
  if (!h || !v) return fail(h, CLV_ERR_ARG, "null argument");
  if (!h->inited) return fail(h, CLV_ERR_STATE, "clv_sweep_injected: call clv_init_state first");
  if (h->comm) return fail(h, CLV_ERR_ARG, "injected sweeps are single-shard only");
  const int D = h->D, K = h->K, S = h->S;
  if (!v->u_z || !v->e_tau || !v->u_tau || !v->iw_chi2 || !v->beta_norm || (S > 0 && (!v->t3_l || !v->t3_m || !v->u_acc)) ||
      (D == 3 && !v->n_eta) || !v->iw_norm)
    return fail(h, CLV_ERR_ARG, "clv_sweep_injected: missing variate array");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const long long C = h->chains, N = h->N;
  const long long n_cn = C * N, n_csn = C * S * N, ntril = D * (D - 1) / 2;
  const long long tot = 3 * n_cn + 3 * n_csn + (D == 3 ? n_cn : 0) + C * (ntril + D + D * K);
  if (h->inj_cap < tot) {
    if (h->d_inj) dfree(h->d_inj);
    h->d_inj = nullptr;
    CK(h, dmalloc(&h->d_inj, (size_t)tot));
    h->inj_cap = tot;
  }
  long long cap = 0;
  if (int r = ensure_run_buffers(h, 1, keep && level1, &cap)) return r;
  double* p = h->d_inj;
  auto up = [&](const double* src, long long n) -> const double* {
    double* dst = p;
    if (n > 0) cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    p += n;
    return dst;
  };
  SweepArgs a = base_args(h);
  Level2Args l2 = base_l2(h);
  a.u_z = up(v->u_z, n_cn); a.e_tau = up(v->e_tau, n_cn); a.u_tau = up(v->u_tau, n_cn);
  a.t3_l = up(v->t3_l, n_csn); a.t3_m = up(v->t3_m, n_csn); a.u_acc = up(v->u_acc, n_csn);
  a.n_eta = (D == 3) ? up(v->n_eta, n_cn) : nullptr;
  l2.iw_norm = up(v->iw_norm, C * ntril); l2.iw_chi2 = up(v->iw_chi2, C * D); l2.beta_norm = up(v->beta_norm, C * D * K);
  l2.injected = 1;
  CK(h, cudaGetLastError());
  a.sweep = l2.sweep = (uint32_t)(h->sweeps_done + 1);
  a.store_zt = 1;
  l2.n_draws = 1; a.loglik_stride = 1;
  if (keep) {
    a.slot = 0; a.draw_index = 0; l2.draw_index = 0;
    if (level1) { a.draws = h->d_draws[0]; a.chunk_cap = cap; }
  }
  if (int r = enqueue_sweep(h, a, l2, MODE_INJECT)) return r;
  if (int r = check_device_error(h)) return r;
  if (keep) {
    if (level1)
      for (long long c = 0; c < C; ++c)
        CK(h, cudaMemcpy(level1 + (size_t)c * N * h->ncol, h->d_draws[0] + (size_t)c * cap * N * h->ncol,
                         (size_t)N * h->ncol * sizeof(double), cudaMemcpyDeviceToHost));
    if (level2) CK(h, cudaMemcpy(level2, h->d_level2, sizeof(double) * C * h->P, cudaMemcpyDeviceToHost));
    if (loglik) {
      std::vector<long long> ll((size_t)C);
      CK(h, cudaMemcpy(ll.data(), h->d_loglik, sizeof(long long) * C, cudaMemcpyDeviceToHost));
      for (long long c = 0; c < C; ++c) loglik[c] = (double)ll[c] * h->h_mc.ll_inv;
    }
  }
  h->resident_draws = 0;
  return CLV_OK;
}

// ---- forecast ---------------------------------------------------------------------------------
static int launch_forecast(const clv_forecast_config* cfg, ForecastArgs a, bool inject, cudaStream_t st) {
  const long long N = cfg->n_customers;
  const long long npairs = ((a.draw_offset + a.n_draws - 1) >> 1) - (a.draw_offset >> 1) + 1;
  int gx = (int)std::min<long long>((N + 255) / 256, 65535);
  int gy = (int)std::max<long long>(1, std::min<long long>(npairs, 148ll * 8 * 4 / std::max(1, gx) + 1));
  gy = std::min(gy, 65535);
  dim3 grid(gx, gy);
  if (cfg->ncol == 4) {
    if (inject) k_forecast<4, true><<<grid, 256, 0, st>>>(a);
    else k_forecast<4, false><<<grid, 256, 0, st>>>(a);
  } else {
    if (inject) k_forecast<5, true><<<grid, 256, 0, st>>>(a);
    else k_forecast<5, false><<<grid, 256, 0, st>>>(a);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "forecast launch failed: %s", cudaGetErrorString(e));
  return 0;
}

static int check_fc(const clv_forecast_config* cfg) {
  if (!cfg) return fail(nullptr, CLV_ERR_ARG, "null forecast config");
  if (cfg->ncol != 4 && cfg->ncol != 5) return fail(nullptr, CLV_ERR_ARG, "ncol must be 4 or 5");
  if (cfg->n_draws_total < 1 || cfg->n_customers < 1) return fail(nullptr, CLV_ERR_ARG, "empty forecast");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CLV_ERR_ARG, "device out of range");
  return 0;
}

static int upload_rk() {
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lock(g_once_mutex);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
  double rk[RK_TABLE + 1];
  rk[0] = 0.0;
  for (int k = 1; k <= RK_TABLE; ++k) rk[k] = 1.0 / (double)k;
  cudaError_t e = cudaMemcpyToSymbol(c_rk, rk, sizeof rk);
  std::vector<float> rkf(RKF_TABLE, 0.0f);
  for (int k = 1; k < RKF_TABLE; ++k) rkf[k] = 1.0f / (float)k;
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_rkf, rkf.data(), sizeof(float) * RKF_TABLE);
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "cudaMemcpyToSymbol failed: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
}

int clv_forecast_dev(const clv_forecast_config* cfg, const double* level1_dev, const double* T_cal_dev,
                     int64_t* x_star_dev, double* spend_dev, void* stream) {
  if (int r = check_fc(cfg)) return r;
  CK(nullptr, cudaSetDevice(cfg->device)); t_alloc_stream = nullptr; ensure_pool(cfg->device);
  if (int r = upload_rk()) return r;
  ForecastArgs a{};
  a.level1 = level1_dev; a.T_cal = T_cal_dev; a.n_draws = cfg->n_draws_total; a.N = cfg->n_customers;
  a.T_star = cfg->T_star; a.sigma_s = cfg->sigma_s; a.seed = cfg->seed;
  a.gid_offset = cfg->gid_offset; a.draw_offset = cfg->draw_offset;
  a.x_out = (long long*)x_star_dev; a.spend_out = cfg->simulate_spend ? spend_dev : nullptr;
  a.rk = round_keys(cfg->seed);
  return launch_forecast(cfg, a, false, (cudaStream_t)stream);
}

// Host-buffer forecast: draws are streamed through the device in chunks on two streams (PCIe-bound).
static int forecast_host(const clv_forecast_config* cfg, const double* level1, const double* T_cal, const double* u,
                         const double* eps, int64_t n_eps, const int64_t* eps_offset, int64_t* x_star, double* spend,
                         bool inject) {
  if (int r = check_fc(cfg)) return r;
  if (!level1 || !T_cal || !x_star) return fail(nullptr, CLV_ERR_ARG, "clv_forecast: null buffer");
  if (inject && !u) return fail(nullptr, CLV_ERR_ARG, "clv_forecast_injected: u is required");
  const bool want_spend = cfg->simulate_spend && cfg->ncol == 5 && spend;
  if (inject && want_spend && (!eps || !eps_offset)) return fail(nullptr, CLV_ERR_ARG, "clv_forecast_injected: eps/eps_offset required for spend");
  CK(nullptr, cudaSetDevice(cfg->device)); t_alloc_stream = nullptr; ensure_pool(cfg->device);
  if (int r = upload_rk()) return r;
  const long long N = cfg->n_customers, nd = cfg->n_draws_total, nc = cfg->ncol;
  // fresh output arrays get their first touch (page allocation + zeroing) from host threads while the first rows travel
  FirstToucher toucher;
  {
    const size_t out_bytes = (size_t)nd * N * 8, piece = 8u << 20;
    std::vector<std::pair<char*, size_t>> ranges;
    if (first_touch_wanted(x_star, out_bytes))
      for (size_t off = 0; off < out_bytes; off += piece) {
        ranges.emplace_back((char*)x_star + off, std::min(piece, out_bytes - off));
        if (want_spend && !is_page_locked(spend)) ranges.emplace_back((char*)spend + off, std::min(piece, out_bytes - off));
      }
    if (!ranges.empty()) toucher.start(std::move(ranges), first_touch_threads());
  }
  const size_t free_b = available_bytes(cfg->device);
  const long long per_draw = N * (nc * 8 + 8 + (want_spend ? 8 : 0) + (inject ? 16 : 0));
  long long chunk = std::max<long long>(1, std::min<long long>(nd, (long long)(free_b * 0.35) / std::max<long long>(1, per_draw)));
  chunk = std::min<long long>(chunk, std::max<long long>(1, (1ll << 28) / std::max<long long>(1, N * nc * 8)));  // ~256 MB pieces pipeline well
  cudaStream_t st[2];
  double *d_l1[2] = {nullptr, nullptr}, *d_sp[2] = {nullptr, nullptr}, *d_u[2] = {nullptr, nullptr}, *d_T = nullptr, *d_eps = nullptr;
  long long *d_x[2] = {nullptr, nullptr}, *d_off[2] = {nullptr, nullptr};
  int rc = 0;
  auto cleanup = [&]() {
    for (int b = 0; b < 2; ++b) {
      if (d_l1[b]) dfree(d_l1[b]);
      if (d_sp[b]) dfree(d_sp[b]);
      if (d_u[b]) dfree(d_u[b]);
      if (d_x[b]) dfree(d_x[b]);
      if (d_off[b]) dfree(d_off[b]);
      cudaStreamDestroy(st[b]);
    }
    if (d_T) dfree(d_T);
    if (d_eps) dfree(d_eps);
  };
  st[0] = st[1] = nullptr;
  for (int b = 0; b < 2; ++b)
    if (cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking) != cudaSuccess) {
      if (st[0]) cudaStreamDestroy(st[0]);
      return fail(nullptr, CLV_ERR_CUDA, "clv_forecast: cannot create a stream");
    }
#define CKF(call) do { cudaError_t e3 = (call); if (e3 != cudaSuccess) { rc = fail(nullptr, CLV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e3)); cleanup(); return rc; } } while (0)
  CKF(dmalloc(&d_T, (size_t)N));
  CKF(cudaMemcpy(d_T, T_cal, (size_t)N * 8, cudaMemcpyHostToDevice));
  for (int b = 0; b < 2; ++b) {
    CKF(dmalloc(&d_l1[b], (size_t)(chunk * N * nc)));
    CKF(dmalloc(&d_x[b], (size_t)(chunk * N)));
    if (want_spend) CKF(dmalloc(&d_sp[b], (size_t)(chunk * N)));
    if (inject) { CKF(dmalloc(&d_u[b], (size_t)(chunk * N))); if (want_spend) CKF(dmalloc(&d_off[b], (size_t)(chunk * N))); }
  }
  if (inject && want_spend) {
    CKF(dmalloc(&d_eps, (size_t)n_eps));
    CKF(cudaMemcpy(d_eps, eps, (size_t)n_eps * 8, cudaMemcpyHostToDevice));
  }
  CKF(cudaStreamSynchronize(nullptr));          // pool allocations were ordered on the legacy stream; they are used on st[]
  int b = 0;
  for (long long d0 = 0; d0 < nd; d0 += chunk, b ^= 1) {
    const long long n = std::min(chunk, nd - d0);
    CKF(copy_to_device_staged(d_l1[b], level1 + (size_t)d0 * N * nc, (size_t)(n * N * nc) * 8, st[b]));
    if (inject) {
      CKF(cudaMemcpyAsync(d_u[b], u + (size_t)d0 * N, (size_t)(n * N) * 8, cudaMemcpyHostToDevice, st[b]));
      if (want_spend) CKF(cudaMemcpyAsync(d_off[b], eps_offset + (size_t)d0 * N, (size_t)(n * N) * 8, cudaMemcpyHostToDevice, st[b]));
    }
    clv_forecast_config c2 = *cfg;
    c2.n_draws_total = n;
    c2.draw_offset = cfg->draw_offset + d0;
    ForecastArgs a{};
    a.level1 = d_l1[b]; a.T_cal = d_T; a.n_draws = n; a.N = N; a.T_star = cfg->T_star; a.sigma_s = cfg->sigma_s;
    a.seed = cfg->seed; a.gid_offset = cfg->gid_offset; a.draw_offset = c2.draw_offset;
    a.u = d_u[b]; a.eps = d_eps; a.eps_offset = d_off[b];
    a.x_out = d_x[b]; a.spend_out = want_spend ? d_sp[b] : nullptr;
    a.rk = round_keys(cfg->seed);
    if ((rc = launch_forecast(&c2, a, inject, st[b]))) { cleanup(); return rc; }
    CKF(copy_to_host_staged(x_star + (size_t)d0 * N, d_x[b], (size_t)(n * N) * 8, st[b]));
    if (want_spend) CKF(copy_to_host_staged(spend + (size_t)d0 * N, d_sp[b], (size_t)(n * N) * 8, st[b]));
  }
  CKF(cudaStreamSynchronize(st[0]));
  CKF(cudaStreamSynchronize(st[1]));
#undef CKF
  cleanup();
  return CLV_OK;
}

int clv_forecast(const clv_forecast_config* cfg, const double* level1, const double* T_cal, int64_t* x_star, double* spend) {
  return forecast_host(cfg, level1, T_cal, nullptr, nullptr, 0, nullptr, x_star, spend, false);
}

int clv_forecast_injected(const clv_forecast_config* cfg, const double* level1, const double* T_cal, const double* u,
                          const double* eps, int64_t n_eps, const int64_t* eps_offset, int64_t* x_star, double* spend) {
  return forecast_host(cfg, level1, T_cal, u, eps, n_eps, eps_offset, x_star, spend, true);
}

int clv_forecast_resident(clv_sampler* h, double T_star, uint64_t seed, int64_t* x_star, double* mean_x_star, double* p_alive,
                          double* kernel_ms) {
  if (!h) return fail(nullptr, CLV_ERR_ARG, "null handle");
  if (h->resident_draws <= 0) return fail(h, CLV_ERR_STATE, "no resident draws: run clv_run with a level1 buffer that fits in one device chunk");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  if (int r = upload_rk()) return r;
  const long long C = h->chains, nd = h->resident_draws, N = h->N;
  double *d_mx = nullptr, *d_pa = nullptr;
  long long* d_x = nullptr;
  CK(h, dmalloc(&d_mx, (size_t)N));
  CK(h, dmalloc(&d_pa, (size_t)N));
  if (x_star) CK(h, dmalloc(&d_x, (size_t)(C * nd * N)));
  ForecastArgs a{};
  // resident layout [chains][nd][N][ncol] with cap == nd is exactly the chain-major (chains*nd, N, ncol) of bi:530-531
  a.level1 = h->d_draws[0]; a.T_cal = h->d_T; a.n_draws = C * nd; a.N = N; a.T_star = T_star; a.sigma_s = 0.5;
  a.seed = seed; a.gid_offset = h->cfg.gid_offset; a.draw_offset = 0; a.x_out = d_x; a.spend_out = nullptr;
  a.rk = round_keys(seed);
  int gx = (int)std::min<long long>((N + 255) / 256, 65535);
  // enough (customer, draw-range) threads to fill the GPU ~8 times over
  const long long npairs = (C * nd + 1) / 2;
  int gy = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(npairs, 64), (long long)h->sm_count * 8 * 8 / std::max(1, gx)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, h->stream);
  // Two passes: the main pass takes every cell through the quick path and lists the cells that need more (x* >= 8, the tie
  // zone, large means: ~3 %) in global memory; k_forecast_deferred works the list off on full warps.  Main pass: register-fed
  // (default: measured faster, profiles/r02_kernel_ab.txt) or TMA-fed (CLV_FC_KERNEL=tma; bivariate 32-byte rows only:
  // every bulk copy must be 16-byte aligned).  The list holds 1/8 of the cells; if that is ever too small the pass is
  // repeated with a list sized for the count.
  const char* fck = getenv("CLV_FC_KERNEL");
  const bool use_tma = fck && std::string(fck) == "tma";
  unsigned long long cap = (unsigned long long)(C * nd * N) / 8 + 65536;
  const bool cap_forced = getenv("CLV_FC_LIST_CAP") != nullptr;
  if (cap_forced) cap = (unsigned long long)std::max(1ll, atoll(getenv("CLV_FC_LIST_CAP")));            // test hook: force the retry
  // Chunks (CLV_FC_CHUNKS, default 1 = one main pass + one second pass): the draw pairs cut into ranges, each with its own
  // list; the second pass of a chunk runs on the copy stream (CLV_FC_SIDE_BLOCKS per SM, so that it fits beside the main
  // pass's blocks) while the main pass of the next chunk streams the rows.  Measured (1 M x 2 000 draws): 11.28 ms in one
  // piece, 11.10 - 11.46 ms in 4 - 32 chunks -- the main pass already fills 78 % of the issue slots and 76 % of the DRAM
  // cycles, so the second pass has nothing to hide in; kept as a knob (same result, tested).
  int chunks = 1;
  if (const char* e = getenv("CLV_FC_CHUNKS")) chunks = (int)std::max(1ll, std::min<long long>(atoll(e), std::min<long long>(npairs, 64)));
  if (use_tma) chunks = 1;
  chunks = std::max(1, std::min(chunks, 64));
  int side_blocks = 1;                            // blocks per SM of a second pass that runs beside a main pass
  if (const char* e = getenv("CLV_FC_SIDE_BLOCKS")) side_blocks = (int)std::max(1ll, std::min(8ll, atoll(e)));
  cudaError_t fe = cudaSuccess;
  for (int attempt = 0; attempt < 2 && fe == cudaSuccess; ++attempt) {
    if (attempt > 0) chunks = 1;                  // the retry after a list overflow: one list sized for the count
    const unsigned long long cap_c = std::max(1ull, cap / (unsigned long long)chunks + ((chunks > 1 && !cap_forced) ? 65536ull : 0ull));
    FcQueued* d_list = nullptr;
    unsigned long long* d_cnt = nullptr;
    fe = dmalloc(&d_list, (size_t)(cap_c * (unsigned long long)chunks));
    if (fe == cudaSuccess) fe = dmalloc(&d_cnt, (size_t)chunks);
    if (fe != cudaSuccess) { if (d_list) dfree(d_list); break; }
    cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * chunks, h->stream);
    const int gd = h->sm_count * 8;
    if (h->ncol == 4 && use_tma) {
      const size_t smem = (size_t)FC_STAGES * 2 * FC_TILE * 32 + (size_t)FC_WARPS * 96 * sizeof(FcQueued) + 2 * FC_STAGES * 8 + sizeof(FcCursor);
      const long long ntiles = (N + FC_TILE - 1) / FC_TILE;
      const int tx = (int)std::min<long long>(ntiles, 65535);
      const int ty = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(npairs, 64), (long long)h->sm_count * 4 * 4 / std::max(1, tx)));
      if (ty > 1 || attempt > 0) { cudaMemsetAsync(d_mx, 0, sizeof(double) * N, h->stream); cudaMemsetAsync(d_pa, 0, sizeof(double) * N, h->stream); }
      if (d_x) {
        cudaFuncSetAttribute(k_forecast_tma<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_forecast_tma<4, true><<<dim3(tx, ty), FC_TILE, smem, h->stream>>>(a, d_mx, d_pa, d_list, d_cnt, cap);
      } else {
        cudaFuncSetAttribute(k_forecast_tma<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_forecast_tma<4, false><<<dim3(tx, ty), FC_TILE, smem, h->stream>>>(a, d_mx, d_pa, d_list, d_cnt, cap);
      }
      h->launches++;
    }
    auto launch_deferred = [&](const ForecastArgs& aa, FcQueued* list, unsigned long long* cnt, int grid, cudaStream_t st) {
      if (h->ncol == 4) {
        if (d_x) k_forecast_deferred<4, true><<<grid, 256, 0, st>>>(aa, list, cnt, cap_c, d_mx);
        else k_forecast_deferred<4, false><<<grid, 256, 0, st>>>(aa, list, cnt, cap_c, d_mx);
      } else {
        if (d_x) k_forecast_deferred<5, true><<<grid, 256, 0, st>>>(aa, list, cnt, cap_c, d_mx);
        else k_forecast_deferred<5, false><<<grid, 256, 0, st>>>(aa, list, cnt, cap_c, d_mx);
      }
      h->launches++;
    };
    if (h->ncol == 4 && use_tma) {
      launch_deferred(a, d_list, d_cnt, gd, h->stream);
    } else {
      if (gy > 1 || attempt > 0 || chunks > 1) { cudaMemsetAsync(d_mx, 0, sizeof(double) * N, h->stream); cudaMemsetAsync(d_pa, 0, sizeof(double) * N, h->stream); }
      cudaEvent_t ev[64] = {};                      // chunks <= 64
      for (int c = 0; c < chunks; ++c) {
        ForecastArgs ac = a;
        if (chunks > 1) { ac.pair_lo = npairs * c / chunks; ac.pair_hi = npairs * (c + 1) / chunks; }
        FcQueued* list = d_list + (size_t)cap_c * c;
        unsigned long long* cnt = d_cnt + c;
        // the draw ranges of a chunk: as many as fill the GPU ~8 times over, at most one per pair
        const long long np_c = chunks > 1 ? ac.pair_hi - ac.pair_lo : npairs;
        const int gy_c = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(np_c, 64), (long long)h->sm_count * 8 * 8 / std::max(1, gx)));
        if (h->ncol == 4) {
          if (d_x) k_forecast_reduce<4, true><<<dim3(gx, gy_c), 256, 0, h->stream>>>(ac, d_mx, d_pa, list, cnt, cap_c);
          else k_forecast_reduce<4, false><<<dim3(gx, gy_c), 256, 0, h->stream>>>(ac, d_mx, d_pa, list, cnt, cap_c);
        } else {
          if (d_x) k_forecast_reduce<5, true><<<dim3(gx, gy_c), 256, 0, h->stream>>>(ac, d_mx, d_pa, list, cnt, cap_c);
          else k_forecast_reduce<5, false><<<dim3(gx, gy_c), 256, 0, h->stream>>>(ac, d_mx, d_pa, list, cnt, cap_c);
        }
        h->launches++;
        if (c + 1 < chunks) {                       // this chunk's second pass beside the next chunk's main pass
          cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
          cudaEventRecord(ev[c], h->stream);
          cudaStreamWaitEvent(h->copy_stream, ev[c], 0);
          launch_deferred(a, list, cnt, h->sm_count * side_blocks, h->copy_stream);
        } else {
          if (chunks > 1) {                         // everything the copy stream still runs adds to the same sums: join first
            cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
            cudaEventRecord(ev[c], h->copy_stream);
            cudaStreamWaitEvent(h->stream, ev[c], 0);
          }
          launch_deferred(a, list, cnt, gd, h->stream);
        }
      }
      for (int c = 0; c < chunks; ++c) if (ev[c]) cudaEventDestroy(ev[c]);
    }
    std::vector<unsigned long long> listed_c(chunks, 0ull);
    fe = cudaMemcpyAsync(listed_c.data(), d_cnt, sizeof(unsigned long long) * chunks, cudaMemcpyDeviceToHost, h->stream);
    if (fe == cudaSuccess) fe = cudaStreamSynchronize(h->stream);
    dfree(d_list); dfree(d_cnt);
    unsigned long long listed = 0;
    bool fits = true;
    for (auto v : listed_c) { listed += v; fits = fits && v <= cap_c; }
    h->fc_deferred_last = listed;
    if (fits) break;
    cap = listed + 65536;                         // a list was too small: once more, one list sized for the count
  }
  if (fe != cudaSuccess) {
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dfree(d_mx); dfree(d_pa);
    if (d_x) dfree(d_x);
    return fail(h, CLV_ERR_CUDA, "clv_forecast_resident failed: %s", cudaGetErrorString(fe));
  }
  k_scale<<<h->sm_count * 4, 256, 0, h->stream>>>(d_mx, d_pa, N, 1.0 / (double)(C * nd));
  h->launches++;
  cudaEventRecord(e1, h->stream);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  if (e == cudaSuccess && kernel_ms) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *kernel_ms = ms; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e == cudaSuccess && mean_x_star) e = cudaMemcpyAsync(mean_x_star, d_mx, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && p_alive) e = cudaMemcpyAsync(p_alive, d_pa, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && x_star) e = cudaMemcpyAsync(x_star, d_x, (size_t)(C * nd * N) * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dfree(d_mx); dfree(d_pa);
  if (d_x) dfree(d_x);
  if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_forecast_resident failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

// ---- "next" rows on the resident draws -----------------------------------------------------------
int clv_upload_draws(clv_sampler* h, const double* level1, int64_t n_draws) {
  if (!h || !level1 || n_draws < 1) return fail(h, CLV_ERR_ARG, "clv_upload_draws: bad argument");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const long long bytes = (long long)h->chains * n_draws * h->N * h->ncol * (long long)sizeof(double);
  if (h->draws_cap_bytes[0] < bytes) {
    if (h->d_draws[0]) dfree(h->d_draws[0]);
    h->d_draws[0] = nullptr; h->draws_cap_bytes[0] = 0;
    cudaError_t e = pool_malloc((void**)&h->d_draws[0], (size_t)bytes);
    if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "cannot allocate %lld bytes for the draws: %s", bytes, cudaGetErrorString(e));
    h->draws_cap_bytes[0] = bytes;
  }
  CK(h, copy_to_device_staged(h->d_draws[0], level1, (size_t)bytes, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->resident_draws = n_draws;
  return CLV_OK;
}

int clv_posterior_summary(clv_sampler* h, double mu_cap, double* out) {
  if (!h || !out) return fail(h, CLV_ERR_ARG, "null argument");
  if (h->resident_draws <= 0) return fail(h, CLV_ERR_STATE, "no resident draws: call clv_run (single device chunk) or clv_run_resident first");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  const long long n_tot = (long long)h->chains * h->resident_draws, N = h->N;
  int n_pad = 1;
  while (n_pad < n_tot) n_pad <<= 1;
  const size_t smem = (size_t)n_pad * sizeof(double);
  if (smem > 220 * 1024) return fail(h, CLV_ERR_ARG, "clv_posterior_summary: %lld draws per customer exceed the in-shared-memory sort (max 28160)", n_tot);
  double* d_out = nullptr;
  CK(h, dmalloc(&d_out, (size_t)N * SUMMARY_COLS));
  cudaError_t e;
  const int grid = (int)std::min<long long>(N, (long long)h->sm_count * 16);
  if (h->ncol == 4) {
    e = cudaFuncSetAttribute(k_posterior_summary<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_posterior_summary<4><<<grid, 256, smem, h->stream>>>(h->d_draws[0], n_tot, N, n_pad, mu_cap, d_out);
  } else {
    e = cudaFuncSetAttribute(k_posterior_summary<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_posterior_summary<5><<<grid, 256, smem, h->stream>>>(h->d_draws[0], n_tot, N, n_pad, mu_cap, d_out);
  }
  h->launches++;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * N * SUMMARY_COLS, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dfree(d_out);
  if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_posterior_summary failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

int clv_weekly_tracking(clv_sampler* h, const double* birth_week, const double* times, int n_weeks, uint64_t seed,
                        double* inc_mean) {
  if (!h || !birth_week || !times || !inc_mean) return fail(h, CLV_ERR_ARG, "null argument");
  if (n_weeks < 1 || n_weeks > 4096) return fail(h, CLV_ERR_ARG, "n_weeks must be in [1, 4096]");
  if (h->resident_draws <= 0) return fail(h, CLV_ERR_STATE, "no resident draws: call clv_run (single device chunk) or clv_run_resident first");
  CK(h, cudaSetDevice(h->cfg.device)); t_alloc_stream = h->stream; ensure_pool(h->cfg.device);
  if (int r = upload_rk()) return r;
  const long long n_tot = (long long)h->chains * h->resident_draws, N = h->N;
  double *d_birth = nullptr, *d_times = nullptr;
  unsigned long long* d_tot = nullptr;
  CK(h, dmalloc(&d_birth, (size_t)N));
  CK(h, dmalloc(&d_times, (size_t)n_weeks));
  CK(h, dmalloc(&d_tot, (size_t)n_weeks));
  cudaError_t e = cudaMemcpyAsync(d_birth, birth_week, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_times, times, sizeof(double) * n_weeks, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long) * n_weeks, h->stream);
  if (e == cudaSuccess) {
    const int gx = (int)std::min<long long>((N + 255) / 256, 65535);
    const int gy = (int)std::max<long long>(1, std::min<long long>(n_tot, (long long)h->sm_count * 8 * 4 / std::max(1, gx) + 1));
    const size_t smem = sizeof(unsigned long long) * n_weeks;
    if (h->ncol == 4) k_weekly_tracking<4><<<dim3(gx, gy), 256, smem, h->stream>>>(h->d_draws[0], n_tot, N, h->cfg.gid_offset, d_birth, d_times, n_weeks, seed, d_tot);
    else k_weekly_tracking<5><<<dim3(gx, gy), 256, smem, h->stream>>>(h->d_draws[0], n_tot, N, h->cfg.gid_offset, d_birth, d_times, n_weeks, seed, d_tot);
    h->launches++;
    e = cudaGetLastError();
  }
  std::vector<unsigned long long> tot((size_t)n_weeks);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tot.data(), d_tot, sizeof(unsigned long long) * n_weeks, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dfree(d_birth); dfree(d_times); dfree(d_tot);
  if (e != cudaSuccess) return fail(h, CLV_ERR_CUDA, "clv_weekly_tracking failed: %s", cudaGetErrorString(e));
  for (int w = 0; w < n_weeks; ++w) inc_mean[w] = (double)tot[w] / (double)n_tot;   // analysis_abe.py:459
  return CLV_OK;
}

// ---- generator --------------------------------------------------------------------------------
int clv_generate(const clv_generate_config* cfg, const double* beta, const double* gamma, int X_given, int T_cal_given,
                 int32_t* x, double* t_x, double* T_cal, double* X, int32_t* x_star, double* lambda_true,
                 double* mu_true, double* tau_true) {
  if (!cfg || !beta || !gamma || !x || !t_x || !T_cal || !X) return fail(nullptr, CLV_ERR_ARG, "clv_generate: null argument");
  if (cfg->n_cov < 1 || cfg->n_cov > CLV_MAX_K || cfg->n < 1) return fail(nullptr, CLV_ERR_ARG, "clv_generate: bad n / n_cov");
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  CK(nullptr, cudaSetDevice(cfg->device)); t_alloc_stream = nullptr; ensure_pool(cfg->device);
  if (int r = upload_rk()) return r;
  const long long n = cfg->n;
  const int K = cfg->n_cov;
  GenerateArgs a{};
  a.n = n; a.gid_offset = cfg->gid_offset; a.K = K; a.seed = cfg->seed;
  a.T_cal_lo = cfg->T_cal_lo; a.T_cal_hi = cfg->T_cal_hi; a.T_star = cfg->T_star;
  std::memcpy(a.beta, beta, sizeof(double) * K * 2);
  if (!chol_host(gamma, a.Lg, 2)) return fail(nullptr, CLV_ERR_NUMERIC, "gamma is not positive definite");
  int rc = 0;
  std::vector<void*> allocs;
  auto dm = [&](size_t bytes) -> void* { void* p = nullptr; if (pool_malloc(&p, bytes) != cudaSuccess) { rc = -1; return nullptr; } allocs.push_back(p); return p; };
  a.x = (int*)dm(n * 4); a.t_x = (double*)dm(n * 8); a.T_cal = (double*)dm(n * 8);
  a.Xc = (double*)dm((size_t)n * 8 * std::max(K - 1, 1));
  a.x_star = x_star ? (int*)dm(n * 4) : nullptr;
  a.lam = lambda_true ? (double*)dm(n * 8) : nullptr;
  a.mu = mu_true ? (double*)dm(n * 8) : nullptr;
  a.tau = tau_true ? (double*)dm(n * 8) : nullptr;
  auto cleanup = [&]() { for (void* p : allocs) dfree(p); };
  if (rc) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_generate: out of device memory"); }
  a.X_given = X_given; a.T_given = T_cal_given;
  if (T_cal_given && cudaMemcpy(a.T_cal, T_cal, n * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = -1;
  for (int k = 1; k < K && X_given && !rc; ++k)
    if (cudaMemcpy2D(a.Xc + (size_t)(k - 1) * n, 8, X + k, (size_t)K * 8, 8, (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) rc = -1;
  if (rc) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_generate: input upload failed"); }
  k_generate<<<(int)std::min<long long>((n + 255) / 256, 148 * 32), 256>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(x, a.x, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(t_x, a.t_x, n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(T_cal, a.T_cal, n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && x_star) e = cudaMemcpy(x_star, a.x_star, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && lambda_true) e = cudaMemcpy(lambda_true, a.lam, n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && mu_true) e = cudaMemcpy(mu_true, a.mu, n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && tau_true) e = cudaMemcpy(tau_true, a.tau, n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) {
    // X row-major n x K with intercept column
    std::vector<double> ones((size_t)n, 1.0);
    e = cudaMemcpy2D(X, (size_t)K * 8, ones.data(), 8, 8, (size_t)n, cudaMemcpyHostToHost);
    for (int k = 1; k < K && e == cudaSuccess; ++k)
      e = cudaMemcpy2D(X + k, (size_t)K * 8, a.Xc + (size_t)(k - 1) * n, 8, 8, (size_t)n, cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "clv_generate failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

// ---- event log -> CBS ---------------------------------------------------------------------------
int clv_elog2cbs(int device, int64_t n_events, const int64_t* cust, const int32_t* day, const double* sales,
                 int32_t T_cal_day, int32_t T_tot_day, double unit_days, int64_t* n_customers, int64_t* cust_out,
                 int32_t* x, double* t_x, double* litt, double* sales_out, double* sales_x, int32_t* first_day,
                 double* T_cal, double* T_star, int32_t* x_star, double* sales_star, double* first_sales) {
  if (!cust || !day || !n_customers || n_events < 1) return fail(nullptr, CLV_ERR_ARG, "clv_elog2cbs: bad argument");
  if (n_events >= (1ll << 31)) return fail(nullptr, CLV_ERR_ARG, "clv_elog2cbs: at most 2^31 - 1 events per call");
  if (!(unit_days > 0)) return fail(nullptr, CLV_ERR_ARG, "clv_elog2cbs: unit_days must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  CK(nullptr, cudaSetDevice(device)); ensure_pool(device);
  cudaStream_t st = nullptr;
  CK(nullptr, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  t_alloc_stream = st;
  const size_t n = (size_t)n_events;
  std::vector<void*> allocs;
  auto dm = [&](size_t bytes) -> void* { void* p = nullptr; if (pool_malloc(&p, bytes) != cudaSuccess) return nullptr; allocs.push_back(p); return p; };
  auto cleanup = [&]() { for (void* p : allocs) dfree(p); cudaStreamSynchronize(st); cudaStreamDestroy(st); t_alloc_stream = nullptr; };
  long long *k0 = (long long*)dm(n * 8), *k1 = (long long*)dm(n * 8);
  int *d0 = (int*)dm(n * 4), *d1 = (int*)dm(n * 4), *dtmp = (int*)dm(n * 4), *head = (int*)dm(n * 4), *idx = (int*)dm(n * 4);
  unsigned *p0 = (unsigned*)dm(n * 4), *p1 = (unsigned*)dm(n * 4);
  double *s0 = sales ? (double*)dm(n * 8) : nullptr, *s1 = sales ? (double*)dm(n * 8) : nullptr;
  if (!k0 || !k1 || !d0 || !d1 || !dtmp || !head || !idx || !p0 || !p1 || (sales && (!s0 || !s1))) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_elog2cbs: out of device memory"); }
  const int gb = (int)std::min<size_t>((n + 255) / 256, 148 * 32);
  cudaError_t e = cudaMemcpyAsync(k0, cust, n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d0, day, n * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && sales) e = cudaMemcpyAsync(s0, sales, n * 8, cudaMemcpyHostToDevice, st);
  // permutation-carrying stable sorts: by day, then by customer  =>  sorted by (cust, day), input order within ties.
  // Keys are signed (days before the epoch, negative ids): CUB's radix sort orders signed integers.
  if (e == cudaSuccess) { k_iota<<<gb, 256, 0, st>>>(p0, (long long)n); e = cudaGetLastError(); }
  size_t tb1 = 0, tb2 = 0, tb3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb1, d0, dtmp, p0, p1, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, k0, k1, p0, p1, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, head, idx, (int)n);
  void* tmp = dm(std::max(tb1, std::max(tb2, tb3)));
  if (!tmp) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_elog2cbs: out of device memory"); }
  // pass 1: permutation sorted by day
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb1, d0, dtmp, p0, p1, (int)n, 0, 32, st);
  // gather customer keys in that order, pass 2: stable sort by customer
  auto gather = [&](auto* dst, const auto* src, const unsigned* perm) {
    using T = std::remove_pointer_t<decltype(dst)>;
    k_gather<T><<<gb, 256, 0, st>>>(dst, src, perm, (long long)n);
  };
  if (e == cudaSuccess) { gather(k1, k0, p1); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb2, k1, k0, p1, p0, (int)n, 0, 64, st);   // k0 = sorted cust, p0 = final perm
  if (e == cudaSuccess) { gather(d1, d0, p0); if (sales) gather(s1, s0, p0); e = cudaGetLastError(); }
  if (e == cudaSuccess) { k_cbs_heads<<<gb, 256, 0, st>>>(k0, (long long)n, head); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb3, head, idx, (int)n, st);
  int last[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(&last[0], head + n - 1, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&last[1], idx + n - 1, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the number of distinct customers sizes the tables below
  if (e != cudaSuccess) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_elog2cbs failed: %s", cudaGetErrorString(e)); }
  const long long nc = (long long)last[0] + last[1];
  long long* starts = (long long*)dm((size_t)nc * 8);
  int* pos = (int*)dm((size_t)nc * 4);
  auto table = [&](CbsOut& o) {
    o.cust = (long long*)dm(nc * 8); o.x = (int*)dm(nc * 4); o.t_x = (double*)dm(nc * 8); o.litt = (double*)dm(nc * 8);
    o.sales = (double*)dm(nc * 8); o.sales_x = (double*)dm(nc * 8); o.first_day = (int*)dm(nc * 4); o.T_cal = (double*)dm(nc * 8);
    o.T_star = (double*)dm(nc * 8); o.x_star = (int*)dm(nc * 4); o.sales_star = (double*)dm(nc * 8); o.first_sales = (double*)dm(nc * 8);
    o.keep = (int*)dm(nc * 4);
    return o.cust && o.x && o.t_x && o.litt && o.sales && o.sales_x && o.first_day && o.T_cal && o.T_star && o.x_star && o.sales_star &&
           o.first_sales && o.keep;
  };
  CbsOut o{}, c{};
  size_t tb4 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb4, (int*)nullptr, (int*)nullptr, (int)nc);
  void* tmp2 = dm(std::max<size_t>(tb4, 8));
  if (!starts || !pos || !tmp2 || !table(o) || !table(c)) { cleanup(); return fail(nullptr, CLV_ERR_CUDA, "clv_elog2cbs: out of device memory"); }
  const int gc = (int)std::min<long long>((nc + 255) / 256, 148 * 32);
  k_cbs_starts<<<gb, 256, 0, st>>>(head, idx, (long long)n, starts);
  k_cbs_customers<<<gc, 256, 0, st>>>(k0, d1, sales ? s1 : nullptr, p0, sales ? s0 : nullptr, starts, (long long)n, nc, T_cal_day, T_tot_day,
                                      unit_days, o);
  e = cudaGetLastError();
  // drop customers without a calibration purchase ON THE DEVICE: scan of the keep flags, scatter of the kept rows
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp2, tb4, o.keep, pos, (int)nc, st);
  if (e == cudaSuccess) { k_cbs_compact<<<gc, 256, 0, st>>>(o, pos, nc, c); e = cudaGetLastError(); }
  int lastk[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(&lastk[0], o.keep + nc - 1, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&lastk[1], pos + nc - 1, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  const long long m = (long long)lastk[0] + lastk[1];
  auto dl = [&](void* dst, const void* src, size_t bytes) { if (e == cudaSuccess && dst && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st); };
  dl(cust_out, c.cust, m * 8); dl(x, c.x, m * 4); dl(first_day, c.first_day, m * 4); dl(x_star, c.x_star, m * 4);
  dl(t_x, c.t_x, m * 8); dl(litt, c.litt, m * 8); dl(sales_out, c.sales, m * 8); dl(sales_x, c.sales_x, m * 8);
  dl(T_cal, c.T_cal, m * 8); dl(T_star, c.T_star, m * 8); dl(sales_star, c.sales_star, m * 8); dl(first_sales, c.first_sales, m * 8);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "clv_elog2cbs failed: %s", cudaGetErrorString(e));
  *n_customers = m;
  return CLV_OK;
}

// ---- covariate standardisation (src/data_processing/2B_cdnow_elog2cbs_full.py:70-101) ----------------------------------
int clv_standardize(int device, int64_t n, const double* v, double scale, double* out, double* mean_out, double* sd_out) {
  if (!v || !out || n < 2) return fail(nullptr, CLV_ERR_ARG, "clv_standardize: need n >= 2 and non-null arrays");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  CK(nullptr, cudaSetDevice(device)); t_alloc_stream = nullptr; ensure_pool(device);
  double *d_v = nullptr, *d_o = nullptr, *d_part = nullptr, *d_ms = nullptr;
  cudaError_t e = dmalloc(&d_v, (size_t)n);
  if (e == cudaSuccess) e = dmalloc(&d_o, (size_t)n);
  if (e == cudaSuccess) e = dmalloc(&d_part, (size_t)COLSUM_BLOCKS);
  if (e == cudaSuccess) e = dmalloc(&d_ms, (size_t)2);
  if (e == cudaSuccess) e = cudaMemcpy(d_v, v, (size_t)n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    // the column is scaled first (first_sales * 1e-3, :68): d_o = v * scale (k_zscore with mean 0, sd 1), then the mean
    // and pandas' std (ddof = 1) of the scaled values, then the z-scores
    const double ms01[2] = {0.0, 1.0};
    e = cudaMemcpy(d_ms, ms01, 16, cudaMemcpyHostToDevice);
    const int gz = (int)std::min<long long>((n + 255) / 256, 148 * 32);
    if (e == cudaSuccess) { k_zscore<<<gz, 256>>>(d_v, n, d_ms, d_ms + 1, scale, d_o); e = cudaGetLastError(); }
    if (e == cudaSuccess) { k_col_sum<<<COLSUM_BLOCKS, COLSUM_THREADS>>>(d_o, n, 0, d_ms, d_part); k_col_fold<<<1, COLSUM_THREADS>>>(d_part, COLSUM_BLOCKS, (double)n, 0, d_ms); }
    if (e == cudaSuccess) { k_col_sum<<<COLSUM_BLOCKS, COLSUM_THREADS>>>(d_o, n, 1, d_ms, d_part); k_col_fold<<<1, COLSUM_THREADS>>>(d_part, COLSUM_BLOCKS, (double)(n - 1), 1, d_ms + 1); }
    if (e == cudaSuccess) { k_zscore<<<gz, 256>>>(d_v, n, d_ms, d_ms + 1, scale, d_o); e = cudaGetLastError(); }
  }
  double ms[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(ms, d_ms, 16, cudaMemcpyDeviceToHost);
  dfree(d_v); dfree(d_o); dfree(d_part); dfree(d_ms);
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "clv_standardize failed: %s", cudaGetErrorString(e));
  if (mean_out) *mean_out = ms[0];
  if (sd_out) *sd_out = ms[1];
  return CLV_OK;
}

int clv_recode(int device, int64_t n, const int32_t* codes, const double* table, int n_table, double* out) {
  if (!codes || !table || !out || n < 1 || n_table < 1) return fail(nullptr, CLV_ERR_ARG, "clv_recode: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  CK(nullptr, cudaSetDevice(device)); t_alloc_stream = nullptr; ensure_pool(device);
  int* d_c = nullptr; double *d_t = nullptr, *d_o = nullptr;
  cudaError_t e = dmalloc(&d_c, (size_t)n);
  if (e == cudaSuccess) e = dmalloc(&d_t, (size_t)n_table);
  if (e == cudaSuccess) e = dmalloc(&d_o, (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpy(d_c, codes, (size_t)n * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_t, table, (size_t)n_table * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { k_recode<<<(int)std::min<long long>((n + 255) / 256, 148 * 32), 256>>>(d_c, n, d_t, n_table, d_o); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost);
  dfree(d_c); dfree(d_t); dfree(d_o);
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "clv_recode failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

// ---- test hook: customer shards of one problem on ONE device, in lockstep ------------------------------------------------
int clv_debug_lockstep_advance(clv_sampler** shards, int n_shards, int64_t n_sweeps) {
  if (!shards || n_shards < 2 || n_shards > P2P_MAX_WORLD || n_sweeps < 0) return fail(nullptr, CLV_ERR_ARG, "clv_debug_lockstep_advance: bad argument");
  clv_sampler* h0 = shards[0];
  for (int r = 0; r < n_shards; ++r) {
    clv_sampler* h = shards[r];
    if (!h || !h->inited) return fail(h0, CLV_ERR_STATE, "lockstep: every shard must be initialised");
    if (h->cfg.device != h0->cfg.device || h->D != h0->D || h->K != h0->K || h->S != h0->S || h->chains != h0->chains ||
        h->cfg.seed != h0->cfg.seed || h->cfg.n_global != h0->cfg.n_global || h->sweeps_done != h0->sweeps_done ||
        h->cfg.rng_mode == CLV_RNG_INJECTED || h->comm || h->p2p)
      return fail(h0, CLV_ERR_ARG, "lockstep: shards must share device, model, chains, seed, n_global and sweep count, and have no communicator");
  }
  CK(h0, cudaSetDevice(h0->cfg.device)); t_alloc_stream = h0->stream; ensure_pool(h0->cfg.device);
  for (int r = 0; r < n_shards; ++r) CK(h0, cudaStreamSynchronize(shards[r]->stream));
  const size_t mail_bytes = sizeof(unsigned long long) * 2 * P2P_MAX_WORLD * (size_t)h0->chains * 2 * NSTAT_MAX;
  std::vector<void*> mail(n_shards, nullptr);
  Level2Args* d_ranks = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(h0->stream);
    for (void* m : mail) if (m) cudaFree(m);
    if (d_ranks) cudaFree(d_ranks);
    for (int r = 0; r < n_shards; ++r) { shards[r]->p2p = false; shards[r]->world = 1; shards[r]->rank = 0; }
  };
#define CKL(call) do { cudaError_t e4 = (call); if (e4 != cudaSuccess) { rc = fail(h0, CLV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e4)); cleanup(); return rc; } } while (0)
  for (int r = 0; r < n_shards; ++r) { CKL(cudaMalloc(&mail[r], mail_bytes)); CKL(cudaMemset(mail[r], 0, mail_bytes)); }
  CKL(cudaMalloc((void**)&d_ranks, sizeof(Level2Args) * n_shards));
  for (int r = 0; r < n_shards; ++r) {
    clv_sampler* h = shards[r];
    h->p2p = true; h->world = n_shards; h->rank = r;
    for (int q = 0; q < n_shards; ++q) h->peer_mail[q] = (unsigned long long*)mail[q];
  }
  std::vector<Level2Args> l2(n_shards);
  const int mode = h0->cfg.rng_mode;
  for (int64_t it = 0; it < n_sweeps; ++it) {
    for (int r = 0; r < n_shards; ++r) {
      clv_sampler* h = shards[r];
      l2[r] = base_l2(h);
      l2[r].sweep = (uint32_t)(h->sweeps_done + 1);
      l2[r].tag = (uint32_t)(((h0->init_epoch % 255ull) + 1ull) << 24) | (l2[r].sweep & 0xffffffu);
    }
    auto do_l2 = [&]() -> cudaError_t {
      cudaError_t e = cudaMemcpyAsync(d_ranks, l2.data(), sizeof(Level2Args) * n_shards, cudaMemcpyHostToDevice, h0->stream);
      if (e != cudaSuccess) return e;
      e = cudaStreamSynchronize(h0->stream);            // l2 (pageable) is rewritten next sweep
      if (e != cudaSuccess) return e;
      void* args[] = {&d_ranks};
      const void* fn = h0->D == 2 ? (const void*)k_level2_ranks<2> : (const void*)k_level2_ranks<3>;
      return cudaLaunchCooperativeKernel(fn, dim3(h0->chains, n_shards), dim3(32), args, 0, h0->stream);
    };
    auto do_sweeps = [&]() -> cudaError_t {
      for (int r = 0; r < n_shards; ++r) {
        clv_sampler* h = shards[r];
        SweepArgs a = base_args(h);
        a.sweep = (uint32_t)(h->sweeps_done + 1);
        a.store_zt = (it + 1 == n_sweeps) ? 1 : 0;
        dim3 grid(h->grid_x, h->chains), block(SWEEP_THREADS);
        cudaError_t e = (h->D == 2)
            ? (mode == MODE_STRICT ? launch_kernel(k_sweep<2, MODE_STRICT>, grid, block, h->stats_smem, h0->stream, false, a)
                                   : launch_kernel(k_sweep<2, MODE_FAST>, grid, block, h->stats_smem, h0->stream, false, a))
            : (mode == MODE_STRICT ? launch_kernel(k_sweep<3, MODE_STRICT>, grid, block, h->stats_smem, h0->stream, false, a)
                                   : launch_kernel(k_sweep<3, MODE_FAST>, grid, block, h->stats_smem, h0->stream, false, a));
        if (e != cudaSuccess) return e;
        h->launches++;
      }
      return cudaSuccess;
    };
    if (h0->D == 2) { CKL(do_l2()); CKL(do_sweeps()); }
    else { CKL(do_sweeps()); CKL(do_l2()); }
    for (int r = 0; r < n_shards; ++r) shards[r]->sweeps_done++;
  }
#undef CKL
  CK(h0, cudaStreamSynchronize(h0->stream));
  int flag = 0;
  for (int r = 0; r < n_shards && !flag; ++r) cudaMemcpy(&flag, shards[r]->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  cleanup();
  if (flag == 2) return fail(h0, CLV_ERR_COMM, "lockstep: mailbox all-reduce timed out");
  if (flag) return fail(h0, CLV_ERR_NUMERIC, "lockstep: level-2 scale matrix not positive definite or non-finite");
  return CLV_OK;
}

// ---- test hook: level-1 variates ------------------------------------------------------------------
int clv_debug_variates(int device, uint64_t seed, uint32_t sweep, int32_t step, int rng_mode, int64_t n, double* t3_l, double* t3_m, double* u_acc) {
  if (!t3_l || !t3_m || !u_acc || n < 1 || step < 0) return fail(nullptr, CLV_ERR_ARG, "clv_debug_variates: bad argument");
  if (rng_mode != CLV_RNG_PHILOX_FAST && rng_mode != CLV_RNG_PHILOX_STRICT) return fail(nullptr, CLV_ERR_ARG, "rng_mode must be fast or strict");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  CK(nullptr, cudaSetDevice(device)); t_alloc_stream = nullptr; ensure_pool(device);
  double* d = nullptr;
  CK(nullptr, dmalloc(&d, (size_t)n * 3));
  const PhiloxRoundKeys rk = round_keys(seed);
  if (rng_mode == CLV_RNG_PHILOX_STRICT) k_debug_variates<MODE_STRICT><<<148 * 4, 256>>>(rk, sweep, step, n, d, d + n, d + 2 * n);
  else k_debug_variates<MODE_FAST><<<148 * 4, 256>>>(rk, sweep, step, n, d, d + n, d + 2 * n);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(t3_l, d, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(t3_m, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(u_acc, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  dfree(d);
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "clv_debug_variates failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

// ---- test hook: parallel host memcpy ----------------------------------------------------------------
int clv_debug_host_copy(void* dst, const void* src, int64_t bytes) {
  if (!dst || !src || bytes < 0) return fail(nullptr, CLV_ERR_ARG, "clv_debug_host_copy: bad argument");
  HostCopyPool& pool = HostCopyPool::get();
  pool.copy(dst, src, (size_t)bytes);
  return pool.threads();
}

// ---- test hook: the first-touch threads of clv_run on a caller's buffer (returns when every page has been touched) ----
int clv_debug_first_touch(void* buf, int64_t bytes, int threads) {
  if (!buf || bytes < 0 || threads < 1) return fail(nullptr, CLV_ERR_ARG, "clv_debug_first_touch: bad argument");
  FirstToucher t;
  std::vector<std::pair<char*, size_t>> ranges;
  const size_t piece = 1u << 20;
  for (size_t off = 0; off < (size_t)bytes; off += piece) ranges.emplace_back((char*)buf + off, std::min(piece, (size_t)bytes - off));
  t.start(std::move(ranges), threads);
  t.finish();
  return CLV_OK;
}

// ---- issue-rate peaks ---------------------------------------------------------------------------
int clv_measure_issue_peaks(int device, double* out4) {
  if (!out4) return fail(nullptr, CLV_ERR_ARG, "null argument");
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) return fail(nullptr, CLV_ERR_CUDA, "no CUDA device available");
  CK(nullptr, cudaSetDevice(device)); t_alloc_stream = nullptr; ensure_pool(device);
  cudaDeviceProp prop;
  CK(nullptr, cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float* d_out = nullptr;
  CK(nullptr, dmalloc(&d_out, (size_t)blocks * threads));
  cudaEvent_t e1, e2;
  cudaEventCreate(&e1); cudaEventCreate(&e2);
  for (int w = 0; w < 4; ++w) {
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e1);
      if (w == 0) k_peak<0><<<blocks, threads>>>(d_out, iters, 1.0f);
      else if (w == 1) k_peak<1><<<blocks, threads>>>(d_out, iters, 1.0f);
      else if (w == 2) k_peak<2><<<blocks, threads>>>(d_out, iters, 1.0f);
      else k_peak<3><<<blocks, threads>>>(d_out, iters, 1.0f);
      cudaEventRecord(e2);
      cudaEventSynchronize(e2);
      float ms = 0;
      cudaEventElapsedTime(&ms, e1, e2);
      double ops = (double)blocks * threads * iters * 8.0;
      if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
    }
    out4[w] = best;
  }
  cudaEventDestroy(e1); cudaEventDestroy(e2);
  dfree(d_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, CLV_ERR_CUDA, "peak kernels failed: %s", cudaGetErrorString(e));
  return CLV_OK;
}

}  // extern "C"
