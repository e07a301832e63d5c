// clv_forecast.cuh — posterior-predictive kernels: draw_future_transactions (bi:506-546, tri:660-749),
// the fused per-customer reductions the analysis scripts take over its output (mean x*, P(alive):
// utils/analysis_bi_helpers.py:75-110), and the synthetic-customer generator (bi:95-187).
// HBM-bound: one 32/40-byte level-1 row in, one int64 (+ one double) out per (draw, customer) cell.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "clv_rng.cuh"

namespace clv {

constexpr int RK_TABLE = 64;
__constant__ double c_rk[RK_TABLE + 1];   // c_rk[k] = 1.0 / k
// Forecast-domain Philox slots (third counter word; one counter -> one variate, ranges disjoint):
//   0                        the Poisson uniforms of a draw pair
//   FC_SLOT_PTRS + t         PTRS attempt t < 4096 of a cell with a large mean
//   FC_SLOT_WEEK + w / 4     weekly-tracking uniforms, week w < 4096
//   FC_SLOT_WEEK_PTRS + 4096 w + t   PTRS attempts of week w
//   FC_SLOT_SPEND + j / 2    per-transaction spend normals (j < 2^31), or the single CLT normal (j = 0) for huge counts
constexpr uint32_t FC_SLOT_PTRS = 0x00010000u, FC_SLOT_WEEK = 0x00020000u, FC_SLOT_WEEK_PTRS = 0x01000000u,
                   FC_SLOT_SPEND = 0x80000000u;
// Above this many transactions the total spend of a cell is drawn from the normal limit of the sum of i.i.d. log-normals
// (mean x m1, variance x v): bounded work per cell whatever the Poisson draw (a PTRS overflow sentinel must never
// become a loop count).
constexpr long long SPEND_EXACT_MAX = 4096;
constexpr int RKF_TABLE = 1032;
__constant__ float c_rkf[RKF_TABLE];      // c_rkf[k] = 1.0f / k (fp32 screen of the Poisson inversion)

// Poisson draw by sequential CDF inversion from zero; arithmetic fixed to match
// oracle/abe_oracle.py:poisson_inversion bit for bit.
__device__ __noinline__ long long poisson_inversion(double m, double u) {
  double p = exp(-m), cdf = p;
  long long k = 0;
  const double cap = floor(4096.0 + 16.0 * m);
  while (u > cdf && (double)k < cap) {
    ++k;
    double rk = (k <= RK_TABLE) ? c_rk[k] : 1.0 / (double)k;
    p = __dmul_rn(__dmul_rn(p, m), rk);   // no FMA contraction: the oracle rounds after every operation
    cdf = __dadd_rn(cdf, p);
  }
  return k;
}

__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// The same map u -> k, screened in fp32: the chop-down runs on the FP32 pipe from fp32 images (mf, uf) of the mean and
// the uniform, four CDF terms per trip and branch-free inside a trip (k = number of CDF values below u; the lanes of a
// warp stay in lockstep, so the reciprocal 1/k is a uniform constant-bank operand).  The result is accepted only when
// uf is further than a guard band (>> the fp32 error of the running CDF, of mf and of uf) from both neighbouring CDF
// values; otherwise `exact()` decides on the fp64 quantities -- which are only built then.  Hence the returned k ALWAYS
// equals poisson_inversion(m, u).
template <typename Exact>
__device__ __forceinline__ long long poisson_inversion_screened(float mf, float uf, Exact exact) {
  if (mf >= 0.0f && mf < 60.0f) {
    float p = ex2_ftz(-1.4426950408889634f * mf), cdf = p, prev = -1.0f;
    int k = 0, base = 0;
    bool searching = uf > cdf;
    while (searching && base < 1024) {
#pragma unroll
      for (int j = 1; j <= 4; ++j) {
        const bool gt = uf > cdf;            // cdf == CDF(k): still below u?
        prev = gt ? cdf : prev;
        k += gt ? 1 : 0;
        p = (p * mf) * c_rkf[base + j];
        cdf = gt ? cdf + p : cdf;            // frozen at CDF(k) once found
      }
      base += 4;
      searching = uf > cdf;
    }
    const float tol = 2e-6f * (8.0f + mf);
    if (!searching && uf < cdf - tol && uf > prev + tol) return (long long)k;
  }
  return exact();
}

// Large means (production path only): Hoermann's transformed rejection "PTRS" (the algorithm NumPy uses for lam >= 10),
// exact, ~1.2 attempts whatever the mean -- a customer with lambda * T_star in the hundreds must not cost hundreds of
// dependent fp64 iterations per draw.  Attempt t takes its two uniforms from Philox block (gid, draw, FC_SLOT_PTRS + t).
__device__ __noinline__ long long poisson_ptrs(double lam, uint32_t gid, uint32_t gdraw, PhiloxKey key, uint32_t slot0 = FC_SLOT_PTRS) {
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam;
  const double a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  const double vr = 0.9277 - 3.6224 / (b - 2.0);
  for (uint32_t t = 0; t < 4096u; ++t) {
    const uint4 r = philox4x32_10(gid, gdraw, slot0 + t, DOM_FORECAST, key);
    const double U = u53(r.x, r.y) - 0.5, V = u53(r.z, r.w);
    const double us = 0.5 - fabs(U);
    const double kd = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return kd < 9.2e18 ? (long long)kd : 0x7fffffffffffffffll;
    if (kd < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kd * loglam - lgamma(kd + 1.0))
      return kd < 9.2e18 ? (long long)kd : 0x7fffffffffffffffll;
  }
  return (long long)lam;
}

constexpr double PTRS_MIN_MEAN = 60.0;   // == the fp32 range limit of the screened inversion

__device__ __forceinline__ double future_horizon(double T_cal, double tau, double zf, double T_star) {
  // bi:535-540: alive -> T_star ; churned -> clip(tau - T_cal, 0, T_star)
  return (zf > 0.5) ? T_star : fmin(fmax(tau - T_cal, 0.0), T_star);
}

struct ForecastArgs {
  const double* level1;   // [n_draws][N][NCOL]
  const double* T_cal;    // [N]
  long long n_draws, N;
  double T_star, sigma_s;
  uint64_t seed;
  long long gid_offset, draw_offset;
  const double* u;               // injected uniforms [n_draws][N]
  const double* eps;             // injected per-transaction normals (flat)
  const long long* eps_offset;   // [n_draws][N]
  long long* x_out;              // [n_draws][N] (nullable)
  double* spend_out;             // [n_draws][N] (nullable)
  PhiloxRoundKeys rk;            // round keys of `seed` (launch constants; filled by the host)
  long long pair_lo, pair_hi;    // k_forecast_reduce: draw pairs [pair_lo, pair_hi) of this launch (pair_hi == 0: all)
};

// Forecast uniforms: one Philox block serves the two draws 2g, 2g+1 of a customer (global draw index, chain-major):
// counter (gid, g, 0, DOM_FORECAST), words (x,y) for the even draw and (z,w) for the odd one.
__device__ __forceinline__ double forecast_u(uint4 r, long long gdraw) {
  return (gdraw & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
}

template <int NCOL>
__device__ __forceinline__ void load_row(const double* row, double& lam, double& tau, double& zf, double& eta) {
  if (NCOL == 4) {
    const double2 r0 = __ldcs(reinterpret_cast<const double2*>(row));       // streamed once: evict-first
    const double2 r1 = __ldcs(reinterpret_cast<const double2*>(row) + 1);
    lam = r0.x; tau = r1.x; zf = r1.y; eta = 0.0;
  } else {
    lam = __ldcs(row); tau = __ldcs(row + 2); zf = __ldcs(row + 3); eta = __ldcs(row + 4);
  }
}

// Slow path of a cell: exact fp64 mean lambda * clip(tau - T_cal, 0, T_star) re-read from the level-1 row, then PTRS for
// large means or the fp64 inversion on the 53-bit uniform.  Out of line: keeps the fast path's register count low.
__device__ __noinline__ long long forecast_cell_slow(const double* row, double T, double T_star, uint32_t ua, uint32_t ub,
                                                     uint32_t gid, uint32_t gdraw, PhiloxKey key) {
  const double lam = row[0], tau = row[2], zf = row[3];
  const double m = lam * future_horizon(T, tau, zf, T_star);
  if (m >= PTRS_MIN_MEAN) return poisson_ptrs(m, gid, gdraw, key);
  return poisson_inversion(m, u53(ua, ub));
}

// x* of one (draw, customer) cell (bi:535-543) from the fp32 images of its row.
__device__ __forceinline__ long long forecast_cell(const double* row, float lamf, float dtf, bool alive, double T, double T_star,
                                                   float T_star_f, uint32_t ua, uint32_t ub, uint32_t gid, uint32_t gdraw,
                                                   PhiloxKey key) {
  const float hf = alive ? T_star_f : fminf(fmaxf(dtf, 0.0f), T_star_f);
  const float mf = lamf * hf;
  if (!(mf < 59.0f)) return forecast_cell_slow(row, T, T_star, ua, ub, gid, gdraw, key);   // near / beyond the PTRS switch, NaN
  return poisson_inversion_screened(mf, u24f(ua), [&]() { return forecast_cell_slow(row, T, T_star, ua, ub, gid, gdraw, key); });
}

// fp32 images of a level-1 row: lambda, tau - T_cal, alive
template <int NCOL>
__device__ __forceinline__ void load_row_f32(const double* row, double T, float& lamf, float& dtf, bool& alive) {
  double lam, tau, zf, eta;
  load_row<NCOL>(row, lam, tau, zf, eta);
  lamf = (float)lam;
  dtf = (float)(tau - T);
  alive = zf > 0.5;
}

// One thread per (customer, pair of draws): x* (and spend) for every cell.
template <int NCOL, bool INJECT>
__global__ void __launch_bounds__(256) k_forecast(ForecastArgs a) {
  const PhiloxKey key = seed_key(a.seed);
  const bool spend = (NCOL == 5) && a.spend_out != nullptr;
  const float T_star_f = (float)a.T_star;
  const long long gp0 = a.draw_offset >> 1, gp1 = (a.draw_offset + a.n_draws - 1) >> 1;   // global pair range
  for (long long gp = gp0 + blockIdx.y; gp <= gp1; gp += gridDim.y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N;
         i += (long long)gridDim.x * blockDim.x) {
      const uint32_t gid = (uint32_t)(a.gid_offset + i);
      const double T = a.T_cal[i];
      uint4 r = make_uint4(0, 0, 0, 0);
      if (!INJECT) r = philox4x32_10_rk(gid, (uint32_t)gp, 0u, DOM_FORECAST, a.rk);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long gdraw = 2 * gp + h, d = gdraw - a.draw_offset;
        if (d < 0 || d >= a.n_draws) continue;
        const long long cell = d * a.N + i;
        double lam, tau, zf, eta;
        load_row<NCOL>(a.level1 + cell * NCOL, lam, tau, zf, eta);
        long long xs;                                                         // bi:543
        const bool alive = zf > 0.5;
        const float dtf = (float)(tau - T);
        if (INJECT) {
          const double u = a.u[cell];
          const float hf = alive ? T_star_f : fminf(fmaxf(dtf, 0.0f), T_star_f);
          xs = poisson_inversion_screened((float)lam * hf, (float)u, [&]() {
            return poisson_inversion(lam * future_horizon(T, tau, zf, a.T_star), u);
          });
        } else {
          xs = forecast_cell(a.level1 + cell * NCOL, (float)lam, dtf, alive, T, a.T_star, T_star_f, h ? r.z : r.x, h ? r.w : r.y,
                             gid, (uint32_t)gdraw, key);
        }
        if (a.x_out) __stcs(&a.x_out[cell], xs);
        if (spend) {
          // tri:730-737: sum of x* log-normal transactions, log-mean = eta column as stored (Q7)
          double tot = 0.0;
          if (!INJECT && xs > SPEND_EXACT_MAX) {
            // sum of xs i.i.d. LogNormal(eta, sigma): normal limit (relative error of the law ~ xs^-1/2 * skewness < 3 %)
            const double s2 = a.sigma_s * a.sigma_s, m1 = exp(eta + 0.5 * s2);
            double nc, ns;
            normal_pair_u53(philox4x32_10(gid, (uint32_t)gdraw, FC_SLOT_SPEND, DOM_FORECAST, key), &nc, &ns);
            tot = fmax(0.0, (double)xs * m1 + sqrt((double)xs * (exp(s2) - 1.0)) * m1 * nc);
          } else {
            for (long long j = 0; j < xs; ++j) {
              double n;
              if (INJECT) n = a.eps[a.eps_offset[cell] + j];
              else {
                double nc, ns;
                normal_pair_u53(philox4x32_10(gid, (uint32_t)gdraw, FC_SLOT_SPEND + (uint32_t)(j >> 1), DOM_FORECAST, key), &nc, &ns);
                n = (j & 1) ? ns : nc;
              }
              tot += exp(eta + a.sigma_s * n);
            }
          }
          a.spend_out[cell] = tot;
        }
      }
    }
  }
}

// ---- fused reductions over the draws still resident in HBM -----------------------------------------------------------
// Quick path of a cell: the chop-down over the first FC_QUICK_TERMS CDF values, fully unrolled and branch-free (the
// reciprocals are immediates), from the fp32 images of the mean and the uniform.  Returns k when the uniform falls
// below CDF(k) for some k < FC_QUICK_TERMS AND is further than the guard band from both neighbouring CDF values (then
// k equals the fp64 inversion, see poisson_inversion_screened); -1 otherwise: the cell is DEFERRED.
constexpr int FC_QUICK_TERMS = 8;
__device__ __forceinline__ int poisson_quick(float mf, float uf) {
  float p = ex2_ftz(-1.4426950408889634f * mf), cdf = p, prev = -1.0f;
  int k = 0;
#pragma unroll
  for (int j = 1; j <= FC_QUICK_TERMS; ++j) {
    const bool gt = uf > cdf;
    prev = gt ? cdf : prev;
    k += gt ? 1 : 0;
    p = (p * mf) * (1.0f / (float)j);
    cdf = gt ? cdf + p : cdf;
  }
  const float tol = 2e-6f * (8.0f + mf);
  const bool ok = !(uf > cdf) && uf < cdf - tol && uf > prev + tol;
  return ok ? k : -1;
}

// Per customer sum of x* and of z (P(alive) = mean z, analysis_bi_helpers.py:98) over the draws resident in HBM
// (draw_offset == 0), x* optionally materialised (WRITE_X).  blockIdx.y splits the draw pairs; partial sums are
// integers, so the atomicAdd order does not matter.
//
// The kernel is bound by instruction issue before it is bound by HBM (profiles/r01_forecast_ncu_summary.md: 77 % of
// the issue slots at 17 of 32 lanes active -- the lanes of a warp wait for the one cell with a large count).  So a
// cell gets only the unrolled quick path in lockstep; the few that need more (x* >= 8, the tie zone, means beyond the
// fp32 screen, PTRS) are DEFERRED to a per-warp queue in shared memory and worked off 32 at a time by the whole
// warp -- full lanes for the divergent work.  Deferred results return to the owner lane through a shared array.
constexpr int FC_WARPS = 8, FC_QCAP = 96;
struct FcQueued { uint32_t gdraw, owner; };      // a deferred cell: global draw index, customer (index within the shard)

// quick path of one cell; a cell that needs more is queued (qn and the queue belong to the warp)
template <bool WRITE_X>
__device__ __forceinline__ void fc_quick_cell(float lamf, float dtf, bool alive, float T_star_f, uint32_t ua, uint32_t gdraw,
                                              uint32_t cust, bool live, int lane, FcQueued* q, int& qn, int& sx, int& sz, long long* x_cell) {
  const float hf = alive ? T_star_f : fminf(fmaxf(dtf, 0.0f), T_star_f);
  const float mf = lamf * hf;
  const int k = (mf < 24.0f) ? poisson_quick(mf, u24f(ua)) : -1;       // beyond ~24 the first 8 terms almost never decide
  const bool defer = live && k < 0;
  const unsigned m = __ballot_sync(0xffffffffu, defer);
  if (defer) q[qn + __popc(m & ((1u << lane) - 1u))] = FcQueued{gdraw, cust};
  qn += __popc(m);
  if (live && k > 0) sx += k;
  if (live && alive) ++sz;
  if (WRITE_X && live && k >= 0) __stcs(x_cell, (long long)k);
}

// The warp moves q[first .. last) to the global list of deferred cells: one atomicAdd per batch, coalesced 8-byte stores.
// Entries beyond the list's capacity are dropped and counted: the host then sizes the list for the count and reruns.
__device__ __forceinline__ void fc_flush(const FcQueued* q, int first, int last, FcQueued* g_list, unsigned long long* g_count,
                                         unsigned long long cap, int lane) {
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(g_count, (unsigned long long)(last - first));
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int b = first + lane; b < last; b += 32) {
    const unsigned long long pos = base + (unsigned long long)(b - first);
    if (pos < cap) g_list[pos] = q[b];
  }
  __syncwarp();
}

#ifndef CLV_FC_MINBLOCKS
#define CLV_FC_MINBLOCKS 6
#endif
// Register-fed main pass: one thread per customer, the two rows of a draw pair in flight in registers.  No function
// call in the kernel (cells that need more than the quick path go to the global list for k_forecast_deferred).
template <int NCOL, bool WRITE_X>
__global__ void __launch_bounds__(256, CLV_FC_MINBLOCKS) k_forecast_reduce(ForecastArgs a, double* sum_x, double* sum_z, FcQueued* g_list,
                                                                          unsigned long long* g_count, unsigned long long g_cap) {
  __shared__ FcQueued s_q[FC_WARPS][FC_QCAP];
  const float T_star_f = (float)a.T_star;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the launch's range of draw pairs (the whole forecast, or one chunk of it: the host overlaps the second pass of a
  // chunk with the main pass of the next), cut into gridDim.y ranges
  const int p_lo = a.pair_hi > 0 ? (int)a.pair_lo : 0, p_hi = a.pair_hi > 0 ? (int)a.pair_hi : (int)((a.n_draws + 1) >> 1);
  const int per = (p_hi - p_lo + gridDim.y - 1) / gridDim.y;
  const int pa = p_lo + blockIdx.y * per, pb = min(p_hi, pa + per);
  const long long stride = a.N * NCOL;                 // doubles between consecutive draws of one customer
  FcQueued* q = s_q[warp];
  const long long nwt = (a.N + 31) / 32;               // warp tiles of 32 consecutive customers
  for (long long wt = (long long)blockIdx.x * FC_WARPS + warp; wt < nwt; wt += (long long)gridDim.x * FC_WARPS) {
    const long long i = wt * 32 + lane;
    const bool valid = i < a.N;
    const long long ic = valid ? i : a.N - 1;          // out-of-range lanes shadow the last customer (results dropped)
    const double T = a.T_cal[ic];
    const uint32_t gid = (uint32_t)(a.gid_offset + ic);
    int sx = 0, sz = 0, qn = 0;                        // quick-path counts are < 8 per cell; qn: queue length (warp uniform)
    const double* row = a.level1 + (2ll * pa * a.N + ic) * NCOL;
    for (int gp = pa; gp < pb; ++gp, row += 2 * stride) {
      const bool v1 = 2ll * gp + 1 < a.n_draws;        // only the very last pair can be half empty
      float lam0, dt0, lam1 = 0.0f, dt1 = 0.0f;
      bool al0, al1 = false;
      load_row_f32<NCOL>(row, T, lam0, dt0, al0);        // both rows in flight before the arithmetic starts
      if (v1) load_row_f32<NCOL>(row + stride, T, lam1, dt1, al1);
      const uint4 r = philox4x32_10_rk(gid, (uint32_t)gp, 0u, DOM_FORECAST, a.rk);
      long long* xc = WRITE_X ? a.x_out + (2ll * gp) * a.N + ic : nullptr;
      fc_quick_cell<WRITE_X>(lam0, dt0, al0, T_star_f, r.x, 2u * (uint32_t)gp, (uint32_t)ic, valid, lane, q, qn, sx, sz, xc);
      fc_quick_cell<WRITE_X>(lam1, dt1, al1, T_star_f, r.z, 2u * (uint32_t)gp + 1u, (uint32_t)ic, valid && v1, lane, q, qn, sx, sz,
                             WRITE_X ? xc + a.N : nullptr);
      if (qn >= 32) {                                    // full batches only: the remainder stays queued
        __syncwarp();
        const int rem = qn & 31;
        fc_flush(q, rem, qn, g_list, g_count, g_cap, lane);
        qn = rem;
      }
    }
    __syncwarp();
    if (qn > 0) fc_flush(q, 0, qn, g_list, g_count, g_cap, lane);
    if (valid) {
      if (gridDim.y == 1 && a.pair_hi == 0) {
        sum_x[i] = (double)sx;
        sum_z[i] = (double)sz;
      } else {
        atomicAdd(&sum_x[i], (double)sx);
        atomicAdd(&sum_z[i], (double)sz);
      }
    }
    __syncwarp();
  }
}

// ---- the same reduction fed by the TMA engine ----------------------------------------------------------------------
// k_forecast_reduce above keeps two level-1 rows per thread in flight in REGISTERS: at the 40 registers that six
// resident blocks allow, the rows plus the Philox state spill (51 local-memory instructions per draw pair), and the loads
// are only in flight while their warp waits for them.  Here the rows never pass through registers on their way in: for
// a tile of 256 consecutive customers one draw is a contiguous 8 KB (10 KB for the trivariate layout) of the resident
// array [draw][customer][col], which ONE thread fetches with a bulk asynchronous copy (cp.async.bulk, the TMA engine;
// UBLKCP in SASS) into a ring of FC_STAGES shared-memory stages of one draw pair each, completion signalled on an
// mbarrier.  Every thread then reads its own 32-byte row from shared memory when it needs it; a warp that has taken its
// rows arrives on the stage's "empty" barrier and the producer refills the stage FC_STAGES pairs ahead.  HBM always has
// (FC_STAGES - 1) x 16 KB per block in flight whatever the warps are doing, and the kernel needs no register for data in
// flight.  Cells: quick path in lockstep + deferred queue, exactly as above (same x*).
constexpr int FC_TILE = 256, FC_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Producer-side cursor over the loads of a block: tiles blockIdx.x, blockIdx.x + gridDim.x, ..., pairs pa .. pb-1 of each, in
// consumption order; load number n goes to stage n % FC_STAGES.  It lives in SHARED memory and only thread 0 touches it:
// the hot loop carries none of it in registers.
struct FcCursor {
  long long left;         // loads still to issue
  int tile, gp, st;       // tile, draw pair and stage of the next load
};

template <int NCOL>
__device__ __forceinline__ void fc_issue(FcCursor* c, const ForecastArgs& a, unsigned char* s_rows, uint64_t* s_full, int pa, int pb) {
  constexpr int ROW = NCOL * 8, DRAW_BYTES = FC_TILE * ROW, STAGE_BYTES = 2 * DRAW_BYTES;
  const int st = c->st, gp = c->gp;
  const long long i0 = (long long)c->tile * FC_TILE;
  const uint32_t bytes = (uint32_t)min((long long)FC_TILE, a.N - i0) * ROW;
  const bool v1 = 2ll * gp + 1 < a.n_draws;
  mbar_expect_tx(&s_full[st], v1 ? 2u * bytes : bytes);
  const double* src = a.level1 + ((2ll * gp) * a.N + i0) * NCOL;
  bulk_load(s_rows + (size_t)st * STAGE_BYTES, src, bytes, &s_full[st]);
  if (v1) bulk_load(s_rows + (size_t)st * STAGE_BYTES + DRAW_BYTES, src + a.N * NCOL, bytes, &s_full[st]);
  c->left -= 1;
  if (gp + 1 == pb) { c->gp = pa; c->tile += (int)gridDim.x; } else c->gp = gp + 1;
  c->st = (st + 1 == FC_STAGES) ? 0 : st + 1;
}

// Main pass: every cell's quick path; cells that need more are listed in global memory for k_forecast_deferred.
// No function call in this kernel: it keeps its registers for the loop (64 at 4 blocks per SM, no spill).
template <int NCOL, bool WRITE_X>
__global__ void __launch_bounds__(FC_TILE, NCOL == 4 ? 4 : 3) k_forecast_tma(ForecastArgs a, double* sum_x, double* sum_z,
                                                                              FcQueued* g_list, unsigned long long* g_count,
                                                                              unsigned long long g_cap) {
  extern __shared__ __align__(128) unsigned char fc_smem[];
  constexpr int ROW = NCOL * 8, DRAW_BYTES = FC_TILE * ROW, STAGE_BYTES = 2 * DRAW_BYTES;
  constexpr int QCAP = 96;
  unsigned char* s_rows = fc_smem;                                                   // [FC_STAGES][2][FC_TILE][ROW]
  FcQueued* s_q = reinterpret_cast<FcQueued*>(fc_smem + FC_STAGES * STAGE_BYTES);    // [FC_WARPS][QCAP]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_q + FC_WARPS * QCAP);             // [FC_STAGES]
  uint64_t* s_empty = s_full + FC_STAGES;                                            // [FC_STAGES]
  FcCursor* s_cur = reinterpret_cast<FcCursor*>(s_empty + FC_STAGES);
  const int tid = threadIdx.x, lane = tid & 31;
  const int npairs = (int)((a.n_draws + 1) >> 1);
  const int per = (npairs + gridDim.y - 1) / gridDim.y;
  const int pa = blockIdx.y * per, pb = min(npairs, pa + per);
  const int ntiles = (int)((a.N + FC_TILE - 1) / FC_TILE);
  if (pb <= pa || (int)blockIdx.x >= ntiles) return;
  if (tid == 0) {
    for (int s = 0; s < FC_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], FC_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_cur->left = (long long)((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * (pb - pa);
    s_cur->tile = (int)blockIdx.x; s_cur->gp = pa; s_cur->st = 0;
    for (int s = 0; s < FC_STAGES && s_cur->left > 0; ++s) fc_issue<NCOL>(s_cur, a, s_rows, s_full, pa, pb);
  }
  __syncthreads();
  FcQueued* q = s_q + (tid >> 5) * QCAP;
  const float T_star_f = (float)a.T_star;
  int st = 0;                               // stage of this consumption
  uint32_t par = 0;                         // parity of the stage's current use
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const bool valid = (long long)tile * FC_TILE + tid < a.N;
    const long long ic = valid ? (long long)tile * FC_TILE + tid : a.N - 1;
    const double T = a.T_cal[ic];
    const uint32_t gid = (uint32_t)(a.gid_offset + ic);
    int sx = 0, sz = 0, qn = 0;             // quick-path counts are < 8 per cell: 32 bits hold 2^28 draws
    for (int gp = pa; gp < pb; ++gp) {
      const bool v1 = 2ll * gp + 1 < a.n_draws;
      const uint4 r = philox4x32_10_rk(gid, (uint32_t)gp, 0u, DOM_FORECAST, a.rk);   // independent of the data: before the wait
      mbar_wait(&s_full[st], par);
      const unsigned char* rows = s_rows + (size_t)st * STAGE_BYTES + (size_t)tid * ROW;
      // copy out what the cells need (lambda, tau, z), then hand the stage back: the refill does not wait for the arithmetic
      const double lam0 = *reinterpret_cast<const double*>(rows);
      const double2 tz0 = *reinterpret_cast<const double2*>(rows + 16);
      double lam1 = 0.0;
      double2 tz1 = make_double2(0.0, 0.0);
      if (v1) {
        lam1 = *reinterpret_cast<const double*>(rows + DRAW_BYTES);
        tz1 = *reinterpret_cast<const double2*>(rows + DRAW_BYTES + 16);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      // the producer refills the stage consumed one iteration ago (its readers have long moved on)
      if (tid == 0 && !first && s_cur->left > 0) {
        mbar_wait(&s_empty[st == 0 ? FC_STAGES - 1 : st - 1], st == 0 ? par ^ 1u : par);
        fc_issue<NCOL>(s_cur, a, s_rows, s_full, pa, pb);
      }
      first = false;
      if (++st == FC_STAGES) { st = 0; par ^= 1u; }
      long long* xc = WRITE_X ? a.x_out + (2ll * gp) * a.N + ic : nullptr;
      fc_quick_cell<WRITE_X>((float)lam0, (float)(tz0.x - T), tz0.y > 0.5, T_star_f, r.x, 2u * (uint32_t)gp, (uint32_t)ic, valid, lane, q, qn,
                             sx, sz, xc);
      fc_quick_cell<WRITE_X>((float)lam1, (float)(tz1.x - T), tz1.y > 0.5, T_star_f, r.z, 2u * (uint32_t)gp + 1u, (uint32_t)ic, valid && v1,
                             lane, q, qn, sx, sz, WRITE_X ? xc + a.N : nullptr);
      if (qn >= 32) {                        // full batches only: the remainder stays queued
        __syncwarp();
        const int rem = qn & 31;
        fc_flush(q, rem, qn, g_list, g_count, g_cap, lane);
        qn = rem;
      }
    }
    __syncwarp();
    if (qn > 0) fc_flush(q, 0, qn, g_list, g_count, g_cap, lane);
    if (valid) {
      if (gridDim.y == 1) {
        sum_x[ic] = (double)sx;
        sum_z[ic] = (double)sz;
      } else {
        atomicAdd(&sum_x[ic], (double)sx);
        atomicAdd(&sum_z[ic], (double)sz);
      }
    }
    __syncwarp();
  }
}

// Second pass: one thread per listed cell, the full path (screened inversion of any length, exact fp64 re-decision, PTRS)
// on full warps.  Its x* is added to the customer's sum (integers in doubles: exact, order independent).
template <int NCOL, bool WRITE_X>
__global__ void __launch_bounds__(256) k_forecast_deferred(ForecastArgs a, const FcQueued* g_list, const unsigned long long* g_count,
                                                           unsigned long long g_cap, double* sum_x) {
  const unsigned long long n = min(*g_count, g_cap);
  const PhiloxKey key = seed_key(a.seed);
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
    const FcQueued c = g_list[e];
    const long long ci = c.owner;
    const double* row = a.level1 + ((long long)c.gdraw * a.N + ci) * NCOL;
    const double Tc = a.T_cal[ci];
    const uint32_t gid = (uint32_t)(a.gid_offset + ci);
    const uint4 r = philox4x32_10_rk(gid, c.gdraw >> 1, 0u, DOM_FORECAST, a.rk);
    float lamf, dtf;
    bool al;
    load_row_f32<NCOL>(row, Tc, lamf, dtf, al);
    const long long x = forecast_cell(row, lamf, dtf, al, Tc, a.T_star, (float)a.T_star, (c.gdraw & 1u) ? r.z : r.x,
                                      (c.gdraw & 1u) ? r.w : r.y, gid, c.gdraw, key);
    if (WRITE_X) __stcs(a.x_out + (long long)c.gdraw * a.N + ci, x);
    if (x) atomicAdd(&sum_x[ci], (double)x);
  }
}

// ------------------------------------------------------------------------------------------------
// "next" rows of SURVEY 8f on the draws resident in HBM
// ------------------------------------------------------------------------------------------------
constexpr int SUMMARY_COLS = 10;   // mean lambda, q2.5, q97.5 | capped mean mu, q2.5, q97.5 | mean z | mean tau | mean mu | mean eta

__device__ __forceinline__ void bitonic_sort_shared(double* v, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n_pad; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const double a = v[t], b = v[p];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { v[t] = b; v[p] = a; }
        }
      }
      __syncthreads();
    }
}

// np.percentile(..., method="linear") of a sorted array
__device__ __forceinline__ double percentile_sorted(const double* v, int n, double q) {
  const double pos = q * 0.01 * (double)(n - 1);
  const int lo = (int)floor(pos);
  const int hi = min(lo + 1, n - 1);
  const double t = pos - (double)lo, a = v[lo], b = v[hi];
  return (t < 0.5) ? a + (b - a) * t : b - (b - a) * (1.0 - t);
}

// Per-customer posterior summaries over every resident draw (Table 4 inputs: utils/analysis_bi_helpers.py:75-110,
// post_mean_lambdas/mus :15-27).  One block per customer; its draws are gathered into shared memory and sorted there.
template <int NCOL>
__global__ void __launch_bounds__(256) k_posterior_summary(const double* level1, long long n_tot, long long N, int n_pad,
                                                           double mu_cap, double* out) {
  extern __shared__ double sh[];
  __shared__ double red[8][8];
  for (long long i = blockIdx.x; i < N; i += gridDim.x) {
    double acc[6] = {0, 0, 0, 0, 0, 0};   // lambda, capped mu, z, tau, mu, eta
    for (int pass = 0; pass < 2; ++pass) {          // pass 0: lambda column, pass 1: mu column
      for (long long d = threadIdx.x; d < n_pad; d += blockDim.x) {
        double val = CUDART_INF;
        if (d < n_tot) {
          const double* row = level1 + (d * N + i) * NCOL;
          val = row[pass];
          if (pass == 0) {
            acc[0] += val; acc[2] += row[3]; acc[3] += row[2];
            if (NCOL == 5) acc[5] += row[4];
          } else {
            acc[1] += fmin(val, mu_cap); acc[4] += val;
          }
        }
        sh[d] = val;
      }
      __syncthreads();
      bitonic_sort_shared(sh, n_pad);
      if (threadIdx.x == 0) {
        out[i * SUMMARY_COLS + 3 * pass + 1] = percentile_sorted(sh, (int)n_tot, 2.5);
        out[i * SUMMARY_COLS + 3 * pass + 2] = percentile_sorted(sh, (int)n_tot, 97.5);
      }
      __syncthreads();
    }
    // block reduction of the six sums
    for (int c = 0; c < 6; ++c) {
      double v = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[6] = {0, 0, 0, 0, 0, 0};
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
        for (int c = 0; c < 6; ++c) tot[c] += red[w][c];
      const double inv = 1.0 / (double)n_tot;
      out[i * SUMMARY_COLS + 0] = tot[0] * inv;
      out[i * SUMMARY_COLS + 3] = tot[1] * inv;
      out[i * SUMMARY_COLS + 6] = tot[2] * inv;
      out[i * SUMMARY_COLS + 7] = tot[3] * inv;
      out[i * SUMMARY_COLS + 8] = tot[4] * inv;
      out[i * SUMMARY_COLS + 9] = tot[5] * inv;
    }
    __syncthreads();
  }
}

// Weekly tracking simulation (Figure 2; bivariate/analysis_abe.py:446-464): for every resident draw and every week t,
// sum over customers of Poisson(lambda_i) while birth_i < t <= birth_i + tau_i.  One thread per (customer, draw);
// uniforms: Philox (gid, draw, FC_SLOT_WEEK + w/4, DOM_FORECAST), word w%4, 32-bit; totals are integers => order independent.
template <int NCOL>
__global__ void __launch_bounds__(256) k_weekly_tracking(const double* level1, long long n_tot, long long N, long long gid_offset,
                                                         const double* birth, const double* times, int n_weeks, uint64_t seed,
                                                         unsigned long long* totals) {
  extern __shared__ unsigned long long s_week[];
  const PhiloxKey key = seed_key(seed);
  for (int w = threadIdx.x; w < n_weeks; w += blockDim.x) s_week[w] = 0ull;
  __syncthreads();
  for (long long d = blockIdx.y; d < n_tot; d += gridDim.y)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
      const double* row = level1 + (d * N + i) * NCOL;
      const double lam = __ldcs(row), tau = __ldcs(row + 2);
      const double b0 = birth[i], b1 = b0 + tau;
      const float lamf = (float)lam;
      const uint32_t gid = (uint32_t)(gid_offset + i);
      uint4 r = make_uint4(0, 0, 0, 0);
      int have = -1;
      for (int w = 0; w < n_weeks; ++w) {
        const double t = times[w];
        if (!(t > b0 && t <= b1)) continue;
        if ((w >> 2) != have) { have = w >> 2; r = philox4x32_10(gid, (uint32_t)d, FC_SLOT_WEEK + (uint32_t)have, DOM_FORECAST, key); }
        const uint32_t word = (w & 3) == 0 ? r.x : (w & 3) == 1 ? r.y : (w & 3) == 2 ? r.z : r.w;
        long long inc;
        if (lam >= PTRS_MIN_MEAN) inc = poisson_ptrs(lam, gid, (uint32_t)d, key, FC_SLOT_WEEK_PTRS + 4096u * (uint32_t)w);
        else inc = poisson_inversion_screened(lamf, u24f(word), [&]() { return poisson_inversion(lam, u32d(word)); });
        if (inc) atomicAdd(&s_week[w], (unsigned long long)inc);
      }
    }
  __syncthreads();
  for (int w = threadIdx.x; w < n_weeks; w += blockDim.x)
    if (s_week[w]) atomicAdd(&totals[w], s_week[w]);
}

__global__ void k_scale(double* a, double* b, long long n, double f) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    a[i] *= f;
    b[i] *= f;
  }
}

// ------------------------------------------------------------------------------------------------
// synthetic customers (law of generate_pareto_abe, bi:95-187, without the per-customer event loop)
// ------------------------------------------------------------------------------------------------
struct GenerateArgs {
  long long n, gid_offset;
  int K;
  uint64_t seed;
  double T_cal_lo, T_cal_hi, T_star;
  double beta[16 * 2];
  double Lg[4];        // chol(gamma), lower
  int* x;
  double* t_x;
  double* T_cal;
  double* Xc;          // [(K-1)][n]
  int X_given, T_given; // covariates / T_cal supplied by the caller instead of drawn
  int* x_star;
  double *lam, *mu, *tau;
};

__global__ void __launch_bounds__(256) k_generate(GenerateArgs a) {
  const PhiloxKey key = seed_key(a.seed);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t gid = (uint32_t)(a.gid_offset + i);
    double m0 = a.beta[0], m1 = a.beta[1];
    for (int k = 1; k < a.K; ++k) {                         // X = [1, U(-1,1)^(K-1)]    bi:118-122
      double xk;
      if (a.X_given) xk = a.Xc[(long long)(k - 1) * a.n + i];
      else {
        uint4 r = philox4x32_10(gid, (uint32_t)k, 0u, DOM_GENERATOR, key);
        xk = 2.0 * u53(r.x, r.y) - 1.0;
        a.Xc[(long long)(k - 1) * a.n + i] = xk;
      }
      m0 = fma(xk, a.beta[k * 2 + 0], m0);
      m1 = fma(xk, a.beta[k * 2 + 1], m1);
    }
    double n0, n1;
    normal_pair_u53(philox4x32_10(gid, 0u, 1u, DOM_GENERATOR, key), &n0, &n1);
    const double lam = exp(m0 + a.Lg[0] * n0);                           // theta = exp(X beta + MVN(0, gamma))  bi:133-136
    const double mu = exp(m1 + a.Lg[2] * n0 + a.Lg[3] * n1);
    uint4 r2 = philox4x32_10(gid, 0u, 2u, DOM_GENERATOR, key);
    const double tau = -log(u53(r2.x, r2.y)) / mu;                       // bi:137
    const double T = a.T_given ? a.T_cal[i] : a.T_cal_lo + (a.T_cal_hi - a.T_cal_lo) * u53(r2.z, r2.w);
    uint4 r3 = philox4x32_10(gid, 0u, 3u, DOM_GENERATOR, key);
    uint4 r4 = philox4x32_10(gid, 0u, 4u, DOM_GENERATOR, key);
    const double Teff = fmin(tau, T);
    const long long x = poisson_inversion(lam * Teff, u53(r3.x, r3.y));  // events in (0, min(tau, T_cal)]  bi:149-162
    const double tx = (x > 0) ? Teff * pow(u53(r3.z, r3.w), 1.0 / (double)x) : 0.0;   // last of x uniform event times
    const double hold = fmax(0.0, fmin(tau, T + a.T_star) - T);
    const long long xs = poisson_inversion(lam * hold, u53(r4.x, r4.y));  // bi:172-181
    a.x[i] = (int)x;
    a.t_x[i] = tx;
    a.T_cal[i] = T;
    if (a.x_star) a.x_star[i] = (int)xs;
    if (a.lam) a.lam[i] = lam;
    if (a.mu) a.mu[i] = mu;
    if (a.tau) a.tau[i] = tau;
  }
}

// ------------------------------------------------------------------------------------------------
// issue-rate micro-benchmarks (roofline denominators for an instruction-bound kernel)
// ------------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) k_peak(float* out, int iters, float seedf) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (WHICH == 0) {          // FFMA
    float a0 = seedf + tid, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  } else if (WHICH == 1) {   // IMAD (32-bit)
    uint32_t a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, a4 = tid + 4, a5 = tid + 5, a6 = tid + 6, a7 = tid + 7;
    const uint32_t m = 0xD2511F53u + (uint32_t)seedf;
    for (int i = 0; i < iters; ++i) {
      a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a4;
      a4 = a4 * m + a5; a5 = a5 * m + a6; a6 = a6 * m + a7; a7 = a7 * m + a0;
    }
    out[tid] = (float)(a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7);
  } else if (WHICH == 2) {   // MUFU.EX2
    float a0 = seedf + 1e-3f * tid, a1 = a0 + .1f, a2 = a0 + .2f, a3 = a0 + .3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
    for (int i = 0; i < iters; ++i) {
      a0 = exp2f(-a0); a1 = exp2f(-a1); a2 = exp2f(-a2); a3 = exp2f(-a3);
      a4 = exp2f(-a4); a5 = exp2f(-a5); a6 = exp2f(-a6); a7 = exp2f(-a7);
    }
    out[tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  } else {                   // DFMA
    double a0 = seedf + tid, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[tid] = (float)(a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7);
  }
}

}  // namespace clv
