// clv_kernels.cuh — sm_100a kernels of the Abe (2009/2015) sampler.
//
//   k_sweep<D,MODE>   level-1 sweep: one thread per (chain, customer): z, tau, S MH steps on (log lambda, log mu),
//                     conjugate log eta (D=3), draw write-out, and the customer's share of the level-2
//                     sufficient statistics (int64 fixed point => order-independent sums).
//   k_level2<D>       one warp per chain: reads the reduced statistics, draws Sigma ~ IW and beta | Sigma.
//   k_stats_only      statistics of the current state (first bivariate sweep, set_state).
//   k_derive_params   P = inv(Sigma) etc. from (beta, Sigma) (initial state, set_state).
//
// Reference semantics: src/models/bivariate/mcmc.py ("bi") and src/models/trivariate/mcmc.py ("tri")
// of lucagem29/mcmc_clv_model; the line ranges are cited at each block.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "clv_rng.cuh"
#include "clv_forecast.cuh"

namespace clv {

constexpr int MAXK = 16;
constexpr int MAXD = 3;
constexpr int NSTAT_MAX = MAXK * MAXD + 6;
constexpr int SWEEP_THREADS = 128;
constexpr int P2P_MAX_WORLD = 16;
// Resident blocks of 128 threads per SM for the one-customer kernels (k_sweep, k_persistent).  Round 1 ran them at 8 (64
// registers: 25 % faster than 80 when this was the kernel of the 10 M-customer sweep, profiles/r01_kernel_ab.txt).  They now
// serve the small problems (customers x chains < 100 000: at most 5 blocks per SM exist), where a sweep is bound by the
// latency of one customer's dependent steps and the spills of the 64-register build sit on that path: 4 blocks per SM
// (<= 128 registers, no spill) is 3-6 % faster there (C1 13.2 -> 12.4, C2 17.4 -> 16.6, C3 19.2 -> 18.6 us per sweep).
#ifndef CLV_MINBLOCKS
#define CLV_MINBLOCKS 4
#endif

enum : int { MODE_FAST = 0, MODE_STRICT = 1, MODE_INJECT = 2 };

// Per-chain level-2 state, written by k_level2 / k_derive_params, read by k_sweep.
struct ChainParams {
  double beta[MAXK * MAXD];  // [k*D + d]
  double Sigma[MAXD * MAXD];
  double P00, P01, P11;      // entries of inv(Sigma) used by the level-1 target (bi:303-305; tri:419-426, Q4)
  double eta_post_var, eta_sd;  // tri:325-326
  // the level-1 quadratic form as a sum of two squares, P/2 = R'R with R = [[qa, qb], [0, qc]], and the inverse of the 2x2
  // block [[P00, P01], [P01, P11]] (s00, s01, s11): what the fp32-screened Metropolis step works with (E32, below)
  double qa, qb, qc, s00, s01, s11;
  int status;                // 0 ok, 1 scale matrix not positive definite / non-finite
  int pad;
};

// Run constants (host-computed once per init).
struct ModelConst {
  int D, K, S, compat;
  long long N, N_global, gid_offset;
  double center[MAXD];       // centring constants c_d (prior intercepts): statistics are of y - c
  double fx_scale, fx_inv;   // fixed-point scale of the level-2 statistics
  double ll_scale, ll_inv;   // fixed-point scale of the log-likelihood sums
  double V[MAXK * MAXK];     // (X'X + A0)^-1                                  bi:249
  double LV[MAXK * MAXK];    // chol(V), lower
  double A0B0c[MAXK * MAXD]; // A0 (B0 - e0 c')
  double Q0[MAXD * MAXD];    // S0 + B0c' A0 B0c
  double B0c[MAXK * MAXD];
  double nu_n;               // nu0 + N_global                                   bi:256
  double omega2;             // tri:494
};

// Fused forecast (SURVEY 8a a8 on the sampler's own kept draws): when a draw is kept, lambda, tau and z of the cell are in
// registers -- x* is simulated there and added to per-(chain, customer) sums, so the posterior-predictive means need no
// second pass over the draws at all (zero re-read; the draws need not even be stored).  Same Philox counters as
// clv_forecast_resident on the stored draws (global draw index = chain * n_draws + draw), hence the same x*.
struct FusedForecast {
  unsigned long long* sum_x;   // [chains][N]
  unsigned int* sum_z;         // [chains][N]
  double T_star;
  uint64_t seed;
  PhiloxRoundKeys rk;          // round keys of `seed`
};

__device__ __noinline__ void fused_forecast_cell(const FusedForecast* f, uint32_t gid, long long gdraw, double lam, double tau,
                                                 double zf, double T, long long idx) {
  const uint4 r = philox4x32_10_rk(gid, (uint32_t)(gdraw >> 1), 0u, DOM_FORECAST, f->rk);
  const uint32_t ua = (gdraw & 1) ? r.z : r.x, ub = (gdraw & 1) ? r.w : r.y;
  const double m = lam * future_horizon(T, tau, zf, f->T_star);
  long long x;
  if (m >= PTRS_MIN_MEAN) x = poisson_ptrs(m, gid, (uint32_t)gdraw, seed_key(f->seed));
  else x = poisson_inversion_screened((float)m, u24f(ua), [&]() { return poisson_inversion(m, u53(ua, ub)); });
  if (x) f->sum_x[idx] += (unsigned long long)x;
  if (zf > 0.5) f->sum_z[idx] += 1u;
}

// per customer: sums over the chains, scaled to means
__global__ void k_fused_forecast_fold(const unsigned long long* sx, const unsigned int* sz, long long N, int chains, double inv,
                                      double* mean_x, double* p_alive) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long tx = 0;
    unsigned long long tz = 0;
    for (int c = 0; c < chains; ++c) { tx += sx[(long long)c * N + i]; tz += sz[(long long)c * N + i]; }
    mean_x[i] = (double)tx * inv;
    p_alive[i] = (double)tz * inv;
  }
}

struct SweepArgs {
  const ModelConst* mc;
  const ChainParams* params;   // [chains]
  // data (SoA)
  const int* x;
  const double* t_x;
  const double* T_cal;
  const double* Xc;            // [(K-1)][N] covariate columns (intercept implicit)
  const double* log_s;
  // state [chains][N]
  double* ll;
  double* lm;
  double* le;
  double* z;                   // last z, tau (kept for get_state / injected parity)
  double* tau;
  // level-2 statistics accumulators [chains][NSTAT_MAX] and log-lik sums [chains][n_draws]
  unsigned long long* acc;
  long long* loglik_acc;
  long long loglik_stride;     // n_draws of the current run
  // draws of the current chunk [chains][chunk_cap][N][ncol]
  double* draws;
  long long chunk_cap;
  long long slot;              // slot inside the chunk, -1: not kept
  long long draw_index;        // index of this draw in the run (for loglik_acc)
  uint32_t sweep;              // 1-based sweep number (Philox counter)
  uint32_t chain_offset;
  uint64_t seed;
  PhiloxRoundKeys rk;          // round keys of `seed` (launch constants)
  int store_zt;                // write z/tau state arrays
  const FusedForecast* fc;     // nullable: simulate x* of every kept draw from the registers (clv_set_fused_forecast)
  const int* error_flag;       // device error flag (2: a peer rank stopped)
  int pdl_early;               // 1: let the next kernel of the stream become resident at once, 0: once this block's tiles are done
  // k_sweep2 with dynamic tiles: blocks draw tile numbers from a counter; tiles [0, n_big) hold 256 customers (two per
  // thread), tiles [n_big, n_big + n_small) the remaining customers 128 at a time (one per thread): short tiles last
  unsigned int* tile_counter;  // [2][chains], alternating by sweep parity (nullable: static grid-stride tiles)
  long long n_big, n_small;
  // injected variates (MODE_INJECT)
  const double *u_z, *e_tau, *u_tau, *t3_l, *t3_m, *u_acc, *n_eta;
};

// Table exp: 2^(j / EXP_N), j < EXP_N, in shared memory + a polynomial for the remainder.  CLV_EXP_BITS = 8 (default):
// 256 entries (2 KB per block), degree-4 remainder (|r| <= ln2/512: r^5/120 < 4e-17, <= 1 ulp); CLV_EXP_BITS = 6: 64
// entries, degree 5 -- one DFMA more per exp (measured: 8 is 1.2 % faster on the 10 M-customer sweep, profiles/r02_kernel_ab.txt).
#ifndef CLV_EXP_BITS
#define CLV_EXP_BITS 8
#endif
constexpr int EXP_BITS = CLV_EXP_BITS, EXP_N = 1 << EXP_BITS;
static_assert(EXP_BITS == 6 || EXP_BITS == 8, "CLV_EXP_BITS must be 6 or 8");
__constant__ double c_exptab[EXP_N];   // 2^(j/EXP_N), uploaded by clv_create
// exp_tab constants: EXP_N/ln2, -ln2/EXP_N (high part with 16 zero low bits, low part), 1/120, 1/24, 1/6
__constant__ double c_expk[6] = {EXP_BITS == 6 ? 92.332482616893656877 : 369.32993046757462751,
                                 EXP_BITS == 6 ? -0x1.62e42fefa0000p-7 : -0x1.62e42fefa0000p-9,
                                 EXP_BITS == 6 ? -0x1.cf79abc9e3b3ap-46 : -0x1.cf79abc9e3b3ap-48,
                                 8.3333333333333332e-03, 4.1666666666666664e-02, 1.6666666666666666e-01};

// Programmatic dependent launch (stream mode): a kernel launched with the programmatic-stream-serialization attribute
// may become resident while its predecessor still runs.  pdl_launch_dependents() lets the NEXT kernel of the stream
// start its prologue; pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are visible.  Both are
// no-ops for an ordinary launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ long long to_fx(double v, double scale) { return __double2ll_rn(v * scale); }

// Warp-wide int64 sum with three REDUX.SUM (32-bit) instead of a 64-bit shuffle tree:
// v = lo + mid 2^26 + hi 2^52 (hi signed), and 32 pieces of 26 bits cannot overflow 32 bits.  All 32 lanes must call.
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#ifdef CLV_SHFL_SUM
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
#endif
  const unsigned lo = (unsigned)v & 0x3ffffffu, mid = (unsigned)(v >> 26) & 0x3ffffffu;
  const int hi = (int)(v >> 52);
  const unsigned slo = __reduce_add_sync(0xffffffffu, lo), smid = __reduce_add_sync(0xffffffffu, mid);
  const int shi = __reduce_add_sync(0xffffffffu, hi);
  return (long long)slo + ((long long)smid << 26) + (long long)((unsigned long long)(long long)shi << 52);
}

// exp(x) for |x| <= 700 (in particular the clip range +-70 of bi:323-324), ~1 ulp, branch-free:
//   n = rint(x E/ln2), r = x - n ln2/E (|r| <= ln2/2E), exp(x) = 2^(n>>B) * 2^((n&(E-1))/E) * (1 + r + ... ),  E = 2^B
// 2^(j/E) comes from the shared-memory table; E = 64 with a degree-5 remainder (< 4e-17), or E = 256 with degree 4.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  // the fp64 constants sit in the constant bank so that DFMA reads them as c[][] operands; written as literals ptxas
  // rebuilds each one with two UMOVs on every call (26 extra issue slots per MH step)
  const double t = fma(x, c_expk[0], 6755399441055744.0);                 // 64/ln2 ; 1.5 * 2^52 rounds to nearest
  const int n = __double2loint(t);
  const double nd = t - 6755399441055744.0;
  double r = fma(nd, c_expk[1], x);                                       // -ln2/64, high part (low 16 bits zero: n*hi exact)
  r = fma(nd, c_expk[2], r);                                              // low part
  double p = (EXP_BITS == 6) ? fma(fma(r, c_expk[3], c_expk[4]), r, c_expk[5]) : fma(r, c_expk[4], c_expk[5]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;
  const double T = tab[n & (EXP_N - 1)];
  double v = fma(T, p, T);
  const int hi = __double2hiint(v) + ((n >> EXP_BITS) << 20);
  return __hiloint2double(hi, __double2loint(v));
}

// The same with the table given as a 32-bit shared-window address held in a register: the load is [reg] instead of
// [reg + uniform base], and the base (S2UR CgaCtaId / UMOV / UIADD3 / ULEA, which ptxas rebuilds in every loop trip)
// disappears from the Metropolis loop.
__device__ __forceinline__ double exp_tab(double x, uint32_t tab_addr) {
  const double t = fma(x, c_expk[0], 6755399441055744.0);
  const int n = __double2loint(t);
  const double nd = t - 6755399441055744.0;
  double r = fma(nd, c_expk[1], x);
  r = fma(nd, c_expk[2], r);
  double p = (EXP_BITS == 6) ? fma(fma(r, c_expk[3], c_expk[4]), r, c_expk[5]) : fma(r, c_expk[4], c_expk[5]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_addr + ((uint32_t)(n & (EXP_N - 1)) << 3)));
  double v = fma(T, p, T);
  const int hi = __double2hiint(v) + ((n >> EXP_BITS) << 20);
  return __hiloint2double(hi, __double2loint(v));
}

// exp for any argument: the table path where it is valid (results stay normal numbers), libm beyond
__device__ __noinline__ double exp_slow(double x) { return exp(x); }
__device__ __forceinline__ double exp_any(double x, const double* __restrict__ tab) {
  return (fabs(x) <= 700.0) ? exp_tab(x, tab) : exp_slow(x);
}

// Level-1 target, bi:291-310, without the lm > 5 cut (the caller applies it).  Tz = z*T_cal + (1-z)*tau, omz = 1-z,
// ll, lm in [-70, 70].  The quadratic form takes the precision pre-scaled per chain, h00 = -P00/2, h01 = -P01,
// h11 = -P11/2, and is folded into the likelihood by Horner steps: 7 fp64 instructions instead of 10.
template <typename Tab>
__device__ __forceinline__ double log_post_open(double ll, double lm, double xd, double omz, double Tz, double m0,
                                                double m1, double h00, double h01, double h11, Tab tab) {
  const double dl = ll - m0, dm = lm - m1;
  const double lik = xd * ll + omz * lm - (exp_tab(ll, tab) + exp_tab(lm, tab)) * Tz;
  return fma(dl, fma(h00, dl, h01 * dm), fma(dm, h11 * dm, lik));
}
// bi:308-309: the target is -inf where log mu > 5
__device__ __forceinline__ double log_post(double ll, double lm, double xd, double omz, double Tz, double m0,
                                           double m1, double h00, double h01, double h11, const double* tab) {
  const double res = log_post_open(ll, lm, xd, omz, Tz, m0, m1, h00, h01, h11, tab);
  return (lm > 5.0) ? -CUDART_INF : res;
}

// MH accept rule of bi:329-330: exp(prop - cur) > u, NaN compares false.  An fp32 SFU screen in log space
// (uf = fp32 image of u) settles all but ~1e-5 of the decisions; the tie zone is re-decided with the exact fp64
// expression, so the outcome always equals `exp(d) > u`.
template <bool GUARD_U, typename ExactU>
__device__ __forceinline__ bool mh_accept(double d, float uf, ExactU exact_u) {
  // one rarely taken branch: accept / reject are decided branch-free unless d - ln u falls in the guard band
  // (d >= 0 gives df >= 0 > ln u; d = -inf or d < -80 gives df < ln u - tol since ln u >= -23; NaN fails both tests)
  const float df = (float)d;
  const float lu = 0.69314718055994531f * lg2_ftz(uf);
  const float tol = 1e-5f * (1.0f + fabsf(lu));
  bool acc = df > lu + tol;
  const bool rej = df < lu - tol;
  bool sure = acc || rej;
  if (GUARD_U) sure = sure && uf > 1e-30f;       // injected uniforms may be 0 or denormal
  if (!sure) acc = (d >= 0.0) || exp(d) > exact_u();
  return acc;
}

// ---- the Metropolis decision with the exponential part of the target in fp32 (CLV_E32, FAST mode) ------------------------
// target(ll, lm) = x ll + (1-z) lm - d'Pd/2 - (e^ll + e^lm) Tz,  d = (ll, lm) - m   (bi:291-310).  Its fp64 evaluation costs
// two table exps (18 DFMA-class instructions, 2 shared-memory loads, ~8 integer ones) plus 7 for the rest, and keeps 16
// registers of per-customer constants alive in the loop.  The decision `exp(prop - cur) > u` only needs prop - cur to the
// accuracy that separates it from ln u, so the hot path works on
//     target = const - Q - E,   Q = (qa ll + qb lm - k0)^2 + (qc lm - k1)^2   (fp64, 5 instructions),   E = (e^ll + e^lm) Tz  (fp32),
// where Q is the linear and the quadratic part with the square completed: P/2 = R'R, R = [[qa, qb], [0, qc]] (ChainParams),
// (k0, k1) = R m', m' = m + P^-1 (x, 1-z) -- two constants per customer instead of four, and no fp64 multiply by x.  E comes
// from the SFU (ex2.approx, 2 ulp) together with a RIGOROUS bound of its error, and the step is decided when the distance
// of prop - cur from ln u exceeds the sum of the bounds.  Otherwise (~1e-4 of the steps) mh_exact() re-decides with the
// fp64 expression log_post_open on the fp64 state, whose inputs wait in shared memory (stash), not in registers.  Every
// decision therefore equals the one of the all-fp64 evaluation: the chain is bit-identical to CLV_E32=0 (same digest).
//   error budget of the screen's value of prop - cur - ln u against the fp64 expression (eps = 2^-53):
//   * E32 = (ex2(a L2E) + ex2(b L2E)) Tz32, a = (float)ll, b = (float)lm: ex2.approx 2^-22 relative; a, L2E and their
//     product each rounded (3 x 2^-24 relative in the argument = 1.8e-7 |a| relative in the result); the sum, Tz32 and the
//     product 3 x 2^-24: <= 4.2e-7 + 1.8e-7 max(|a|, |b|).  dE = Ep - Ec and dt = dL32 - dE add 2^-24 (|dE| + |dt|) <=
//     1.2e-7 (Ep + Ec) + 6e-8 |dL32|.  Carried: tolE = E (1.3e-6 + 4e-7 max(|a|, |b|)) (margin > 2), 5e-7 E for the
//     initial E (an fp64 value rounded once); dL32 = (float)(Qc - Qp): 6e-8 |dL32|; carried 3e-7 |dL32|.
//   * fp64 rounding of Q and of the exact expression itself.  u = qa ll + qb lm - k0 cancels terms of size up to
//     Tm = (|qa| + |qb| + |qc|)(70 + |m'0| + |m'1|): |du| <= eps (Tm + |u|), plus an offset common to prop and cur from the
//     rounding of k (<= 4 eps Tm).  Within a sweep the current state obeys |u|, |v| <= B = sqrt(Q0 + E0 + 25 S): an
//     accepted step raises Q + E by at most -ln u + tol < 24.5 (u >= 2^-33).  With Dm = max |u_p - u_c|: the error of
//     Qc - Qp is <= 24 eps (Tm + B + Dm)(B + Dm); the exact expression's own rounding is <= 6 eps (70 x + 70 + E + 2 q),
//     q = d'Pd/2 <= 2 (B + Dm)^2 + 2 Rm^2, Rm = |R P^-1 (x, 1-z)| (its E part is covered by tolE).  For Dm <= 3 B both
//     are below 100 eps (Tm + 4B + Rm)(4B + Rm) + 420 eps x, which e32_begin requires to be < 2e-6 (else the customer's
//     Tz32 is NaN for this sweep and the exact path decides every step: thousands of transactions, a nearly singular
//     Sigma); for Dm > 3 B, |Qc - Qp| >= 0.22 Dm^2 and both errors are below 3e-7 |Qc - Qp| as long as Tm < 1e8 (also
//     required) -- the relative term of the tolerance.  The 2e-6 sits in the floor 1.2e-5, which keeps the
//     1e-5 (1 + |ln u|) of mh_accept for lg2.approx and the fp32 image of u.
//   * Overflow / NaN anywhere make the comparison of the screen false and so lead to the exact path; so does a proposal
//     with max(|a|, |b|) >= 69.5, which the exact path clips to +-70 (bi:323-324) -- the hot path never clips.
#ifndef CLV_E32
#define CLV_E32 1
#endif
// (ex2_ftz: clv_forecast.cuh)
constexpr int E32_STASH = 5;      // doubles per customer in the stash: x, 1-z, Tz, m0, m1
__device__ __forceinline__ double q_part(double ll, double lm, double qa, double qb, double qc, double k0, double k1) {
  const double u = fma(qa, ll, fma(qb, lm, -k0)), v = fma(qc, lm, -k1);
  return fma(u, u, v * v);
}
// fp32 image of E at (ll, lm), the bound of its error, and max(|a|, |b|)
__device__ __forceinline__ void e_part32(double ll, double lm, float Tz32, float& E, float& tolE, float& amax) {
  const float a = (float)ll, b = (float)lm;
  E = (ex2_ftz(a * 1.44269504088896341f) + ex2_ftz(b * 1.44269504088896341f)) * Tz32;
  amax = fmaxf(fabsf(a), fabsf(b));
  tolE = E * fmaf(amax, 4e-7f, 1.3e-6f);
}
// the exact decision (bi:329-330 on the fp64 target) from the fp64 state; cur is -inf where log mu > 5 (bi:308-309).
// stash: this customer's x, 1-z, Tz, m0, m1 at stride SWEEP_THREADS doubles.
__device__ __noinline__ bool mh_exact(double ll, double lm, double pl, double pm, const double* stash, const ChainParams* cp,
                                      const double* tab, uint32_t ur) {
  const double xd = stash[0], omz = stash[SWEEP_THREADS], Tz = stash[2 * SWEEP_THREADS], m0 = stash[3 * SWEEP_THREADS],
               m1 = stash[4 * SWEEP_THREADS];
  const double h00 = -0.5 * cp->P00, h01 = -cp->P01, h11 = -0.5 * cp->P11;
  const double cur = (lm > 5.0) ? -CUDART_INF : log_post_open(ll, lm, xd, omz, Tz, m0, m1, h00, h01, h11, tab);
  const double prop = log_post_open(pl, pm, xd, omz, Tz, m0, m1, h00, h01, h11, tab);
  const double d = prop - cur;
  return (d >= 0.0) || exp(d) > u32d(ur);
}
// The state of one customer inside the Metropolis loop (13 registers) ...
struct E32State {
  double ll, lm, Qc, k0, k1;
  float Ec, tEc, Tz32;
};
// ... its set-up: stores the exact path's inputs, completes the square, evaluates the current target's parts
__device__ __forceinline__ void e32_begin(E32State& st, double ll, double lm, double lam, double mu, double xd, double omz,
                                          double Tz, double m0, double m1, const ChainParams& cp, int S, double* stash) {
  stash[0] = xd; stash[SWEEP_THREADS] = omz; stash[2 * SWEEP_THREADS] = Tz; stash[3 * SWEEP_THREADS] = m0;
  stash[4 * SWEEP_THREADS] = m1;
  const double qa = cp.qa, qb = cp.qb, qc = cp.qc;
  const double d0 = fma(cp.s00, xd, cp.s01 * omz), d1 = fma(cp.s01, xd, cp.s11 * omz);      // P^-1 (x, 1-z)
  const double n0 = m0 + d0, n1 = m1 + d1;                                                     // m'
  st.ll = ll; st.lm = lm;
  st.k0 = fma(qa, n0, qb * n1);
  st.k1 = qc * n1;
  st.Qc = q_part(ll, lm, qa, qb, qc, st.k0, st.k1);
  st.Ec = (lm > 5.0) ? CUDART_NAN_F : (float)((lam + mu) * Tz);     // NaN: target -inf, the exact path decides
  st.tEc = st.Ec * 5e-7f;
  // the fp64 rounding guard of the error budget (fp32 arithmetic with 1 % margins; NaN / inf fail it)
  const float Tm = 1.01f * (float)((fabs(qa) + fabs(qb) + fabs(qc)) * (70.0 + fabs(n0) + fabs(n1)));
  const float Rm = 1.01f * (float)(fabs(fma(qa, d0, qb * d1)) + fabs(qc * d1));
  const float B4 = 4.04f * sqrtf((float)st.Qc + st.Ec + 25.0f * (float)S);
  const float W2 = B4 + Rm, W1 = Tm + W2;
  const bool ok = fmaf(100.0f * W1, W2, 420.0f * fabsf((float)xd)) < 1.8e10f && Tm < 1e8f;
  st.Tz32 = ok ? (float)Tz : CUDART_NAN_F;
  asm volatile("" : "+f"(st.Tz32));             // opaque: ptxas would otherwise rebuild it inside the loop
}
// one Metropolis step (bi:312-335) in two halves, so that a caller with several customers per thread can interleave them
struct E32Prop {
  double pl, pm, Qp;
  float Ep, tEp, amax;
};
__device__ __forceinline__ void e32_propose(const E32State& st, E32Prop& pr, float tl32, float tm32, double s_l, double s_m,
                                            double qa, double qb, double qc) {
  const double tl = (double)tl32, tm = (double)tm32;
  pr.pl = st.ll + s_l * tl;                     // bi:318-324 (clip: in the exact path)
  pr.pm = st.lm + s_m * tm;
  pr.Qp = q_part(pr.pl, pr.pm, qa, qb, qc, st.k0, st.k1);
  e_part32(pr.pl, pr.pm, st.Tz32, pr.Ep, pr.tEp, pr.amax);
}
__device__ __forceinline__ void e32_decide(E32State& st, E32Prop& pr, uint32_t ur, double qa, double qb, double qc,
                                           const double* stash, const ChainParams* cp, const double* tab) {
  const float dL32 = (float)(st.Qc - pr.Qp);
  const float dt = dL32 - (pr.Ep - st.Ec);
  const float lg = lg2_ftz(u32f(ur));
  const float gap = fmaf(lg, -0.69314718055994531f, dt);                       // prop - cur - ln u
  float tol = (pr.tEp + st.tEc) + fmaf(fabsf(dL32), 3e-7f, 1.2e-5f);
  tol = fmaf(fabsf(lg), 7e-6f, tol);                                            // 1e-5 |ln u|
  bool acc = gap > 0.0f;
  if (!(fabsf(gap) > tol && pr.amax < 69.5f)) {                                 // NaN / inf anywhere: not decided here
    pr.pl = fmin(fmax(pr.pl, -70.0), 70.0);
    pr.pm = fmin(fmax(pr.pm, -70.0), 70.0);
    pr.Qp = q_part(pr.pl, pr.pm, qa, qb, qc, st.k0, st.k1);
    e_part32(pr.pl, pr.pm, st.Tz32, pr.Ep, pr.tEp, pr.amax);
    acc = mh_exact(st.ll, st.lm, pr.pl, pr.pm, stash, cp, tab, ur);
  }
  // a proposal with log mu > 5 has target -inf and is never accepted (bi:308-309); cur itself can be -inf only at the
  // start (Ec = NaN), and then every admissible proposal is accepted by the exact path (d = +inf)
  if (acc && !(pr.pm > 5.0)) {
    st.ll = pr.pl; st.lm = pr.pm; st.Qc = pr.Qp; st.Ec = pr.Ep; st.tEc = pr.tEp;
  }
}

// np.clip(v, -70, 70) of bi:323-324 for both proposals; the test runs on the high words so the common case costs a few
// integer ops and one rarely taken branch (|v| >= 70 needs a t3 variate beyond ~47: about 1 proposal in 50,000).
// Tried and dropped: leaving the loop for a clipping copy of the step instead of the call -- ptxas merges the copies
// back into one loop and reloads the Philox keys through vector registers (spills).
__device__ __noinline__ double clip70_slow(double v) { return fmin(fmax(v, -70.0), 70.0); }   // by value: no stack traffic
__device__ __forceinline__ void clip70_pair(double& a, double& b) {
  const unsigned ha = (unsigned)__double2hiint(a) & 0x7fffffffu, hb = (unsigned)__double2hiint(b) & 0x7fffffffu;
  if (max(ha, hb) >= 0x40518000u) {
    a = clip70_slow(a);
    b = clip70_slow(b);
  }
}

// Level-2 sufficient statistics.  Every thread owns one column of the dynamic shared array s_priv[(stat)][128]
// (int64 fixed point) and adds its customers' terms there, tile after tile; flush_stats() reduces the columns once per
// block and sweep.  All sums are integers => the result does not depend on any of this.
// XC_STASH: the first covariates of a customer wait in shared memory between the prologue of its tile (where they enter the
// prior means) and the statistics at its end, instead of being fetched from global memory twice (k_sweep2)
constexpr int XC_STASH = 4;
template <int D>
__device__ __forceinline__ void accumulate_stats(const ModelConst& mc, const double* __restrict__ Xc, long long N,
                                                 long long i, bool valid, double yc0, double yc1, double yc2,
                                                 long long* s_priv, const double* xs = nullptr, int xs_stride = 0) {
  if (!valid) return;
  // to_fx(xk y, sc) = rn((xk y) sc): sc is a power of two, so rn(xk (y sc)) is the same integer -- the responses are
  // scaled once per customer instead of once per term
  const double sc = mc.fx_scale;
  const int K = mc.K;
  const double y[3] = {yc0, yc1, yc2};
  const double ys[3] = {yc0 * sc, yc1 * sc, yc2 * sc};
  long long* col = s_priv + threadIdx.x;
  // (a 64-bit shared-memory atomicAdd would be one instruction in source but compiles to an ATOMS.CAS loop: plain
  // load / add / store on the thread's own column it is)
  for (int k = 0; k < K; ++k) {
    const double xk = (k == 0) ? 1.0 : (xs && k <= XC_STASH) ? xs[(k - 1) * xs_stride] : Xc[(long long)(k - 1) * N + i];
#pragma unroll
    for (int d = 0; d < D; ++d) col[(k * D + d) * SWEEP_THREADS] += __double2ll_rn(xk * ys[d]);
  }
  int t = K * D;
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int e = d; e < D; ++e) {
      col[t * SWEEP_THREADS] += __double2ll_rn(ys[d] * y[e]);
      ++t;
    }
}

// columns 0..nstat-1: statistics, column nstat: log-likelihood sum of a kept draw.  Adds the block totals to s_acc
// (slot NSTAT_MAX for the log-likelihood) and clears the columns.
__device__ __forceinline__ void flush_stats(long long* s_priv, unsigned long long* s_acc, int nstat, bool with_loglik) {
  const int lane = threadIdx.x & 31;
  long long* col = s_priv + threadIdx.x;
  const int ncol = nstat + (with_loglik ? 1 : 0);
  for (int t = 0; t < ncol; ++t) {
    const long long v = warp_sum_ll(col[t * SWEEP_THREADS]);
    col[t * SWEEP_THREADS] = 0;
    if (lane == 0 && v) atomicAdd(&s_acc[t == nstat ? NSTAT_MAX : t], (unsigned long long)v);
  }
}

__device__ __forceinline__ void clear_stats(long long* s_priv, int nstat) {
  for (int t = 0; t <= nstat; ++t) s_priv[t * SWEEP_THREADS + threadIdx.x] = 0;
}

// Per-sweep scalars (kernel arguments in stream mode, computed on the device in persistent mode).
struct SweepStep {
  uint32_t sweep;        // 1-based sweep number (Philox counter)
  int keep;              // this sweep is a kept draw
  int store_zt;          // write z / tau state arrays
  long long slot;        // slot inside the draw chunk
  long long chunk_cap;
  long long draw_index;  // index of the kept draw within the run
  double* draws;         // nullable
};

// One tile of 128 customers of one chain through one Gibbs sweep (blocks a1-a3, a5 of SURVEY 8a) plus its share of
// the level-2 statistics.  cp / s_beta / s_tab / s_acc live in shared memory.
template <int D, int MODE, bool FUSE_FC = false>
__device__ __forceinline__ void sweep_tile(const SweepArgs& a, const ModelConst& mc, const ChainParams& cp,
                                           const double* s_beta, const double* s_tab, long long* s_priv,
                                           const SweepStep& sw, int chain, long long tile, uint32_t c3, double* s_stash) {
  const int tid = threadIdx.x;
  const int K = mc.K, S = mc.S;
  const long long N = mc.N;
  const long long cN = (long long)chain * N;
  const double h00 = -0.5 * cp.P00, h01 = -cp.P01, h11 = -0.5 * cp.P11;
  // proposal scales are variances (bi:316-317, Q2); t3_fast returns t / sqrt(3)
  const double t3s = (MODE == MODE_FAST) ? 1.7320508075688772 : 1.0;
  const double s_l = cp.Sigma[0] * t3s, s_m = cp.Sigma[D + 1] * t3s;
  const bool keep = sw.keep != 0;
  const long long i = tile * SWEEP_THREADS + tid;
  const bool valid = i < N;
  double yc0 = 0.0, yc1 = 0.0, yc2 = 0.0, lik = 0.0;
  if (valid) {
    const uint32_t gid = (uint32_t)(mc.gid_offset + i);
    const double xd = (double)a.x[i];
    const double tx = a.t_x[i], T = a.T_cal[i];
    double ll = a.ll[cN + i], lm = a.lm[cN + i];
    // prior means (X beta)[i, :]   bi:284
    double m0 = s_beta[0], m1 = s_beta[1], m2 = (D == 3) ? s_beta[2] : 0.0;
    for (int k = 1; k < K; ++k) {
      double xk = a.Xc[(long long)(k - 1) * N + i];
      m0 = fma(xk, s_beta[k * D + 0], m0);
      m1 = fma(xk, s_beta[k * D + 1], m1);
      if (D == 3) m2 = fma(xk, s_beta[k * D + 2], m2);
    }
    // ---- z (bi:193-200) and tau (bi:203-227) from the current lambda, mu -------------------------
    const double lam = exp_tab(ll, s_tab), mu = exp_tab(lm, s_tab);   // |ll|, |lm| <= 70 by construction
    double uz, ut, et;
    if (MODE == MODE_INJECT) {
      uz = a.u_z[cN + i];
      ut = a.u_tau[cN + i];
      et = a.e_tau[cN + i];
    } else {
      uint4 r = philox4x32_10_rk(gid, sw.sweep, 0u, c3, a.rk);
      uz = u53(r.x, r.y);
      ut = u53(r.z, r.w);
      et = 0.0;
    }
    const double ml = mu + lam;
    const double e = exp_any(-(ml * (T - tx)), s_tab);
    const double pa = (ml * e) / (ml * e + mu * (1.0 - e));
    const bool alive = uz < pa;
    double tau;
    if (MODE == MODE_INJECT) {
      if (alive) {
        tau = T + (1.0 / mu) * et;
      } else {
        double mtx = fmin(700.0, ml * tx), mT = fmin(700.0, ml * T);
        tau = -log((1.0 - ut) * exp(-mtx) + ut * exp(-mT)) / ml;
      }
    } else {
      // both cases share one logarithm and one division (no divergent branches): alive -> T - ln(u)/mu (an Exp(mu)
      // beyond T_cal, bi:217); churned -> -ln((1-u) e^{-ml t_x} + u e^{-ml T}) / ml (bi:223-226)
      const double mtx = fmin(700.0, ml * tx), mT = fmin(700.0, ml * T);
      const double mix = (1.0 - ut) * exp_tab(-mtx, s_tab) + ut * exp_tab(-mT, s_tab);
      const double lg = -log(alive ? ut : mix);
      tau = (alive ? T : 0.0) + lg / (alive ? mu : ml);
    }
    const double zf = alive ? 1.0 : 0.0;
    const double omz = 1.0 - zf;
    const double Tz = alive ? T : tau;          // z*T_cal + (1-z)*tau, bi:298
    // ---- S Metropolis steps (bi:312-335) -------------------------------------------------------
    constexpr bool E32 = (CLV_E32 != 0) && MODE == MODE_FAST;   // fp32 E part + exact re-decision (e32_decide)
    if constexpr (E32) {
      E32State st;
      double* stash = s_stash + tid;
      e32_begin(st, ll, lm, lam, mu, xd, omz, Tz, m0, m1, cp, S, stash);
      const double qa = cp.qa, qb = cp.qb, qc = cp.qc;
      for (int s = 0; s < S; ++s) {
        const uint4 A = philox4x32_10_rk(gid, sw.sweep, 1u + (uint32_t)s, c3, a.rk);   // word layout: clv_rng.cuh
        E32Prop pr;
        e32_propose(st, pr, t3_fast(A.x, A.y), t3_fast(A.z, A.w), s_l, s_m, qa, qb, qc);
        e32_decide(st, pr, low_bytes(A.x, A.y, A.z, A.w), qa, qb, qc, stash, &cp, s_tab);
      }
      ll = st.ll;
      lm = st.lm;
    } else {
    double cur = log_post(ll, lm, xd, omz, Tz, m0, m1, h00, h01, h11, s_tab);
    uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(s_tab);
    asm volatile("" : "+r"(tab_addr));                 // opaque: keeps ptxas from rebuilding it inside the loop
    // one step from the four words of its Philox block (ignored when the variates are injected)
    auto mh_step = [&](int s, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
      double tl, tm, ua = 0.0;
      float uaf;
      uint32_t ur = 0u;
      if (MODE == MODE_INJECT) {
        long long o = ((long long)chain * S + s) * N + i;
        tl = a.t3_l[o];
        tm = a.t3_m[o];
        ua = a.u_acc[o];
        uaf = (float)ua;
      } else {
        if (MODE == MODE_STRICT) {
          tl = t3_strict(w0, w1);
          tm = t3_strict(w2, w3);
        } else {
          tl = (double)t3_fast(w0, w1);
          tm = (double)t3_fast(w2, w3);
        }
        ur = low_bytes(w0, w1, w2, w3);
        uaf = u32f(ur);
      }
      double pl = ll + s_l * tl, pm = lm + s_m * tm;           // bi:318-324
      clip70_pair(pl, pm);
      // a proposal with log mu > 5 has target -inf and is never accepted (exp(-inf - cur) = 0, or NaN when cur is
      // -inf too); cur itself can be -inf only at the start, and then every admissible proposal is accepted (d = +inf)
      const bool admissible = !(pm > 5.0);
      const double prop = log_post_open(pl, pm, xd, omz, Tz, m0, m1, h00, h01, h11, tab_addr);
      if (mh_accept<MODE == MODE_INJECT>(prop - cur, uaf, [&]() { return MODE == MODE_INJECT ? ua : u32d(ur); }) &&
          admissible) {
        ll = pl;
        lm = pm;
        cur = prop;
      }
    };
    for (int s = 0; s < S; ++s) {
      uint4 A = make_uint4(0u, 0u, 0u, 0u);
      if (MODE != MODE_INJECT) A = philox4x32_10_rk(gid, sw.sweep, 1u + (uint32_t)s, c3, a.rk);   // word layout: clv_rng.cuh
      mh_step(s, A.x, A.y, A.z, A.w);
    }
    }
    a.ll[cN + i] = ll;
    a.lm[cN + i] = lm;
    // ---- eta (tri:306-333, 524-526) ------------------------------------------------------------
    double le = 0.0;
    if (D == 3) {
      double n;
      if (MODE == MODE_INJECT) {
        n = a.n_eta[cN + i];
      } else {
        double ns;
        normal_pair_u53(philox4x32_10_rk(gid, sw.sweep, 1u + 2u * (uint32_t)S, c3, a.rk), &n, &ns);
      }
      const double prior_var = cp.Sigma[8];
      double post_mean = cp.eta_post_var * (a.log_s[i] / mc.omega2 + m2 / prior_var);
      le = post_mean + cp.eta_sd * n;
      a.le[cN + i] = le;
    }
    if (sw.store_zt) {
      a.z[cN + i] = zf;
      a.tau[cN + i] = tau;
    }
    // ---- kept draw: lambda, mu, tau, z(, eta)   bi:407-410, tri:544-548 -------------------------
    if (keep) {
      const double lam_n = exp_tab(ll, s_tab), mu_n = exp_tab(lm, s_tab);
      constexpr int NC = (D == 2) ? 4 : 5;
      if (sw.draws) {           // level-1 storage is optional (clv_run with level1 == NULL)
        double* o = sw.draws + (((long long)chain * sw.chunk_cap + sw.slot) * N + i) * NC;
        if (D == 2) {
          reinterpret_cast<double2*>(o)[0] = make_double2(lam_n, mu_n);
          reinterpret_cast<double2*>(o)[1] = make_double2(tau, zf);
        } else {
          o[0] = lam_n; o[1] = mu_n; o[2] = tau; o[3] = zf; o[4] = exp(le);
        }
      }
      if (FUSE_FC && a.fc) fused_forecast_cell(a.fc, gid, (long long)chain * a.loglik_stride + sw.draw_index, lam_n, tau, zf, T, cN + i);
      // (E32: x, 1-z, Tz come back from the stash, so they do not occupy registers across the Metropolis loop)
      const double xk = E32 ? s_stash[tid] : xd, ok = E32 ? s_stash[SWEEP_THREADS + tid] : omz,
                   Tk = E32 ? s_stash[2 * SWEEP_THREADS + tid] : Tz;
      lik = xk * ll + ok * lm - (lam_n + mu_n) * Tk;         // bi:423-427
      lik = fmin(fmax(lik, -1048576.0), 1048576.0);
    }
    yc0 = ll - mc.center[0];
    yc1 = lm - mc.center[1];
    if (D == 3) yc2 = le - mc.center[2];
  }
  accumulate_stats<D>(mc, a.Xc, N, i, valid, yc0, yc1, yc2, s_priv);
  if (keep && valid) s_priv[(K * D + D * (D + 1) / 2) * SWEEP_THREADS + tid] += to_fx(lik, mc.ll_scale);
}

// ---- two customers per thread ---------------------------------------------------------------------------------------------
// The Metropolis loop is a chain of dependent instructions per customer (Philox rounds, SFU, fp64 exp, accept) and the
// sweep kernel issues on only ~68 % of its scheduler slots with the 8 warps per scheduler that 64 registers allow
// (stalls: fixed-latency `wait`, pipe throttles).  sweep_tile2 gives every thread TWO independent customers (i and
// i + 128 of a 256-customer tile) and walks them through the steps together, so each warp carries two independent
// instruction streams for the scheduler to interleave.  Per customer the arithmetic is exactly sweep_tile's (same
// functions, same order), so chains are bit-identical whichever variant runs.  Philox modes only (FAST / STRICT).
#ifndef CLV_CPT
#define CLV_CPT 2
#endif
constexpr int CPT = CLV_CPT;
template <int D, int MODE>
struct Cust2 {
  double ll, lm, cur, xd, omz, Tz, m0, m1, m2, tau, zf;
  E32State st;               // E32: the Metropolis loop's state (ll, lm live there; x, 1-z, Tz, m0, m1 in the stash)
  uint32_t gid;
  long long i;
  bool valid;
};
// a customer's inputs, fetched for all customers of the thread before any of them is used (the loads overlap)
struct Cust2In {
  double xd, tx, T, ll, lm, xc[XC_STASH];
};
template <int D>
__device__ __forceinline__ void cust_load(Cust2In& in, const SweepArgs& a, const ModelConst& mc, long long cN, long long i) {
  const long long N = mc.N;
  const long long ii = (i < N) ? i : N - 1;   // a lane beyond the end shadows the last customer; nothing of it is stored
  in.xd = (double)a.x[ii];
  in.tx = a.t_x[ii];
  in.T = a.T_cal[ii];
  in.ll = a.ll[cN + ii];
  in.lm = a.lm[cN + ii];
#pragma unroll
  for (int k = 0; k < XC_STASH; ++k) in.xc[k] = (k + 1 < mc.K) ? a.Xc[(long long)k * N + ii] : 0.0;
}

template <int D, int MODE>
__device__ __forceinline__ void cust_begin(Cust2<D, MODE>& c, const Cust2In& in, const SweepArgs& a, const ModelConst& mc,
                                           const ChainParams& cp, const double* s_beta, const double* s_tab, const SweepStep& sw,
                                           int chain, long long cN, long long i, uint32_t c3, double h00, double h01, double h11,
                                           double* stash, double* xs) {
  const int K = mc.K;
  const long long N = mc.N;
  c.valid = i < N;
  c.i = c.valid ? i : N - 1;
  c.gid = (uint32_t)(mc.gid_offset + c.i);
  const double xd = in.xd, tx = in.tx, T = in.T, ll = in.ll, lm = in.lm;
  double m0 = s_beta[0], m1 = s_beta[1];
  c.m2 = (D == 3) ? s_beta[2] : 0.0;
#pragma unroll
  for (int k = 1; k <= XC_STASH; ++k)
    if (k < K) {
      const double xk = in.xc[k - 1];
      xs[(k - 1) * (CPT * SWEEP_THREADS)] = xk;
      m0 = fma(xk, s_beta[k * D + 0], m0);
      m1 = fma(xk, s_beta[k * D + 1], m1);
      if (D == 3) c.m2 = fma(xk, s_beta[k * D + 2], c.m2);
    }
  for (int k = XC_STASH + 1; k < K; ++k) {
    const double xk = a.Xc[(long long)(k - 1) * N + c.i];
    m0 = fma(xk, s_beta[k * D + 0], m0);
    m1 = fma(xk, s_beta[k * D + 1], m1);
    if (D == 3) c.m2 = fma(xk, s_beta[k * D + 2], c.m2);
  }
  const double lam = exp_tab(ll, s_tab), mu = exp_tab(lm, s_tab);
  const uint4 r = philox4x32_10_rk(c.gid, sw.sweep, 0u, c3, a.rk);
  const double uz = u53(r.x, r.y), ut = u53(r.z, r.w);
  const double ml = mu + lam;
  const double e = exp_any(-(ml * (T - tx)), s_tab);
  const double pa = (ml * e) / (ml * e + mu * (1.0 - e));
  const bool alive = uz < pa;
  const double mtx = fmin(700.0, ml * tx), mT = fmin(700.0, ml * T);
  const double mix = (1.0 - ut) * exp_tab(-mtx, s_tab) + ut * exp_tab(-mT, s_tab);
  const double lg = -log(alive ? ut : mix);
  c.tau = (alive ? T : 0.0) + lg / (alive ? mu : ml);
  c.zf = alive ? 1.0 : 0.0;
  const double omz = 1.0 - c.zf;
  const double Tz = alive ? T : c.tau;
  if constexpr ((CLV_E32 != 0) && MODE == MODE_FAST) {
    // z and tau are final here: they leave now (state arrays, columns 2-3 of a kept draw's row), so that nothing but the
    // Metropolis state is alive in the loop
    if (c.valid) {
      if (sw.store_zt) {
        a.z[cN + c.i] = c.zf;
        a.tau[cN + c.i] = c.tau;
      }
      if (sw.keep && sw.draws) {
        constexpr int NC = (D == 2) ? 4 : 5;
        double* o = sw.draws + (((long long)chain * sw.chunk_cap + sw.slot) * N + c.i) * NC;
        if (D == 2) reinterpret_cast<double2*>(o)[1] = make_double2(c.tau, c.zf);
        else { o[2] = c.tau; o[3] = c.zf; }
      }
    }
    e32_begin(c.st, ll, lm, lam, mu, xd, omz, Tz, m0, m1, cp, mc.S, stash);
  } else {
    c.ll = ll; c.lm = lm; c.xd = xd; c.omz = omz; c.Tz = Tz; c.m0 = m0; c.m1 = m1;
    c.cur = log_post(ll, lm, xd, omz, Tz, m0, m1, h00, h01, h11, s_tab);
  }
}

template <int D, int MODE>
__device__ __forceinline__ void cust_end(Cust2<D, MODE>& c, const SweepArgs& a, const ModelConst& mc, const ChainParams& cp,
                                         const double* s_tab, long long* s_priv, const SweepStep& sw, int chain, long long cN, uint32_t c3,
                                         long long i, const double* stash, const double* xs) {
  const int K = mc.K, S = mc.S;
  const long long N = mc.N;
  double le = 0.0, lik = 0.0;
  constexpr bool E32 = (CLV_E32 != 0) && MODE == MODE_FAST;
  if constexpr (E32) {
    c.ll = c.st.ll; c.lm = c.st.lm;
    c.xd = stash[0]; c.omz = stash[SWEEP_THREADS]; c.Tz = stash[2 * SWEEP_THREADS];
    c.valid = i < N;                             // recomputed, not carried through the loop
    c.i = c.valid ? i : N - 1;
    c.gid = (uint32_t)(mc.gid_offset + c.i);
  }
  if (c.valid) {
    a.ll[cN + c.i] = c.ll;
    a.lm[cN + c.i] = c.lm;
    if (D == 3) {
      double n, ns;
      normal_pair_u53(philox4x32_10_rk(c.gid, sw.sweep, 1u + 2u * (uint32_t)S, c3, a.rk), &n, &ns);
      const double post_mean = cp.eta_post_var * (a.log_s[c.i] / mc.omega2 + c.m2 / cp.Sigma[8]);
      le = post_mean + cp.eta_sd * n;
      a.le[cN + c.i] = le;
    }
    if (!E32 && sw.store_zt) {
      a.z[cN + c.i] = c.zf;
      a.tau[cN + c.i] = c.tau;
    }
    if (sw.keep) {
      const double lam_n = exp_tab(c.ll, s_tab), mu_n = exp_tab(c.lm, s_tab);
      constexpr int NC = (D == 2) ? 4 : 5;
      if (sw.draws) {
        double* o = sw.draws + (((long long)chain * sw.chunk_cap + sw.slot) * N + c.i) * NC;
        if (D == 2) {
          reinterpret_cast<double2*>(o)[0] = make_double2(lam_n, mu_n);
          if (!E32) reinterpret_cast<double2*>(o)[1] = make_double2(c.tau, c.zf);
        } else {
          o[0] = lam_n; o[1] = mu_n; o[4] = exp(le);
          if (!E32) { o[2] = c.tau; o[3] = c.zf; }
        }
      }
      lik = c.xd * c.ll + c.omz * c.lm - (lam_n + mu_n) * c.Tz;
      lik = fmin(fmax(lik, -1048576.0), 1048576.0);
    }
  }
  accumulate_stats<D>(mc, a.Xc, N, c.i, c.valid, c.ll - mc.center[0], c.lm - mc.center[1], (D == 3) ? le - mc.center[2] : 0.0, s_priv, xs,
                      CPT * SWEEP_THREADS);
  if (sw.keep && c.valid) s_priv[(K * D + D * (D + 1) / 2) * SWEEP_THREADS + threadIdx.x] += to_fx(lik, mc.ll_scale);
}

template <int D, int MODE>
__device__ __forceinline__ void sweep_tile2(const SweepArgs& a, const ModelConst& mc, const ChainParams& cp, const double* s_beta,
                                            const double* s_tab, long long* s_priv, const SweepStep& sw, int chain, long long tile,
                                            uint32_t c3, double* s_stash, double* s_xc) {
  static_assert(MODE != MODE_INJECT, "two customers per thread: Philox modes only");
  constexpr bool E32 = (CLV_E32 != 0) && MODE == MODE_FAST;
  const int S = mc.S;
  const long long cN = (long long)chain * mc.N;
  const double h00 = -0.5 * cp.P00, h01 = -cp.P01, h11 = -0.5 * cp.P11;
  const double t3s = (MODE == MODE_FAST) ? 1.7320508075688772 : 1.0;
  const double s_l = cp.Sigma[0] * t3s, s_m = cp.Sigma[D + 1] * t3s;
  Cust2<D, MODE> c[CPT];
  // customer j of this thread keeps its exact-path inputs at s_stash[(f CPT + j) 128 + tid], f = 0..4, and its first
  // covariates at s_xc[(k CPT + j) 128 + tid]
  {
    Cust2In in[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
      cust_load<D>(in[j], a, mc, cN, tile * (CPT * SWEEP_THREADS) + j * SWEEP_THREADS + threadIdx.x);
#pragma unroll
    for (int j = 0; j < CPT; ++j)
      cust_begin<D, MODE>(c[j], in[j], a, mc, cp, s_beta, s_tab, sw, chain, cN, tile * (CPT * SWEEP_THREADS) + j * SWEEP_THREADS + threadIdx.x,
                          c3, h00, h01, h11, s_stash + (size_t)j * E32_STASH * SWEEP_THREADS + threadIdx.x,
                          s_xc + j * SWEEP_THREADS + threadIdx.x);
  }
  if constexpr (E32) {
    const double qa = cp.qa, qb = cp.qb, qc = cp.qc;
    for (int s = 0; s < S; ++s) {
      E32Prop pr[CPT];
      uint32_t ur[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint4 A = philox4x32_10_rk(c[j].gid, sw.sweep, 1u + (uint32_t)s, c3, a.rk);
        ur[j] = low_bytes(A.x, A.y, A.z, A.w);
        e32_propose(c[j].st, pr[j], t3_fast(A.x, A.y), t3_fast(A.z, A.w), s_l, s_m, qa, qb, qc);
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        e32_decide(c[j].st, pr[j], ur[j], qa, qb, qc, s_stash + (size_t)j * E32_STASH * SWEEP_THREADS + threadIdx.x, &cp, s_tab);
    }
  } else {
    uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(s_tab);
    asm volatile("" : "+r"(tab_addr));
    for (int s = 0; s < S; ++s) {
      double pl[CPT], pm[CPT], prop[CPT];
      uint32_t ur[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint4 A = philox4x32_10_rk(c[j].gid, sw.sweep, 1u + (uint32_t)s, c3, a.rk);
        double tl, tm;
        if (MODE == MODE_STRICT) {
          tl = t3_strict(A.x, A.y);
          tm = t3_strict(A.z, A.w);
        } else {
          tl = (double)t3_fast(A.x, A.y);
          tm = (double)t3_fast(A.z, A.w);
        }
        ur[j] = low_bytes(A.x, A.y, A.z, A.w);
        pl[j] = c[j].ll + s_l * tl;
        pm[j] = c[j].lm + s_m * tm;
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) clip70_pair(pl[j], pm[j]);
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        prop[j] = log_post_open(pl[j], pm[j], c[j].xd, c[j].omz, c[j].Tz, c[j].m0, c[j].m1, h00, h01, h11, tab_addr);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t u = ur[j];
        const bool admissible = !(pm[j] > 5.0);
        if (mh_accept<false>(prop[j] - c[j].cur, u32f(u), [&]() { return u32d(u); }) && admissible) {
          c[j].ll = pl[j];
          c[j].lm = pm[j];
          c[j].cur = prop[j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j)
    cust_end<D, MODE>(c[j], a, mc, cp, s_tab, s_priv, sw, chain, cN, c3, tile * (CPT * SWEEP_THREADS) + j * SWEEP_THREADS + threadIdx.x,
                      s_stash + (size_t)j * E32_STASH * SWEEP_THREADS + threadIdx.x, s_xc + j * SWEEP_THREADS + threadIdx.x);
}

// the sweep kernel with CPT (= 2) customers per thread (tiles of 128 CPT customers)
#ifndef CLV_MINBLOCKS2
#define CLV_MINBLOCKS2 4
#endif
template <int D, int MODE>
__global__ void __launch_bounds__(SWEEP_THREADS, CLV_MINBLOCKS2) k_sweep2(SweepArgs a) {
  extern __shared__ long long s_priv[];          // [(nstat + 1)][SWEEP_THREADS]
  __shared__ double s_beta[MAXK * MAXD];
  __shared__ double s_tab[EXP_N];
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  __shared__ double s_stash[E32_STASH * CPT * SWEEP_THREADS];     // the exact path's inputs (E32)
  __shared__ double s_xc[XC_STASH * CPT * SWEEP_THREADS];         // the first covariates of the tile's customers
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y;
  const int tid = threadIdx.x;
  const int K = mc.K;
  const int nstat = K * D + D * (D + 1) / 2;
  const ChainParams& cp = a.params[chain];
  if (a.pdl_early) pdl_launch_dependents();
  for (int t = tid; t < EXP_N; t += SWEEP_THREADS) s_tab[t] = c_exptab[t];
  for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
  clear_stats(s_priv, nstat);
  const bool peer_stopped = a.error_flag && *(volatile const int*)a.error_flag == 2;
  pdl_wait();
  if (peer_stopped) return;
  for (int t = tid; t < K * D; t += SWEEP_THREADS) s_beta[t] = cp.beta[t];
  __syncthreads();
  const uint32_t c3 = dom_word(DOM_SAMPLER, a.chain_offset + (uint32_t)chain);
  SweepStep sw;
  sw.sweep = a.sweep; sw.keep = a.slot >= 0; sw.store_zt = a.store_zt; sw.slot = a.slot; sw.chunk_cap = a.chunk_cap;
  sw.draws = a.draws; sw.draw_index = a.draw_index;
  if (a.tile_counter) {
    // Dynamic tiles.  One synchronisation per sweep means the sweep ends when its LAST tile ends: with a static split the
    // blocks that drew one tile more than the others run alone for a whole tile latency (~30 us with two customers per
    // thread).  Here the blocks of the (one-wave) grid draw tile numbers from a counter, and the tail of the tile list is
    // cut finer: the last customers go 128 at a time, one per thread -- half the latency per tile.
    // The number of a block's NEXT tile is drawn when the current tile starts and published when it ends, so the atomic's
    // round trip hides behind the tile and one barrier per tile suffices (two slots, alternating).
    __shared__ unsigned int s_next[2];
    unsigned int* ctr = a.tile_counter + (size_t)(a.sweep & 1u) * gridDim.y + chain;
    if (blockIdx.x == 0 && tid == 0) a.tile_counter[(size_t)((a.sweep + 1u) & 1u) * gridDim.y + chain] = 0u;   // for the next sweep
    const long long nt = a.n_big + a.n_small;
    if (tid == 0) s_next[0] = atomicAdd(ctr, 1u);
    __syncthreads();
    for (int p = 0;; p ^= 1) {
      const long long t = s_next[p];
      if (t >= nt) break;
      unsigned int nxt = 0u;
      if (tid == 0) nxt = atomicAdd(ctr, 1u);
      if (CPT >= 2 && t >= a.n_big) sweep_tile<D, MODE>(a, mc, cp, s_beta, s_tab, s_priv, sw, chain, CPT * a.n_big + (t - a.n_big), c3, s_stash);
      else sweep_tile2<D, MODE>(a, mc, cp, s_beta, s_tab, s_priv, sw, chain, t, c3, s_stash, s_xc);
      if (tid == 0) s_next[p ^ 1] = nxt;
      __syncthreads();
    }
  } else {
    const long long ntiles = (mc.N + CPT * SWEEP_THREADS - 1) / (CPT * SWEEP_THREADS);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      sweep_tile2<D, MODE>(a, mc, cp, s_beta, s_tab, s_priv, sw, chain, tile, c3, s_stash, s_xc);
  }
  if (!a.pdl_early) pdl_launch_dependents();
  flush_stats(s_priv, s_acc, nstat, sw.keep != 0);
  __syncthreads();
  for (int t = tid; t < nstat; t += SWEEP_THREADS)
    if (s_acc[t]) atomicAdd(&a.acc[chain * NSTAT_MAX + t], s_acc[t]);
  if (sw.keep && tid == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&a.loglik_acc[chain * a.loglik_stride + a.draw_index]),
              s_acc[NSTAT_MAX]);
}

// FUSE_FC: the instantiation that also simulates x* of a kept draw (clv_set_fused_forecast).  A separate instantiation, so
// that the out-of-line call in the keep branch costs the ordinary sweep kernel no register (it put a spill reload into
// the Metropolis loop when it was a run-time option).
template <int D, int MODE, bool FUSE_FC = false>
__global__ void __launch_bounds__(SWEEP_THREADS, CLV_MINBLOCKS) k_sweep(SweepArgs a) {
  extern __shared__ long long s_priv[];          // [(nstat + 1)][SWEEP_THREADS]
  __shared__ double s_beta[MAXK * MAXD];
  __shared__ double s_tab[EXP_N];
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  __shared__ double s_stash[E32_STASH * SWEEP_THREADS];           // the exact path's inputs (E32)
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y;
  const int tid = threadIdx.x;
  const int K = mc.K;
  const int nstat = K * D + D * (D + 1) / 2;
  const ChainParams& cp = a.params[chain];
  // prologue (programmatic dependent launch): nothing here depends on the level-2 kernel that precedes this launch
  if (a.pdl_early) pdl_launch_dependents();
  for (int t = tid; t < EXP_N; t += SWEEP_THREADS) s_tab[t] = c_exptab[t];
  for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
  clear_stats(s_priv, nstat);
  // a peer rank stopped (flag raised by an EARLIER level-2 kernel): do not sample on partial sums.  Read before the wait, off
  // the critical path; the sweep right after the failed exchange still runs, on the unchanged parameters of the sweep before
  const bool peer_stopped = a.error_flag && *(volatile const int*)a.error_flag == 2;
  pdl_wait();                                   // (beta, Sigma) of this sweep and the state of the previous one are complete
  if (peer_stopped) return;
  for (int t = tid; t < K * D; t += SWEEP_THREADS) s_beta[t] = cp.beta[t];
  __syncthreads();
  const uint32_t c3 = dom_word(DOM_SAMPLER, a.chain_offset + (uint32_t)chain);
  SweepStep sw;
  sw.sweep = a.sweep; sw.keep = a.slot >= 0; sw.store_zt = a.store_zt; sw.slot = a.slot; sw.chunk_cap = a.chunk_cap;
  sw.draws = a.draws; sw.draw_index = a.draw_index;
  const long long ntiles = (mc.N + SWEEP_THREADS - 1) / SWEEP_THREADS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    sweep_tile<D, MODE, FUSE_FC>(a, mc, cp, s_beta, s_tab, s_priv, sw, chain, tile, c3, s_stash);
  if (!a.pdl_early) pdl_launch_dependents();
  flush_stats(s_priv, s_acc, nstat, sw.keep != 0);
  __syncthreads();
  for (int t = tid; t < nstat; t += SWEEP_THREADS)
    if (s_acc[t]) atomicAdd(&a.acc[chain * NSTAT_MAX + t], s_acc[t]);
  if (sw.keep && tid == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&a.loglik_acc[chain * a.loglik_stride + a.draw_index]),
              s_acc[NSTAT_MAX]);
}

// Statistics of the current state only (first level-2 draw of the bivariate order, bi:393).
template <int D>
__global__ void __launch_bounds__(SWEEP_THREADS) k_stats_only(SweepArgs a) {
  extern __shared__ long long s_priv[];
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y, tid = threadIdx.x;
  const long long N = mc.N, cN = (long long)chain * N;
  const int nstat = mc.K * D + D * (D + 1) / 2;
  for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
  clear_stats(s_priv, nstat);
  __syncthreads();
  const long long ntiles = (N + SWEEP_THREADS - 1) / SWEEP_THREADS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i = tile * SWEEP_THREADS + tid;
    const bool valid = i < N;
    double yc0 = 0, yc1 = 0, yc2 = 0;
    if (valid) {
      yc0 = a.ll[cN + i] - mc.center[0];
      yc1 = a.lm[cN + i] - mc.center[1];
      if (D == 3) yc2 = a.le[cN + i] - mc.center[2];
    }
    accumulate_stats<D>(mc, a.Xc, N, i, valid, yc0, yc1, yc2, s_priv);
  }
  flush_stats(s_priv, s_acc, nstat, false);
  __syncthreads();
  for (int t = tid; t < nstat; t += SWEEP_THREADS)
    if (s_acc[t]) atomicAdd(&a.acc[chain * NSTAT_MAX + t], s_acc[t]);
}

// Test hook: the level-1 variates of MH step `step` (proposal t3 for log lambda / log mu, accept uniform) that customer
// gid of chain 0 consumes in sweep `sweep`, exactly as sweep_tile generates them (FAST: times the sqrt(3) the kernel
// folds into the proposal scale).
template <int MODE>
__global__ void k_debug_variates(PhiloxRoundKeys rk, uint32_t sweep, int step, long long n, double* t3l, double* t3m,
                                 double* uacc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t gid = (uint32_t)i, c3 = dom_word(DOM_SAMPLER, 0u);
    const uint4 A = philox4x32_10_rk(gid, sweep, 1u + (uint32_t)step, c3, rk);
    const uint32_t w[4] = {A.x, A.y, A.z, A.w};
    if (MODE == MODE_STRICT) {
      t3l[i] = t3_strict(w[0], w[1]);
      t3m[i] = t3_strict(w[2], w[3]);
    } else {
      t3l[i] = 1.7320508075688772 * (double)t3_fast(w[0], w[1]);
      t3m[i] = 1.7320508075688772 * (double)t3_fast(w[2], w[3]);
    }
    uacc[i] = u32d(low_bytes(w[0], w[1], w[2], w[3]));
  }
}

// ------------------------------------------------------------------------------------------------
// initialisation statistics (bi:367-374, tri:488-499) as exact integer sums: every term is rounded once to
// fixed point (bits chosen from the global max |term|) and added in int64 (lo 32 bits / hi part separately),
// so the totals do not depend on thread, block, shard or GPU count.  Mirrors mcmc_clv_model_b200/hostmath.py.
// ------------------------------------------------------------------------------------------------
constexpr int NQ_MAX = 3 + MAXK * (MAXK + 1) / 2;

struct InitQArgs {
  const int* x;
  const double *t_x, *T_cal, *Xc, *log_s;
  long long N;
  int K, D, phase, mode;          // phase 0: x, tden, log_s, X_a X_b ; phase 1: mu_init, (log_s - m)^2.  mode 0: max |v| ; 1: sums
  double lam_init, mean_log_s;
  double scale[NQ_MAX];           // 2^bits per quantity (mode 1)
  unsigned long long* out_max;    // [NQ] bit patterns of max |v|
  long long* out_sum;             // [NQ][2]: lo (low 32 bits of each term), hi (term >> 32)
};

__global__ void __launch_bounds__(256) k_init_quantities(InitQArgs a) {
  __shared__ unsigned long long s_max[NQ_MAX];
  __shared__ unsigned long long s_sum[2 * NQ_MAX];
  for (int t = threadIdx.x; t < NQ_MAX; t += blockDim.x) { s_max[t] = 0ull; s_sum[2 * t] = 0ull; s_sum[2 * t + 1] = 0ull; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nwarp_tiles = (a.N + 31) / 32;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nq = a.phase == 0 ? 3 + a.K * (a.K + 1) / 2 : 2;
  for (long long wt = wid; wt < nwarp_tiles; wt += nw) {
    const long long i = wt * 32 + lane;
    const bool valid = i < a.N;
    double xv[MAXK];
    xv[0] = 1.0;
    double tx = 0, T = 0, ls = 0, xd = 0;
    if (valid) {
      xd = (double)a.x[i]; tx = a.t_x[i]; T = a.T_cal[i];
      if (a.D == 3) ls = a.log_s[i];
      for (int k = 1; k < a.K; ++k) xv[k] = a.Xc[(long long)(k - 1) * a.N + i];
    }
    int pa = 0, pb = 0;
    for (int q = 0; q < nq; ++q) {
      double v = 0.0;
      if (a.phase == 0) {
        if (q == 0) v = xd;
        else if (q == 1) v = (tx == 0.0) ? T : tx;                      // bi:368
        else if (q == 2) v = ls;
        else { v = xv[pa] * xv[pb]; if (++pb == a.K) { ++pa; pb = pa; } }  // pairs a <= b, row-major
      } else {
        if (q == 0) v = 1.0 / (tx + 0.5 / a.lam_init);                    // bi:370
        else { double d = ls - a.mean_log_s; v = d * d; }                 // tri:494
      }
      if (!valid) v = 0.0;
      if (a.mode == 0) {
        unsigned long long m = (unsigned long long)__double_as_longlong(fabs(v));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_down_sync(0xffffffffu, m, o); m = t > m ? t : m; }
        if (lane == 0 && m) atomicMax(&s_max[q], m);
      } else {
        const long long term = __double2ll_rn(v * a.scale[q]);
        const long long lo = warp_sum_ll(term & 0xffffffffll), hi = warp_sum_ll(term >> 32);
        if (lane == 0) {
          if (lo) atomicAdd(&s_sum[2 * q], (unsigned long long)lo);
          if (hi) atomicAdd(&s_sum[2 * q + 1], (unsigned long long)hi);
        }
      }
    }
  }
  // one global atomic per quantity per block (the per-warp ones go to shared memory)
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    if (a.mode == 0) {
      if (s_max[q]) atomicMax(&a.out_max[q], s_max[q]);
    } else {
      if (s_sum[2 * q]) atomicAdd((unsigned long long*)&a.out_sum[2 * q], s_sum[2 * q]);
      if (s_sum[2 * q + 1]) atomicAdd((unsigned long long*)&a.out_sum[2 * q + 1], s_sum[2 * q + 1]);
    }
  }
}

// The same quantities for a compile-time K <= 5 (every configuration of BASELINE.json): each thread keeps the maxima / the
// lo and hi sums of its customers in registers (statically indexed) and the warp and block reductions happen ONCE at the
// end instead of once per 32 customers (k_init_quantities: 18 quantities x two 64-bit warp sums per warp tile; 10 M
// customers: 0.65 + 0.80 + 0.14 + 0.15 ms for the four passes of one initialisation).  Integer sums (mod 2^64) and maxima
// do not depend on the order, so the totals are bit-identical to k_init_quantities' and to hostmath.py's.
template <int KT>
__global__ void __launch_bounds__(256) k_init_quantities_t(InitQArgs a) {
  constexpr int NQ0 = 3 + KT * (KT + 1) / 2;
  __shared__ unsigned long long s_max[NQ0];
  __shared__ unsigned long long s_sum[2 * NQ0];
  for (int t = threadIdx.x; t < NQ0; t += blockDim.x) { s_max[t] = 0ull; s_sum[2 * t] = 0ull; s_sum[2 * t + 1] = 0ull; }
  __syncthreads();
  unsigned long long acc[2 * NQ0];          // mode 0: acc[q] = bit pattern of max |v| ; mode 1: acc[2q], acc[2q+1] = lo, hi sums
#pragma unroll
  for (int t = 0; t < 2 * NQ0; ++t) acc[t] = 0ull;
  const bool sums = a.mode == 1;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto take = [&](int q, double v) {
    if (!sums) {
      const unsigned long long m = (unsigned long long)__double_as_longlong(fabs(v));
      acc[q] = m > acc[q] ? m : acc[q];
    } else {
      const long long term = __double2ll_rn(v * a.scale[q]);
      acc[2 * q] += (unsigned long long)(term & 0xffffffffll);
      acc[2 * q + 1] += (unsigned long long)(term >> 32);
    }
  };
  if (a.phase == 0) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += stride) {
      double xv[KT];
      xv[0] = 1.0;
#pragma unroll
      for (int k = 1; k < KT; ++k) xv[k] = a.Xc[(long long)(k - 1) * a.N + i];
      const double tx = a.t_x[i], T = a.T_cal[i];
      take(0, (double)a.x[i]);
      take(1, (tx == 0.0) ? T : tx);                                       // bi:368
      take(2, (a.D == 3) ? a.log_s[i] : 0.0);
      int q = 3;
#pragma unroll
      for (int pa = 0; pa < KT; ++pa)
#pragma unroll
        for (int pb = pa; pb < KT; ++pb) take(q++, xv[pa] * xv[pb]);       // pairs a <= b, row-major
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += stride) {
      take(0, 1.0 / (a.t_x[i] + 0.5 / a.lam_init));                        // bi:370
      const double d = ((a.D == 3) ? a.log_s[i] : 0.0) - a.mean_log_s;     // tri:494
      take(1, d * d);
    }
  }
  const int nq = a.phase == 0 ? NQ0 : 2;
#pragma unroll
  for (int q = 0; q < NQ0; ++q) {
    if (q >= nq) break;
    if (!sums) {
      unsigned long long m = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_down_sync(0xffffffffu, m, o); m = t > m ? t : m; }
      if (lane == 0 && m) atomicMax(&s_max[q], m);
    } else {
      const long long lo = warp_sum_ll((long long)acc[2 * q]), hi = warp_sum_ll((long long)acc[2 * q + 1]);
      if (lane == 0) {
        if (lo) atomicAdd(&s_sum[2 * q], (unsigned long long)lo);
        if (hi) atomicAdd(&s_sum[2 * q + 1], (unsigned long long)hi);
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    if (!sums) {
      if (s_max[q]) atomicMax(&a.out_max[q], s_max[q]);
    } else {
      if (s_sum[2 * q]) atomicAdd((unsigned long long*)&a.out_sum[2 * q], s_sum[2 * q]);
      if (s_sum[2 * q + 1]) atomicAdd((unsigned long long*)&a.out_sum[2 * q + 1], s_sum[2 * q + 1]);
    }
  }
}

// initial level-1 state: lambda_i = lam_init, mu_i = 1/(t_x + 0.5/lam_init), eta_i = 1   (bi:369-370, tri:493)
__global__ void __launch_bounds__(256) k_init_state(const double* t_x, long long N, int chains, int D, double lam_init,
                                                    double* ll, double* lm, double* le) {
  const double ll0 = log(lam_init);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const double lm0 = log(1.0 / (t_x[i] + 0.5 / lam_init));
    for (int c = 0; c < chains; ++c) {
      ll[(long long)c * N + i] = ll0;
      lm[(long long)c * N + i] = lm0;
      if (D == 3) le[(long long)c * N + i] = 0.0;
    }
  }
}

// row-major (N, K) design matrix -> SoA covariate columns [(K-1)][N].  Column 0, the intercept, is implicit; it is
// checked here (*bad_intercept = 1 unless it is all ones) instead of by a strided host pass over the matrix.
__global__ void __launch_bounds__(256) k_split_columns(const double* X, long long N, int K, double* Xc, int* bad_intercept) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N * K; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / K;
    const int k = (int)(t - i * K);
    const double v = X[t];
    if (k > 0) Xc[(long long)(k - 1) * N + i] = v;
    else if (v != 1.0) *bad_intercept = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// level 2
// ------------------------------------------------------------------------------------------------
// P = inv(Sigma) (entries 00, 01, 11) and the eta conjugate scalars.
template <int D>
__device__ inline void derive_params(ChainParams& cp, double omega2) {
  const double* S = cp.Sigma;
  if (D == 2) {
    const double rdet = 1.0 / (S[0] * S[3] - S[1] * S[2]);
    cp.P00 = S[3] * rdet;
    cp.P01 = -S[1] * rdet;
    cp.P11 = S[0] * rdet;
    cp.eta_post_var = 0.0;
    cp.eta_sd = 0.0;
  } else {
    double c00 = S[4] * S[8] - S[5] * S[7];
    double c01 = S[5] * S[6] - S[3] * S[8];
    double c02 = S[3] * S[7] - S[4] * S[6];
    const double rdet = 1.0 / (S[0] * c00 + S[1] * c01 + S[2] * c02);
    cp.P00 = c00 * rdet;
    cp.P01 = (S[2] * S[7] - S[1] * S[8]) * rdet;
    cp.P11 = (S[0] * S[8] - S[2] * S[6]) * rdet;
    double post_precision = 1.0 / omega2 + 1.0 / S[8];      // tri:325
    cp.eta_post_var = 1.0 / post_precision;
    cp.eta_sd = sqrt(cp.eta_post_var);
  }
  // E32 constants (three independent long operations; a non-positive-definite block gives NaN, which sends every
  // decision of the screen to the exact path)
  const double ra = rsqrt(0.5 * cp.P00);
  cp.qa = 0.5 * cp.P00 * ra;
  cp.qb = 0.5 * cp.P01 * ra;
  const double c2 = fma(-cp.qb, cp.qb, 0.5 * cp.P11);
  cp.qc = c2 * rsqrt(c2);
  const double rd = 1.0 / fma(cp.P00, cp.P11, -cp.P01 * cp.P01);
  cp.s00 = cp.P11 * rd;
  cp.s01 = -cp.P01 * rd;
  cp.s11 = cp.P00 * rd;
}

template <int D>
__global__ void k_derive_params(const ModelConst* mc, ChainParams* params, int chains) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < chains) {
    params[c].status = 0;
    derive_params<D>(params[c], mc->omega2);
  }
}

struct Level2Args {
  const ModelConst* mc;
  ChainParams* params;
  unsigned long long* acc;      // [chains][NSTAT_MAX]; zeroed after reading
  double* level2_draws;         // device [chains][n_draws][P]
  long long n_draws;
  long long draw_index;         // -1: not kept
  uint32_t sweep;
  uint32_t chain_offset;
  uint64_t seed;
  int injected;
  const double *iw_norm, *iw_chi2, *beta_norm;   // injected variates [chains][...]
  int* error_flag;
  // customer-sharded runs with peer mailboxes (NVLink/NVSwitch P2P stores): the all-reduce of the statistics is done
  // HERE, inside the level-2 kernel, instead of a separate NCCL call.  world == 0: not used.
  int world, rank, n_chains;
  int pdl_early;                                  // 1: the next sweep kernel may become resident at once, 0: after pdl_wait
  uint32_t tag;                                   // never 0; differs from the tag of sweep - 2 and of any earlier init
  long long timeout_ns;                           // give up (error flag 2) when a peer's words do not arrive in time
  unsigned long long* peer_mail[P2P_MAX_WORLD];   // rank r's mailbox [2][P2P_MAX_WORLD][chains][2 * NSTAT_MAX] words
};

// The sweep-invariant constants of the level-2 draw, staged in shared memory (they used to be re-read from global
// memory in each of ~7 dependent phases).
struct Level2Const {
  double V[MAXK * MAXK], LV[MAXK * MAXK], A0B0c[MAXK * MAXD], Q0[MAXD * MAXD], center[MAXD];
  double fx_inv, nu_n, omega2;
  int K, compat;
};

template <int D>
__device__ __forceinline__ void load_level2_const(const ModelConst& mc, Level2Const& lc, int tid, int nthreads) {
  const int K = mc.K;
  for (int t = tid; t < K * K; t += nthreads) { lc.V[t] = mc.V[t]; lc.LV[t] = mc.LV[t]; }
  for (int t = tid; t < K * D; t += nthreads) lc.A0B0c[t] = mc.A0B0c[t];
  for (int t = tid; t < D * D; t += nthreads) lc.Q0[t] = mc.Q0[t];
  for (int t = tid; t < D; t += nthreads) lc.center[t] = mc.center[t];
  if (tid == 0) { lc.fx_inv = mc.fx_inv; lc.nu_n = mc.nu_n; lc.omega2 = mc.omega2; lc.K = K; lc.compat = mc.compat; }
}

// Everything of a level-2 draw that depends only on (seed, chain, sweep) -- never on the state: the Bartlett normals
// and chi-squares (as the inverse of the Bartlett factor A) and the beta normals (as W = chol(V) z per response).
// Produced OFF the critical path: in stream mode by k_level2 before it waits for the sweep kernel (programmatic
// dependent launch), in the persistent kernel by an otherwise idle warp one sweep ahead.
struct Level2Variates {
  double trn[3], chi[3], zb[MAXD * MAXK];
  double W[MAXD * MAXK];        // [d*K + k]
  double Ainv[MAXD * MAXD];     // inverse of the lower-triangular Bartlett factor (scipy: normals below, sqrt(chi2) on the diagonal)
};

template <int D>
__device__ __forceinline__ void level2_variates(const Level2Const& lc, Level2Variates& lv, PhiloxKey key, uint32_t c3,
                                                uint32_t sweep, int injected, const double* iw_norm, const double* iw_chi2,
                                                const double* beta_norm, int lane) {
  const int K = lc.K;
  constexpr int ntril = D * (D - 1) / 2;
  const int nb = D * K;
  // fp64 Philox transforms are long dependent chains: one per lane
  for (int t = lane; t < ntril + D + nb; t += 32) {
    if (t < ntril) lv.trn[t] = injected ? iw_norm[t] : level2_normal(key, c3, sweep, (uint32_t)t);
    else if (t < ntril + D) {
      const int i = t - ntril;
      lv.chi[i] = injected ? iw_chi2[i] : level2_chi2(key, c3, sweep, 16u + i, lc.nu_n - D + 1 + i);   // chi2(nu_n - D + 1 + i)
    } else {
      const int j = t - ntril - D;
      lv.zb[j] = injected ? beta_norm[j] : level2_normal(key, c3, sweep, 32u + (uint32_t)j);
    }
  }
  __syncwarp();
  // W = chol(V) z (per response)
  for (int t = lane; t < nb; t += 32) {
    const int d = t / K, k = t - d * K;
    double s = 0.0;
    for (int m = 0; m <= k; ++m) s += lc.LV[k * K + m] * lv.zb[d * K + m];
    lv.W[t] = s;
  }
  // A^-1: A[i][i] = sqrt(chi_i), A[i][j] = normal (row-major fill below the diagonal)
  if (lane == 0) {
    double inv[D];
#pragma unroll
    for (int i = 0; i < D; ++i) inv[i] = rsqrt(lv.chi[i]);
#pragma unroll
    for (int t = 0; t < D * D; ++t) lv.Ainv[t] = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) lv.Ainv[i * D + i] = inv[i];
    if (D == 2) {
      lv.Ainv[2] = -lv.trn[0] * inv[0] * inv[1];
    } else {
      const double a10 = lv.trn[0], a20 = lv.trn[1], a21 = lv.trn[2];
      const double i10 = -a10 * inv[0] * inv[1];
      lv.Ainv[3] = i10;
      lv.Ainv[7] = -a21 * inv[1] * inv[2];
      lv.Ainv[6] = -(a20 * inv[0] + a21 * i10) * inv[2];
    }
  }
  __syncwarp();
}

// Shared-memory scratch of one level-2 draw (one warp).
struct Level2Scratch {
  double st[NSTAT_MAX];            // reduced statistics: X'Yc [k*D+d], then triu(Yc'Yc)
  double R[MAXK * MAXD], Bc[MAXK * MAXD];
  double Sn[MAXD * MAXD], CA[MAXD * MAXD];
  int ok;
};

// Conjugate multivariate regression draw (bi:233-262, tri:340-380) from the reduced statistics, by ONE WARP:
// every small matrix product is spread over the lanes; the D x D factorisations run on lane 0 (three long fp64
// operations for D = 2: rsqrt, sqrt, one reciprocal).  Works in centred responses y - c (c = prior intercept row),
// which leaves E and beta - B0 unchanged.  sc.st must be filled (and visible to the warp) on entry, lv holds the
// variates of this draw; cp (shared or global memory) receives beta, Sigma, P.
template <int D>
__device__ __forceinline__ void level2_algebra(const Level2Const& lc, Level2Scratch& sc, const Level2Variates& lv, ChainParams& cp,
                                               int lane) {
  const int K = lc.K;
  const int nb = D * K;
  // R = X'Yc + A0 B0c                                           bi:250
  for (int t = lane; t < nb; t += 32) sc.R[t] = sc.st[t] + lc.A0B0c[t];
  __syncwarp();
  // Bc = V R
  for (int t = lane; t < nb; t += 32) {
    const int k = t / D, d = t - k * D;
    double s = 0.0;
    for (int m = 0; m < K; ++m) s += lc.V[k * K + m] * sc.R[m * D + d];
    sc.Bc[t] = s;
  }
  __syncwarp();
  // S_n = S0 + E'E + C'A0C = Q0 + Yc'Yc - Bc' R                bi:253-255
  if (lane < D * D) {
    const int d = lane / D, e = lane - d * D;
    const int lo = d < e ? d : e, hi = d < e ? e : d;
    const int t = nb + lo * D - lo * (lo - 1) / 2 + (hi - lo);   // index of (lo, hi) in the row-major upper triangle
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < K; ++k) {
      s1 += sc.Bc[k * D + d] * sc.R[k * D + e];
      s2 += sc.Bc[k * D + e] * sc.R[k * D + d];
    }
    sc.Sn[lane] = sc.st[t] + lc.Q0[lane] - 0.5 * (s1 + s2);        // symmetrised
  }
  __syncwarp();
  if (lane == 0) {
    // C = chol(S_n), lower; reciprocal square roots instead of sqrt + divisions
    double C[D * D];
    bool ok = true;
#pragma unroll
    for (int t = 0; t < D * D; ++t) C[t] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double piv = sc.Sn[j * D + j];
#pragma unroll
      for (int m = 0; m < j; ++m) piv -= C[j * D + m] * C[j * D + m];
      if (!(piv > 0.0) || !isfinite(piv)) { ok = false; piv = 1.0; }
      const double r = rsqrt(piv);
      C[j * D + j] = piv * r;
#pragma unroll
      for (int i = j + 1; i < D; ++i) {
        double s = sc.Sn[i * D + j];
#pragma unroll
        for (int m = 0; m < j; ++m) s -= C[i * D + m] * C[j * D + m];
        C[i * D + j] = s * r;
      }
    }
    // Sigma ~ IW(nu_n, S_n): scipy's Bartlett construction (bi:258): CA = C A^-1 (lower)  =>  Sigma = CA CA'
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m)
          if (m >= j && m <= r) s += C[r * D + m] * lv.Ainv[m * D + j];
        sc.CA[r * D + j] = s;
      }
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int e = 0; e < D; ++e) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < D; ++m) s += sc.CA[d * D + m] * sc.CA[e * D + m];
        cp.Sigma[d * D + e] = s;
        if (!isfinite(s)) ok = false;
      }
    derive_params<D>(cp, lc.omega2);
    cp.status = ok ? 0 : 1;
    sc.ok = ok ? 1 : 0;
  }
  __syncwarp();
  // beta | Sigma: noise = kron(chol Sigma, chol V) z, ordered d*K+k           bi:261
  for (int j = lane; j < nb; j += 32) {
    const int k = j / D, d = j - k * D;
    const int src = (lc.compat == 0) ? j : d * K + k;                  // Q1: reference adds kron-ordered noise to ravel()
    const int dn = src / K, kn = src - dn * K;
    double nz = 0.0;
    for (int m = 0; m <= dn; ++m) nz += sc.CA[dn * D + m] * lv.W[m * K + kn];
    cp.beta[j] = sc.Bc[j] + (k == 0 ? lc.center[d] : 0.0) + nz;      // B_hat = Bc + e0 c'
  }
  __syncwarp();
}

// level_2 row: beta.T.ravel() then the upper triangle of Sigma (bi:411-412, tri:549-554)
template <int D>
__device__ __forceinline__ void write_level2_row(int K, const ChainParams& cp, double* o, int lane) {
  for (int t = lane; t < D * K; t += 32) {
    const int d = t / K, k = t - d * K;
    o[t] = cp.beta[k * D + d];
  }
  if (lane == 0) {
    int t = D * K;
    for (int d = 0; d < D; ++d)
      for (int e = d; e < D; ++e) o[t++] = cp.Sigma[d * D + e];
  }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One-shot all-reduce of a chain's int64 statistics over peer memory, fused into the level-2 kernel (one warp).
// Low-latency ("LL") protocol: every 64-bit word that crosses NVLink carries 32 bits of payload and a 32-bit tag, and
// an aligned 8-byte store is single-copy atomic, so a word IS its own arrival flag -- no fence, no separate flag
// round trip.  Every rank stores the two halves of each partial sum into every rank's mailbox (its own included) and
// polls its own mailbox until all world x 2 nstat words carry this sweep's tag; halves are added mod 2^64 (exact
// integers => the same totals on every rank whatever the arrival order).  Mailboxes are double-buffered by sweep
// parity: a rank can only reach sweep s+2 after every peer sent s+1, i.e. after they consumed parity s.
// Returns false on time-out (a rank stopped): the caller raises error flag 2.
__device__ __forceinline__ bool ll_allreduce_stats(const Level2Args& a, int chain, int nstat, unsigned long long* s_tot, int lane) {
  const int par = (int)(a.sweep & 1u), nw = 2 * nstat;
  const size_t words = 2 * (size_t)NSTAT_MAX;
  const size_t send_base = (((size_t)par * P2P_MAX_WORLD + a.rank) * a.n_chains + chain) * words;
  const unsigned long long tag = (unsigned long long)a.tag << 32;
  for (int t = lane; t < nstat; t += 32) s_tot[t] = 0ull;
  for (int w = lane; w < nw; w += 32) {
    const unsigned long long v = a.acc[chain * NSTAT_MAX + (w >> 1)];
    const unsigned long long word = ((w & 1) ? (v >> 32) : (v & 0xffffffffull)) | tag;
    for (int r = 0; r < a.world; ++r) st_volatile_u64(a.peer_mail[r] + send_base + w, word);
  }
  __syncwarp();
  for (int t = lane; t < nstat; t += 32) a.acc[chain * NSTAT_MAX + t] = 0ull;
  const int total = nw * a.world;
  constexpr int BATCH = 8;
  const unsigned long long* mine = a.peer_mail[a.rank];
  bool ok = true;
  long long t0 = 0;
  for (int q0 = lane; q0 < total && ok; q0 += 32 * BATCH) {
    unsigned long long x[BATCH];
    const unsigned long long* p[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int q = q0 + 32 * b;
      const int r = q / nw, w = q - r * nw;
      p[b] = mine + (((size_t)par * P2P_MAX_WORLD + r) * a.n_chains + chain) * words + w;
      x[b] = (q < total) ? ld_volatile_u64(p[b]) : tag;          // all loads of a batch in flight together
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int q = q0 + 32 * b;
      if (q >= total) continue;
      unsigned spins = 0;
      while ((x[b] & 0xffffffff00000000ull) != tag) {
        x[b] = ld_volatile_u64(p[b]);
        if ((++spins & 1023u) == 0u) {
          const long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > a.timeout_ns) { ok = false; break; }
        }
      }
      if (!ok) break;
      const int w = q % nw;
      atomicAdd(&s_tot[w >> 1], (w & 1) ? (x[b] << 32) : (x[b] & 0xffffffffull));
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  return ok;
}

// One chain's level-2 draw by one warp (the body of k_level2): prologue that needs nothing of the preceding sweep kernel
// (constants, variates), then the statistics (own accumulators, or the peer-mailbox all-reduce), then the algebra.
template <int D>
__device__ __forceinline__ void level2_chain(const Level2Args& a, int chain, int lane) {
  __shared__ Level2Scratch sc;
  __shared__ Level2Const lc;
  __shared__ Level2Variates lv;
  __shared__ unsigned long long s_tot[NSTAT_MAX];
  const ModelConst& mc = *a.mc;
  constexpr int ntril = D * (D - 1) / 2;
  // ---- prologue ---------------------------------------------------------------------------------------------------
  if (a.pdl_early) pdl_launch_dependents();  // the next sweep kernel may become resident and run ITS prologue
  load_level2_const<D>(mc, lc, lane, 32);
  __syncwarp();
  const int K = lc.K;
  const int nstat = K * D + D * (D + 1) / 2;
  level2_variates<D>(lc, lv, seed_key(a.seed), dom_word(DOM_LEVEL2, a.chain_offset + (uint32_t)chain), a.sweep, a.injected,
                     a.injected ? a.iw_norm + chain * ntril : nullptr, a.injected ? a.iw_chi2 + chain * D : nullptr,
                     a.injected ? a.beta_norm + chain * D * K : nullptr, lane);
  // ---- the statistics of the sweep kernel ------------------------------------------------------------------------------
  const bool peer_stopped = *(volatile int*)a.error_flag == 2;   // raised by an earlier exchange; read off the critical path
  pdl_wait();
  if (!a.pdl_early) pdl_launch_dependents();
  if (peer_stopped) return;                               // a peer stopped: leave the state as it is
  if (a.world <= 1) {
    for (int t = lane; t < nstat; t += 32) {
      unsigned long long* p = &a.acc[chain * NSTAT_MAX + t];
      sc.st[t] = (double)(long long)(*p) * lc.fx_inv;
      *p = 0ull;
    }
  } else {
    if (!ll_allreduce_stats(a, chain, nstat, s_tot, lane)) {
      if (lane == 0) *a.error_flag = 2;
      return;
    }
    __syncwarp();
    for (int t = lane; t < nstat; t += 32) sc.st[t] = (double)(long long)s_tot[t] * lc.fx_inv;
  }
  __syncwarp();
  ChainParams& cp = a.params[chain];
  level2_algebra<D>(lc, sc, lv, cp, lane);
  if (lane == 0 && !sc.ok) *a.error_flag = 1;
  if (a.draw_index >= 0)
    write_level2_row<D>(K, cp, a.level2_draws + ((long long)chain * a.n_draws + a.draw_index) * (D * K + D * (D + 1) / 2), lane);
}

template <int D>
__global__ void __launch_bounds__(32) k_level2(Level2Args a) {
  level2_chain<D>(a, blockIdx.x, threadIdx.x);
}

// Test hook (clv_debug_lockstep_advance): the level-2 kernels of G ranks as ONE cooperative launch, block (chain, r)
// playing rank r with rank r's accumulators, mailbox and parameters.  Kernels of separate launches must not wait on
// each other on one GPU; blocks of a cooperative launch are co-resident, so the mailbox protocol runs unchanged.
template <int D>
__global__ void __launch_bounds__(32) k_level2_ranks(const Level2Args* ranks) {
  level2_chain<D>(ranks[blockIdx.y], blockIdx.x, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// persistent cooperative kernel: whole sweeps fused, one grid-wide barrier per sweep
// ------------------------------------------------------------------------------------------------
struct PersistArgs {
  SweepArgs sw;                 // data, state, draw chunk (sw.draws, sw.chunk_cap), loglik_acc / loglik_stride
  ChainParams* params;          // [chains] (read at entry, written back at exit)
  unsigned long long* acc3;     // [3][chains][NSTAT_MAX] rotating accumulators; acc3[first_sweep % 3 ... ] see below
  double* level2_draws;
  long long n_draws;            // draws of the run (stride of level2_draws / loglik_acc)
  uint32_t first_sweep;         // 1-based number of the first sweep of this launch
  uint32_t n_sweeps;
  long long run_step0;          // step index (1-based, within the run) of the first sweep of this launch
  long long burnin, thin;       // keep iff step > burnin && (step-1-burnin) % thin == 0   (bi:402)
  long long chunk_base;         // first draw index held by the chunk
  int store_zt_last;
  int* error_flag;
  unsigned int* barrier;        // [2] grid barrier: arrival counter, generation
};

// Sense-reversing grid barrier on two global words (all blocks are co-resident: cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen = bar + 1;
    const unsigned int my = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1u) {
      *bar = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == my) { }
    }
    __threadfence();
  }
  __syncthreads();
}

// Accumulator rotation: sweep s ADDS its statistics to slot s%3, the level-2 draw that follows (bivariate: at the start
// of sweep s+1; trivariate: at the end of sweep s) READS slot s%3, and slot (s+1)%3 is ZEROED during sweep s, one full
// barrier after its last reader and one before its next writer.
template <int D, int MODE>
__global__ void __launch_bounds__(SWEEP_THREADS, CLV_MINBLOCKS) k_persistent(PersistArgs pa) {
  extern __shared__ long long s_priv[];
  __shared__ double s_beta[MAXK * MAXD];
  __shared__ double s_tab[EXP_N];
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  __shared__ ChainParams s_cp;
  __shared__ double s_stash[E32_STASH * SWEEP_THREADS];           // the exact path's inputs (E32)
  __shared__ Level2Scratch sc;
  __shared__ Level2Const lc;
  __shared__ Level2Variates lv[2];         // variates of sweep s live in lv[s & 1]; drawn one sweep ahead by warp 1
  const SweepArgs& a = pa.sw;
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = mc.K;
  const int nstat = K * D + D * (D + 1) / 2;
  const unsigned int nblocks = gridDim.x * gridDim.y;
  const long long ntiles = (mc.N + SWEEP_THREADS - 1) / SWEEP_THREADS;
  const PhiloxKey key = seed_key(a.seed);
  const uint32_t c3 = dom_word(DOM_SAMPLER, a.chain_offset + (uint32_t)chain);
  const uint32_t c3_l2 = dom_word(DOM_LEVEL2, a.chain_offset + (uint32_t)chain);
  const long long csz = (long long)gridDim.y * NSTAT_MAX;
  for (int t = tid; t < EXP_N; t += SWEEP_THREADS) s_tab[t] = c_exptab[t];
  clear_stats(s_priv, nstat);
  load_level2_const<D>(mc, lc, tid, SWEEP_THREADS);
  {
    const double* src = reinterpret_cast<const double*>(&pa.params[chain]);
    double* dst = reinterpret_cast<double*>(&s_cp);
    for (int t = tid; t < (int)(sizeof(ChainParams) / sizeof(double)); t += SWEEP_THREADS) dst[t] = src[t];
  }
  __syncthreads();
  // the variates of the first level-2 draw of this launch (they depend on (seed, chain, sweep) only)
  if (warp == 1) level2_variates<D>(lc, lv[pa.first_sweep & 1u], key, c3_l2, pa.first_sweep, 0, nullptr, nullptr, nullptr, lane);

  auto level2_phase = [&](uint32_t sweep, long long draw_index, unsigned long long* slot_read, bool more) {
    // every block of the chain draws the same (beta, Sigma) from the same totals and the same Philox key: warp 0 does
    // the algebra of this sweep while warp 1 (idle otherwise) prepares the variates of the next one
    if (warp == 0) {
      for (int t = lane; t < nstat; t += 32)
        sc.st[t] = (double)(long long)__ldcg(&slot_read[chain * NSTAT_MAX + t]) * lc.fx_inv;
      __syncwarp();
      level2_algebra<D>(lc, sc, lv[sweep & 1u], s_cp, lane);
      if (lane == 0 && !sc.ok) *pa.error_flag = 1;
      if (blockIdx.x == 0 && draw_index >= 0)
        write_level2_row<D>(K, s_cp, pa.level2_draws + ((long long)chain * pa.n_draws + draw_index) * (D * K + D * (D + 1) / 2), lane);
    } else if (warp == 1 && more) {
      level2_variates<D>(lc, lv[(sweep + 1u) & 1u], key, c3_l2, sweep + 1u, 0, nullptr, nullptr, nullptr, lane);
    }
    __syncthreads();
  };

  for (uint32_t it = 0; it < pa.n_sweeps; ++it) {
    const uint32_t sweep = pa.first_sweep + it;
    const long long step = pa.run_step0 + it;
    const bool kept = step > pa.burnin && (step - 1 - pa.burnin) % pa.thin == 0;
    const long long draw = kept ? (step - 1 - pa.burnin) / pa.thin : -1;
    const bool more = it + 1 < pa.n_sweeps;
    unsigned long long* slot_prev = pa.acc3 + (long long)((sweep + 2u) % 3u) * csz;   // (sweep-1) % 3
    unsigned long long* slot_cur = pa.acc3 + (long long)(sweep % 3u) * csz;
    unsigned long long* slot_next = pa.acc3 + (long long)((sweep + 1u) % 3u) * csz;
    if (D == 2) {
      grid_barrier(pa.barrier, nblocks);         // statistics of sweep-1 (or the initial state) are complete
      level2_phase(sweep, draw, slot_prev, more);      // bi:393
    }
    if (blockIdx.x == 0)
      for (int t = tid; t < nstat; t += SWEEP_THREADS) slot_next[chain * NSTAT_MAX + t] = 0ull;
    for (int t = tid; t < K * D; t += SWEEP_THREADS) s_beta[t] = s_cp.beta[t];
    for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
    __syncthreads();
    SweepStep sw;
    sw.sweep = sweep; sw.keep = kept; sw.store_zt = (pa.store_zt_last && it + 1 == pa.n_sweeps);
    sw.slot = kept ? draw - pa.chunk_base : 0; sw.chunk_cap = a.chunk_cap; sw.draws = a.draws; sw.draw_index = draw;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      sweep_tile<D, MODE>(a, mc, s_cp, s_beta, s_tab, s_priv, sw, chain, tile, c3, s_stash);
    flush_stats(s_priv, s_acc, nstat, kept);
    __syncthreads();
    for (int t = tid; t < nstat; t += SWEEP_THREADS)
      if (s_acc[t]) atomicAdd(&slot_cur[chain * NSTAT_MAX + t], s_acc[t]);
    if (kept && tid == 0)
      atomicAdd(reinterpret_cast<unsigned long long*>(&a.loglik_acc[chain * a.loglik_stride + draw]), s_acc[NSTAT_MAX]);
    if (D == 3) {
      grid_barrier(pa.barrier, nblocks);         // tri:529-536: level-2 closes the sweep
      level2_phase(sweep, draw, slot_cur, more);
    }
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    double* dst = reinterpret_cast<double*>(&pa.params[chain]);
    const double* src = reinterpret_cast<const double*>(&s_cp);
    for (int t = tid; t < (int)(sizeof(ChainParams) / sizeof(double)); t += SWEEP_THREADS) dst[t] = src[t];
  }
}

}  // namespace clv
