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

namespace clv {

constexpr int MAXK = 16;
constexpr int MAXD = 3;
constexpr int NSTAT_MAX = MAXK * MAXD + 6;
constexpr int SWEEP_THREADS = 128;

enum : int { MODE_FAST = 0, MODE_STRICT = 1, MODE_INJECT = 2 };

// Per-chain level-2 state, written by k_level2 / k_derive_params, read by k_sweep.
struct ChainParams {
  double beta[MAXK * MAXD];  // [k*D + d]
  double Sigma[MAXD * MAXD];
  double P00, P01, P11;      // entries of inv(Sigma) used by the level-1 target (bi:303-305; tri:419-426, Q4)
  double eta_post_var, eta_sd;  // tri:325-326
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
  int store_zt;                // write z/tau state arrays
  // injected variates (MODE_INJECT)
  const double *u_z, *e_tau, *u_tau, *t3_l, *t3_m, *u_acc, *n_eta;
};

__device__ __forceinline__ long long to_fx(double v, double scale) { return __double2ll_rn(v * scale); }

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Level-1 target, bi:291-310.  Tz = z*T_cal + (1-z)*tau, omz = 1-z.
__device__ __forceinline__ double log_post(double ll, double lm, double xd, double omz, double Tz, double m0,
                                           double m1, double P00, double P01, double P11) {
  double dl = ll - m0, dm = lm - m1;
  double lik = xd * ll + omz * lm - (exp(ll) + exp(lm)) * Tz;
  double prior = -0.5 * (dl * dl * P00 + 2.0 * dl * dm * P01 + dm * dm * P11);
  double res = lik + prior;
  return (lm > 5.0) ? -CUDART_INF : res;
}

// MH accept rule of bi:329-330: exp(prop - cur) > u, NaN compares false.
__device__ __forceinline__ bool mh_accept(double d, double u) {
  if (d >= 0.0) return true;       // exp(d) >= 1 > u
  return exp(d) > u;               // NaN -> false; d = -inf -> 0 > u false
}

template <int D>
__device__ __forceinline__ void accumulate_stats(const ModelConst& mc, const double* __restrict__ Xc, long long N,
                                                 long long i, bool valid, double yc0, double yc1, double yc2,
                                                 unsigned long long* s_acc, int lane) {
  const double sc = mc.fx_scale;
  const int K = mc.K;
  double y[3] = {yc0, yc1, yc2};
  for (int k = 0; k < K; ++k) {
    double xk = 0.0;
    if (valid) xk = (k == 0) ? 1.0 : Xc[(long long)(k - 1) * N + i];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      long long v = warp_sum_ll(valid ? to_fx(xk * y[d], sc) : 0ll);
      if (lane == 0) atomicAdd(&s_acc[k * D + d], (unsigned long long)v);
    }
  }
  int t = K * D;
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int e = d; e < D; ++e) {
      long long v = warp_sum_ll(valid ? to_fx(y[d] * y[e], sc) : 0ll);
      if (lane == 0) atomicAdd(&s_acc[t], (unsigned long long)v);
      ++t;
    }
}

template <int D, int MODE>
__global__ void __launch_bounds__(SWEEP_THREADS) k_sweep(SweepArgs a) {
  __shared__ double s_beta[MAXK * MAXD];
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = mc.K, S = mc.S;
  const long long N = mc.N;
  const ChainParams& cp = a.params[chain];
  for (int t = tid; t < K * D; t += SWEEP_THREADS) s_beta[t] = cp.beta[t];
  for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
  __syncthreads();
  const double P00 = cp.P00, P01 = cp.P01, P11 = cp.P11;
  const double s_l = cp.Sigma[0], s_m = cp.Sigma[D + 1];   // proposal scales are variances (bi:316-317, Q2)
  const PhiloxKey key = chain_key(a.seed, a.chain_offset + (uint32_t)chain);
  const bool keep = a.slot >= 0;
  const long long ntiles = (N + SWEEP_THREADS - 1) / SWEEP_THREADS;
  const long long cN = (long long)chain * N;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
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
      const double lam = exp(ll), mu = exp(lm);
      double uz, ut, et;
      if (MODE == MODE_INJECT) {
        uz = a.u_z[cN + i];
        ut = a.u_tau[cN + i];
        et = a.e_tau[cN + i];
      } else {
        uint4 r = philox4x32_10(gid, a.sweep, 0u, DOM_SAMPLER, key);
        uz = u53(r.x, r.y);
        ut = u53(r.z, r.w);
        et = 0.0;
      }
      const double ml = mu + lam;
      const double e = exp(-(ml * (T - tx)));
      const double pa = (ml * e) / (ml * e + mu * (1.0 - e));
      const bool alive = uz < pa;
      double tau;
      if (alive) {
        if (MODE != MODE_INJECT) et = -log(ut);
        tau = T + (1.0 / mu) * et;
      } else {
        double mtx = fmin(700.0, ml * tx), mT = fmin(700.0, ml * T);
        tau = -log((1.0 - ut) * exp(-mtx) + ut * exp(-mT)) / ml;
      }
      const double zf = alive ? 1.0 : 0.0;
      const double omz = 1.0 - zf;
      const double Tz = alive ? T : tau;          // z*T_cal + (1-z)*tau, bi:298
      // ---- S Metropolis steps (bi:312-335) -------------------------------------------------------
      double cur = log_post(ll, lm, xd, omz, Tz, m0, m1, P00, P01, P11);
      for (int s = 0; s < S; ++s) {
        double tl, tm, ua;
        if (MODE == MODE_INJECT) {
          long long o = ((long long)chain * S + s) * N + i;
          tl = a.t3_l[o];
          tm = a.t3_m[o];
          ua = a.u_acc[o];
        } else {
          uint4 ra = philox4x32_10(gid, a.sweep, 1u + 2u * s, DOM_SAMPLER, key);
          uint4 rb = philox4x32_10(gid, a.sweep, 2u + 2u * s, DOM_SAMPLER, key);
          if (MODE == MODE_STRICT) {
            tl = t3_strict(ra.x, ra.y, ra.z);
            tm = t3_strict(ra.w, rb.x, rb.y);
          } else {
            tl = (double)t3_fast(ra.x, ra.y, ra.z);
            tm = (double)t3_fast(ra.w, rb.x, rb.y);
          }
          ua = u32d(rb.z);
        }
        double pl = fmin(fmax(ll + s_l * tl, -70.0), 70.0);     // bi:318-324
        double pm = fmin(fmax(lm + s_m * tm, -70.0), 70.0);
        double prop = log_post(pl, pm, xd, omz, Tz, m0, m1, P00, P01, P11);
        if (mh_accept(prop - cur, ua)) {
          ll = pl;
          lm = pm;
          cur = prop;
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
          normal_pair_u53(philox4x32_10(gid, a.sweep, 1u + 2u * (uint32_t)S, DOM_SAMPLER, key), &n, &ns);
        }
        const double prior_var = cp.Sigma[8];
        double post_mean = cp.eta_post_var * (a.log_s[i] / mc.omega2 + m2 / prior_var);
        le = post_mean + cp.eta_sd * n;
        a.le[cN + i] = le;
      }
      if (a.store_zt) {
        a.z[cN + i] = zf;
        a.tau[cN + i] = tau;
      }
      // ---- kept draw: lambda, mu, tau, z(, eta)   bi:407-410, tri:544-548 -------------------------
      if (keep) {
        const double lam_n = exp(ll), mu_n = exp(lm);
        constexpr int NC = (D == 2) ? 4 : 5;
        double* o = a.draws + (((long long)chain * a.chunk_cap + a.slot) * N + i) * NC;
        if (D == 2) {
          reinterpret_cast<double2*>(o)[0] = make_double2(lam_n, mu_n);
          reinterpret_cast<double2*>(o)[1] = make_double2(tau, zf);
        } else {
          o[0] = lam_n; o[1] = mu_n; o[2] = tau; o[3] = zf; o[4] = exp(le);
        }
        lik = xd * ll + omz * lm - (lam_n + mu_n) * Tz;         // bi:423-427
        lik = fmin(fmax(lik, -1048576.0), 1048576.0);
      }
      yc0 = ll - mc.center[0];
      yc1 = lm - mc.center[1];
      if (D == 3) yc2 = le - mc.center[2];
    }
    accumulate_stats<D>(mc, a.Xc, N, i, valid, yc0, yc1, yc2, s_acc, lane);
    if (keep) {
      long long v = warp_sum_ll(valid ? to_fx(lik, mc.ll_scale) : 0ll);
      if (lane == 0) atomicAdd(&s_acc[NSTAT_MAX], (unsigned long long)v);
    }
  }
  __syncthreads();
  const int nstat = K * D + D * (D + 1) / 2;
  for (int t = tid; t < nstat; t += SWEEP_THREADS)
    if (s_acc[t]) atomicAdd(&a.acc[chain * NSTAT_MAX + t], s_acc[t]);
  if (keep && tid == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&a.loglik_acc[chain * a.loglik_stride + a.draw_index]),
              s_acc[NSTAT_MAX]);
}

// Statistics of the current state only (first level-2 draw of the bivariate order, bi:393).
template <int D>
__global__ void __launch_bounds__(SWEEP_THREADS) k_stats_only(SweepArgs a) {
  __shared__ unsigned long long s_acc[NSTAT_MAX + 1];
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const long long N = mc.N, cN = (long long)chain * N;
  for (int t = tid; t < NSTAT_MAX + 1; t += SWEEP_THREADS) s_acc[t] = 0ull;
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
    accumulate_stats<D>(mc, a.Xc, N, i, valid, yc0, yc1, yc2, s_acc, lane);
  }
  __syncthreads();
  const int nstat = mc.K * D + D * (D + 1) / 2;
  for (int t = tid; t < nstat; t += SWEEP_THREADS)
    if (s_acc[t]) atomicAdd(&a.acc[chain * NSTAT_MAX + t], s_acc[t]);
}

// ------------------------------------------------------------------------------------------------
// level 2
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ inline bool chol_lower(const double* A, double* L) {
  bool ok = true;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * D + j];
      for (int m = 0; m < j; ++m) s -= L[i * D + m] * L[j * D + m];
      if (i == j) {
        if (!(s > 0.0) || !isfinite(s)) { ok = false; s = 1.0; }
        L[i * D + i] = sqrt(s);
      } else {
        L[i * D + j] = s / L[j * D + j];
      }
    }
  for (int i = 0; i < D; ++i)
    for (int j = i + 1; j < D; ++j) L[i * D + j] = 0.0;
  return ok;
}

// P = inv(Sigma) (entries 00, 01, 11) and the eta conjugate scalars.
template <int D>
__device__ inline void derive_params(ChainParams& cp, double omega2) {
  const double* S = cp.Sigma;
  if (D == 2) {
    double det = S[0] * S[3] - S[1] * S[2];
    cp.P00 = S[3] / det;
    cp.P01 = -S[1] / det;
    cp.P11 = S[0] / det;
    cp.eta_post_var = 0.0;
    cp.eta_sd = 0.0;
  } else {
    double c00 = S[4] * S[8] - S[5] * S[7];
    double c01 = S[5] * S[6] - S[3] * S[8];
    double c02 = S[3] * S[7] - S[4] * S[6];
    double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
    cp.P00 = c00 / det;
    cp.P01 = (S[2] * S[7] - S[1] * S[8]) / det;
    cp.P11 = (S[0] * S[8] - S[2] * S[6]) / det;
    double post_precision = 1.0 / omega2 + 1.0 / S[8];      // tri:325
    cp.eta_post_var = 1.0 / post_precision;
    cp.eta_sd = sqrt(cp.eta_post_var);
  }
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
};

// Conjugate multivariate regression draw (bi:233-262, tri:340-380) from the reduced statistics.
// Works in centred responses y - c (c = prior intercept row), which leaves E and beta - B0 unchanged.
template <int D>
__global__ void __launch_bounds__(32) k_level2(Level2Args a) {
  __shared__ double st[NSTAT_MAX];
  const ModelConst& mc = *a.mc;
  const int chain = blockIdx.x, lane = threadIdx.x;
  const int K = mc.K;
  const int nstat = K * D + D * (D + 1) / 2;
  for (int t = lane; t < nstat; t += 32) {
    unsigned long long* p = &a.acc[chain * NSTAT_MAX + t];
    st[t] = (double)(long long)(*p) * mc.fx_inv;
    *p = 0ull;
  }
  // variates: every lane draws its share (fp64 Philox transforms are long dependent chains)
  __shared__ double s_trn[3], s_chi[3], s_zb[MAXD * MAXK];
  const PhiloxKey key = chain_key(a.seed, a.chain_offset + (uint32_t)chain);
  {
    const int ntril = D * (D - 1) / 2, nb = D * K;
    for (int t = lane; t < ntril + D + nb; t += 32) {
      if (t < ntril) {
        s_trn[t] = a.injected ? a.iw_norm[chain * ntril + t] : level2_normal(key, a.sweep, (uint32_t)t);
      } else if (t < ntril + D) {
        int i = t - ntril;
        s_chi[i] = a.injected ? a.iw_chi2[chain * D + i] : level2_chi2(key, a.sweep, 16u + i, mc.nu_n - D + 1 + i);
      } else {
        int j = t - ntril - D;
        s_zb[j] = a.injected ? a.beta_norm[chain * nb + j] : level2_normal(key, a.sweep, 32u + (uint32_t)j);
      }
    }
  }
  __syncwarp();
  if (lane != 0) return;
  ChainParams& cp = a.params[chain];

  // R = X'Yc + A0 B0c ; Bc = V R                               bi:250
  double R[MAXK * MAXD], Bc[MAXK * MAXD];
  for (int t = 0; t < K * D; ++t) R[t] = st[t] + mc.A0B0c[t];
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < D; ++d) {
      double s = 0.0;
      for (int m = 0; m < K; ++m) s += mc.V[k * K + m] * R[m * D + d];
      Bc[k * D + d] = s;
    }
  // S_n = S0 + E'E + C'A0C = Q0 + Yc'Yc - Bc' R                bi:253-255
  double Sn[D * D];
  {
    int t = K * D;
    for (int d = 0; d < D; ++d)
      for (int e = d; e < D; ++e) {
        Sn[d * D + e] = st[t];
        Sn[e * D + d] = st[t];
        ++t;
      }
  }
  for (int d = 0; d < D; ++d)
    for (int e = 0; e < D; ++e) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += Bc[k * D + d] * R[k * D + e];
      Sn[d * D + e] += mc.Q0[d * D + e] - s;
    }
  for (int d = 0; d < D; ++d)
    for (int e = d + 1; e < D; ++e) {
      double m = 0.5 * (Sn[d * D + e] + Sn[e * D + d]);
      Sn[d * D + e] = m;
      Sn[e * D + d] = m;
    }
  double C[D * D];
  bool ok = chol_lower<D>(Sn, C);
  // Sigma ~ IW(nu_n, S_n): scipy's Bartlett construction (bi:258)
  double A[D * D];
  for (int t = 0; t < D * D; ++t) A[t] = 0.0;
  {
    int t = 0;
    for (int i = 1; i < D; ++i)
      for (int j = 0; j < i; ++j) {
        A[i * D + j] = s_trn[t];
        ++t;
      }
    for (int i = 0; i < D; ++i) {
      A[i * D + i] = sqrt(s_chi[i]);                              // chi2(nu_n - D + 1 + i)
    }
  }
  // CA = C A^-1 (lower triangular)  =>  Sigma = CA CA'
  double CA[D * D];
  for (int r = 0; r < D; ++r)
    for (int j = D - 1; j >= 0; --j) {
      double s = C[r * D + j];
      for (int m = j + 1; m < D; ++m) s -= CA[r * D + m] * A[m * D + j];
      CA[r * D + j] = s / A[j * D + j];
    }
  for (int d = 0; d < D; ++d)
    for (int e = 0; e < D; ++e) {
      double s = 0.0;
      for (int m = 0; m < D; ++m) s += CA[d * D + m] * CA[e * D + m];
      cp.Sigma[d * D + e] = s;
      if (!isfinite(s)) ok = false;
    }
  // beta | Sigma: noise = kron(chol Sigma, chol V) z, z ordered d*K+k           bi:261
  double W[MAXD * MAXK];
  for (int d = 0; d < D; ++d)
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (int m = 0; m <= k; ++m) s += mc.LV[k * K + m] * s_zb[d * K + m];
      W[d * K + k] = s;
    }
  double Ef[MAXD * MAXK];  // noise in kron(Sigma, V) order d*K+k
  for (int d = 0; d < D; ++d)
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (int m = 0; m <= d; ++m) s += CA[d * D + m] * W[m * K + k];
      Ef[d * K + k] = s;
    }
  for (int k = 0; k < K; ++k)
    for (int d = 0; d < D; ++d) {
      double bh = Bc[k * D + d] + (k == 0 ? mc.center[d] : 0.0);   // B_hat = Bc + e0 c'
      int j = k * D + d;
      double nz = (mc.compat == 0) ? Ef[j] : Ef[d * K + k];        // Q1: reference adds kron-ordered noise to ravel()
      cp.beta[j] = bh + nz;
    }
  derive_params<D>(cp, mc.omega2);
  cp.status = ok ? 0 : 1;
  if (!ok) *a.error_flag = 1;
  if (a.draw_index >= 0) {
    const int P = D * K + D * (D + 1) / 2;
    double* o = a.level2_draws + ((long long)chain * a.n_draws + a.draw_index) * P;
    for (int d = 0; d < D; ++d)
      for (int k = 0; k < K; ++k) o[d * K + k] = cp.beta[k * D + d];   // beta.T.ravel()   bi:411
    int t = D * K;
    for (int d = 0; d < D; ++d)
      for (int e = d; e < D; ++e) o[t++] = cp.Sigma[d * D + e];        // bi:412, tri:550-554
  }
}

}  // namespace clv
