/* clv_b200.h — C-ABI of the B200-native Abe (2009/2015) hierarchical Pareto/NBD sampler.
 *
 * The reference (lucagem29/mcmc_clv_model) is pure Python and has no native boundary; the
 * functions below are what a maintainer binds (ctypes) behind the
 * reference's own entry points.  Each entry cites the reference interface it replaces
 * ("bi" = src/models/bivariate/mcmc.py, "tri" = src/models/trivariate/mcmc.py).
 *
 * Conventions
 *   - every function returns 0 on success or a negative clv_status; nothing throws across
 *     the ABI; clv_last_error() gives the message of the last failure on that handle
 *     (clv_last_error(NULL) for failures of clv_create / handle-less calls).
 *   - host pointers are owned by the caller and need only stay valid during the call.
 *   - the library owns all device memory of a handle; *_dev accessors expose raw device
 *     pointers (valid until clv_destroy) for zero-copy consumers (torch.from_blob etc.).
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - all matrices are row-major doubles unless noted.
 */
#ifndef CLV_B200_H
#define CLV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLV_ABI_VERSION 2
#define CLV_MAX_K 16 /* design-matrix columns incl. intercept */

typedef enum {
  CLV_OK = 0,
  CLV_ERR_ARG = -1,     /* bad argument / call order */
  CLV_ERR_CUDA = -2,    /* CUDA runtime failure (no device, OOM, launch error) */
  CLV_ERR_NUMERIC = -3, /* non-finite level-2 statistics, non-PD scale matrix */
  CLV_ERR_COMM = -4,    /* NCCL failure */
  CLV_ERR_STATE = -5    /* handle not initialised for this call */
} clv_status;

typedef enum {
  CLV_RNG_PHILOX_FAST = 0,   /* Philox4x32-10; proposal variates through the SFU in fp32; target/accept in fp64 */
  CLV_RNG_PHILOX_STRICT = 1, /* Philox4x32-10; every transform in fp64 (replayable by oracle/philox_np.py) */
  CLV_RNG_INJECTED = 2       /* variates supplied by the caller (clv_sweep_injected) */
} clv_rng_mode;

typedef enum {
  CLV_COMPAT_REFERENCE = 0, /* reproduce the reference's beta-covariance ordering (bi:261) */
  CLV_COMPAT_PAPER = 1      /* matrix-normal beta | Sigma as in Abe (2009) */
} clv_compat;

typedef enum {
  CLV_SWEEP_AUTO = 0,      /* the faster of the two for the problem; currently always stream */
  CLV_SWEEP_STREAM = 1,    /* two kernels per sweep on a stream (k_sweep, k_level2), chained by programmatic dependent launch */
  CLV_SWEEP_RESERVED = 2,  /* treated as CLV_SWEEP_STREAM */
  CLV_SWEEP_PERSISTENT = 3 /* one cooperative kernel, grid barrier per sweep (single shard only) */
} clv_sweep_mode;

typedef struct clv_sampler clv_sampler;

/* Replaces the keyword arguments of mcmc_draw_parameters (bi:437-447) /
 * mcmc_draw_parameters_rfm_m (tri:580-590) plus the placement the reference does not have. */
typedef struct {
  int32_t model_dim;    /* D: 2 = bivariate (lambda, mu); 3 = trivariate (+ eta) */
  int32_t n_cov;        /* K: columns of the design matrix, intercept included (bi:469-472) */
  int32_t n_chains;     /* chains held by this handle (bi:485) */
  int32_t chain_offset; /* global index of this handle's chain 0: chain c draws from key seed+chain_offset+c (bi:486) */
  int32_t n_mh_steps;   /* bi:446 */
  int32_t rng_mode;     /* clv_rng_mode */
  int32_t compat;       /* clv_compat */
  int32_t sweep_mode;   /* clv_sweep_mode */
  int32_t device;       /* CUDA ordinal */
  int32_t reserved;
  int64_t n_local;      /* customers in this shard */
  int64_t n_global;     /* customers in the whole problem (== n_local when unsharded) */
  int64_t gid_offset;   /* global id of local customer 0 (RNG counters use global ids) */
  uint64_t seed;        /* bi:444 */
} clv_config;

/* Global (all shards) initialisation statistics, bi:367-374 / tri:488-499.  The host computes them
 * (exact integer sums, mcmc_clv_model_b200/hostmath.py) so they do not depend on the shard count. */
typedef struct {
  double lam_init;      /* mean(x) / mean(where(t_x==0, T_cal, t_x))            bi:368 */
  double mean_mu_init;  /* mean(1/(t_x + 0.5/lam_init))                          bi:370,374 */
  double mean_log_s;    /* mean(log_s)                     (D=3)                 tri:499 */
  double omega2;        /* var(log_s, ddof=1)              (D=3)                 tri:494 */
  double max_abs_x;     /* max |X| over all customers and columns (fixed-point headroom) */
  const double* xtx;    /* K*K, X'X over all customers                           bi:248 */
} clv_init_stats;

typedef void (*clv_progress_cb)(void* user, int64_t step, int64_t total_steps);

/* ---- lifecycle -------------------------------------------------------------------------- */
int clv_abi_version(void);
int clv_create(clv_sampler** out, const clv_config* cfg);
void clv_destroy(clv_sampler* h);
const char* clv_last_error(const clv_sampler* h);

/* ---- inputs: the CBS frame of bi:459-470 / tri:612-618 ------------------------------------ */
/* x, t_x, T_cal: columns "x", "t_x", "T_cal"; X: n_local x K row-major, column 0 == 1 (bi:468-470);
 * log_s: column "log_s" (tri:329), NULL for D=2. */
int clv_set_data(clv_sampler* h, const int32_t* x, const double* t_x, const double* T_cal,
                 const double* X, const double* log_s);
/* The same frame handed over column by column, the way the DataFrame of bi:459-470 holds it: cov[k] (k = 0 .. K-2) is the
 * column named covariates[k] (n_local doubles); the intercept column the reference prepends (bi:468-470) is implicit.
 * No row-major matrix has to be assembled on the host and the intercept's 8 B per customer never cross PCIe. */
int clv_set_data_columns(clv_sampler* h, const int32_t* x, const double* t_x, const double* T_cal,
                         const double* const* cov, const double* log_s);
/* hyper dict of bi:474-479 / tri:622-626: beta0 K x D, A0 K x K, nu0, gamma0 D x D.
 * Row 0 of beta0 is overwritten from the init statistics exactly as bi:373-374 / tri:497-499 do. */
int clv_set_hyper(clv_sampler* h, const double* beta0, const double* A0, double nu0, const double* gamma0);
/* Initial lambda, mu, eta, beta, Sigma of bi:367-379 / tri:488-504.  stats == NULL: the statistics are computed
 * on the device as exact integer sums (and all-reduced over the communicator when clv_comm_init was called first),
 * bit-identical to mcmc_clv_model_b200/hostmath.py for any sharding. */
int clv_init_state(clv_sampler* h, const clv_init_stats* stats);
/* The statistics the last clv_init_state(h, NULL) computed; xtx_out (nullable) receives K*K doubles. */
int clv_get_init_stats(clv_sampler* h, clv_init_stats* out, double* xtx_out);
/* Customer-sharded mode: join the NCCL communicator used for the per-sweep all-reduce of the
 * level-2 sufficient statistics.  unique_id is the 128-byte ncclUniqueId from clv_comm_unique_id
 * on rank 0, broadcast by the host (torch.distributed).  The communicator is created once per
 * (device, rank, world) and reused by later handles of the same process (unique_id is then ignored). */
int clv_comm_unique_id(void* out128);
int clv_comm_init(clv_sampler* h, const void* unique_id128, int rank, int world);

/* Optional, after clv_comm_init + clv_init_state: replace the per-sweep NCCL all-reduce by a one-shot all-reduce over
 * peer memory FUSED INTO the level-2 kernel.  Each rank stores the halves of its int64 partial sums into every rank's
 * mailbox over NVLink/NVSwitch as 8-byte words of 32 payload bits + a 32-bit (init epoch, sweep) tag -- an aligned 8-byte
 * store is atomic, so a word is its own arrival flag: no fence, no flag round trip -- and adds the words it received.
 * Every rank exports its mailbox (64-byte cudaIpcMemHandle), the host all-gathers the handles, every rank connects.
 * clv_p2p_connect clears the rank's own mailbox: the caller must synchronise the ranks (a host barrier) between
 * clv_p2p_connect and the first sweep.  A rank that does not receive its peers' words within CLV_P2P_TIMEOUT_S seconds
 * (environment, default 120) raises CLV_ERR_COMM; the remaining sweeps of the call are skipped.
 * world <= 16, one process per GPU. */
int clv_p2p_export(clv_sampler* h, void* handle64);
int clv_p2p_connect(clv_sampler* h, const void* handles /* world x 64 bytes */, int rank, int world);
/* Mailboxes and their peer mappings live for the life of the process; returns 1 when a connected set already exists for
 * this (device, rank, world, n_chains): clv_p2p_connect can then be called with handles == NULL. */
int clv_p2p_is_cached(clv_sampler* h, int rank, int world);

/* ---- the chain driver, _run_chain (bi:346-431, tri:465-574) ------------------------------- */
/* Runs burnin+mcmc sweeps on every chain of the handle and writes, per chain c,
 *   level1 [c][n_draws][n_local][4|5] = lambda, mu, tau, z(, eta)       (bi:407-410, tri:544-548)
 *   level2 [c][n_draws][D*K + D(D+1)/2] = beta.T.ravel(), triu(Sigma)   (bi:411-412, tri:549-554)
 *   loglik [c][n_draws]  = SUM over local customers of the likelihood part (bi:423-427); the host divides
 *                          by n_global (and adds shards) to get the reference's per-draw mean (bi:428)
 * with n_draws = (mcmc-1)/thin + 1 (bi:360).  level1 may be NULL (draws are then not stored).
 * cb (nullable) is called every `trace` sweeps (bi:384-385).  May be called repeatedly; sweeps continue. */
int clv_run(clv_sampler* h, int64_t burnin, int64_t mcmc, int64_t thin, double* level1, double* level2,
            double* loglik, clv_progress_cb cb, void* user, int64_t trace);
/* Same as clv_run, but the level-1 draws stay on the device (they must fit its draw buffer): the input of
 * clv_forecast_resident and of zero-copy consumers (clv_resident_draws -> [chains][n_draws][n_local][4|5]). */
int clv_run_resident(clv_sampler* h, int64_t burnin, int64_t mcmc, int64_t thin, double* level2, double* loglik,
                     clv_progress_cb cb, void* user, int64_t trace);
int clv_resident_draws(clv_sampler* h, const double** level1_dev, int64_t* n_draws);
/* Advance n sweeps without storing draws (burn-in, benchmarks).  Asynchronous unless sync != 0; a synchronous call
 * also leaves z / tau of its last sweep in the state arrays clv_get_state reads. */
int clv_advance(clv_sampler* h, int64_t n_sweeps, int sync);
/* Same, synchronous, bracketed by CUDA events on the handle's stream: *elapsed_ms = device time of the n sweeps. */
int clv_advance_timed(clv_sampler* h, int64_t n_sweeps, double* elapsed_ms);
int64_t clv_sweeps_done(const clv_sampler* h);
/* Resume: after clv_set_state on a fresh handle, set the number of sweeps the restored chains have already made, so
 * that the Philox counters (sweep = sweeps_done + 1, ...) continue where the checkpointed run stopped. */
int clv_set_sweeps_done(clv_sampler* h, int64_t n);
/* kernels launched by this handle so far (bench.py's gpu_launches) */
int64_t clv_kernel_launches(const clv_sampler* h);
/* CUDA-event time (ms) of the level-1 sweep kernel summed over the sweeps since the last call (0 if
 * timing is off); enable with clv_set_timing(h, 1). */
int clv_set_timing(clv_sampler* h, int on);
int clv_kernel_time_ms(clv_sampler* h, double* sweep_kernel_ms, double* level2_kernel_ms, int64_t* n_sweeps);

/* ---- state access (tests, checkpoint/resume) ---------------------------------------------- */
/* Per chain: log lambda, log mu, log eta (NULL for D=2), z (0/1), tau — each n_local doubles;
 * beta K x D, Sigma D x D.  Any pointer may be NULL. */
int clv_get_state(clv_sampler* h, int chain, double* log_lambda, double* log_mu, double* log_eta,
                  double* z, double* tau, double* beta, double* Sigma);
int clv_set_state(clv_sampler* h, int chain, const double* log_lambda, const double* log_mu,
                  const double* log_eta, const double* beta, const double* Sigma);

/* One sweep of every chain with caller-supplied variates, arrays chain-major:
 * u_z, e_tau, u_tau [chains][n_local]; t3_l, t3_m, u_acc [chains][S][n_local]; n_eta [chains][n_local] (D=3);
 * iw_norm [chains][D(D-1)/2]; iw_chi2 [chains][D]; beta_norm [chains][D*K].
 * The test hook behind "identical injected streams => z bit-exact, continuous 1e-6" (BASELINE.json). */
typedef struct {
  const double *u_z, *e_tau, *u_tau, *t3_l, *t3_m, *u_acc, *n_eta, *iw_norm, *iw_chi2, *beta_norm;
} clv_injected;
int clv_sweep_injected(clv_sampler* h, const clv_injected* v, int keep, double* level1, double* level2,
                       double* loglik);

/* ---- forecast, draw_future_transactions (bi:506-546, tri:660-749) -------------------------- */
typedef struct {
  int32_t device;
  int32_t ncol;           /* 4 (bivariate draws) or 5 (trivariate) */
  int64_t n_draws_total;  /* chains * n_draws, chain-major as bi:530-531 */
  int64_t n_customers;
  int64_t gid_offset;     /* global id of customer 0 of this shard */
  int64_t draw_offset;    /* global index of draw 0 of this call */
  double T_star;          /* bi:506 */
  uint64_t seed;          /* bi:526 */
  int32_t simulate_spend; /* tri:665 */
  int32_t reserved;
  double sigma_s;         /* tri:666 */
} clv_forecast_config;
/* level1: host [n_draws_total][n_customers][ncol]; T_cal host [n_customers];
 * x_star out host int64 [n_draws_total][n_customers]; spend out (nullable) same shape, double. */
int clv_forecast(const clv_forecast_config* cfg, const double* level1, const double* T_cal,
                 int64_t* x_star, double* spend);
/* Same with device pointers on `stream` (a cudaStream_t passed as void*; NULL = default stream). */
int clv_forecast_dev(const clv_forecast_config* cfg, const double* level1_dev, const double* T_cal_dev,
                     int64_t* x_star_dev, double* spend_dev, void* stream);
/* Injected: u [n_draws_total][n_customers] uniforms (Poisson by CDF inversion); eps (nullable) n_eps flat
 * per-transaction normals, eps_offset [n_draws_total][n_customers] index of each cell's first normal. */
int clv_forecast_injected(const clv_forecast_config* cfg, const double* level1, const double* T_cal,
                          const double* u, const double* eps, int64_t n_eps, const int64_t* eps_offset,
                          int64_t* x_star, double* spend);
/* Forecast straight from the draws still resident on the device after the last clv_run (no PCIe):
 * x_star host int64 [chains*n_draws][n_local] (nullable), p_alive/mean_x host [n_local] (nullable). */
int clv_forecast_resident(clv_sampler* h, double T_star, uint64_t seed, int64_t* x_star, double* mean_x_star,
                          double* p_alive, double* kernel_ms /* nullable: CUDA-event time of the kernel */);

/* Fused forecast: while clv_run / clv_run_resident keeps a draw, x* of that draw (bi:535-543) is simulated inside the sweep
 * kernel from lambda, tau, z still in registers and added to per-(chain, customer) sums -- the posterior-predictive means
 * need no pass over the stored draws at all (level1 may even be NULL).  Same Philox counters as clv_forecast_resident /
 * clv_forecast on the same draws (seed, global customer id, draw index chain * n_draws + draw), hence the same x*.
 * Enable before the run; the sums restart with every run. */
int clv_set_fused_forecast(clv_sampler* h, int enable, double T_star, uint64_t seed);
/* mean x* and P(alive) = mean z per customer over all kept draws of all chains of the last run (host [n_local], nullable). */
int clv_fused_forecast_result(clv_sampler* h, double* mean_x_star, double* p_alive, int64_t* n_draws_total);

/* ---- analysis reductions on the resident draws (SURVEY 8f "next" rows) ---------------------- */
/* Place host draws [chains][n_draws][n_local][4|5] (e.g. an unpickled "level_1" list, chain-major) into the handle's
 * resident buffer, so that the reductions below / clv_forecast_resident can serve draws that were not produced by this
 * handle.  Needs only clv_create (clv_forecast_resident additionally needs clv_set_data for T_cal). */
int clv_upload_draws(clv_sampler* h, const double* level1, int64_t n_draws);
#define CLV_SUMMARY_COLS 10
/* Per-customer posterior summaries over all resident draws of all chains (post_mean_lambdas/mus and compute_table4,
 * src/models/utils/analysis_bi_helpers.py:15-27, 75-110): out [n_local][CLV_SUMMARY_COLS] =
 * mean lambda, 2.5 % and 97.5 % of lambda | mean of min(mu, mu_cap), 2.5 % and 97.5 % of mu (raw) | mean z = P(alive) |
 * mean tau | mean mu (raw) | mean eta (0 for D=2).  Percentiles are np.percentile's (linear). */
int clv_posterior_summary(clv_sampler* h, double mu_cap, double* out);
/* Weekly tracking simulation (Figure 2; bivariate/analysis_abe.py:446-464): for each resident draw and each time
 * times[w], the sum over customers of Poisson(lambda_i) while birth_week_i < t <= birth_week_i + tau_i; returns the mean
 * over draws inc_mean[n_weeks] (the caller takes the cumulative sum, analysis_abe.py:460). */
int clv_weekly_tracking(clv_sampler* h, const double* birth_week, const double* times, int n_weeks, uint64_t seed,
                        double* inc_mean);

/* ---- synthetic customers, generate_pareto_abe (bi:95-187) ---------------------------------- */
/* Device-side generator of the CBS law (x, t_x, x_star | lambda, mu, tau) for n customers with design
 * matrix X = [1, U(-1,1)^(K-1)] (bi:118-122) and theta = exp(X beta + MVN(0, gamma)) (bi:133-136).
 * T_cal per customer ~ U(T_cal_lo, T_cal_hi).  Outputs are host arrays of length n (X: n x K). */
typedef struct {
  int32_t device;
  int32_t n_cov;
  int64_t n;
  int64_t gid_offset;
  uint64_t seed;
  double T_cal_lo, T_cal_hi, T_star;
} clv_generate_config;
/* X_given / T_cal_given != 0: the X (n x K, column 0 = 1) / T_cal arrays are inputs (bi:123-130, bi:140-142)
 * instead of being drawn. */
int clv_generate(const clv_generate_config* cfg, const double* beta /*K x 2*/, const double* gamma /*2 x 2*/,
                 int X_given, int T_cal_given, int32_t* x, double* t_x, double* T_cal, double* X, int32_t* x_star,
                 double* lambda_true, double* mu_true, double* tau_true);

/* ---- event log -> CBS (src/models/utils/elog2cbs2param.py:33-94) --------------------------------- */
/* Events in any order: cust ids, day numbers (days since any fixed epoch), sales (nullable: 1 per event, :45).  Same
 * (cust, day) events are one transaction with summed sales (:62).  T_cal_day / T_tot_day: last calibration / observation
 * day; unit_days: 7 for weeks.  Outputs (any nullable) have one row per customer with a calibration purchase, ascending
 * cust id; the caller sizes them for the number of distinct customers (<= n_events); *n_customers returns the row count.
 * first_sales (nullable): the sales of the customer's first event in INPUT order -- pandas' groupby("cust")["sales"].first()
 * of src/data_processing/2B_cdnow_elog2cbs_full.py:62-68.  Sorting, grouping, per-customer statistics and the dropping of
 * customers without a calibration purchase all run on the device; only the kept rows are copied back. */
int clv_elog2cbs(int device, int64_t n_events, const int64_t* cust, const int32_t* day, const double* sales,
                 int32_t T_cal_day, int32_t T_tot_day, double unit_days, int64_t* n_customers, int64_t* cust_out,
                 int32_t* x, double* t_x, double* litt, double* sales_out, double* sales_x, int32_t* first_day,
                 double* T_cal, double* T_star, int32_t* x_star, double* sales_star, double* first_sales);

/* ---- covariate standardisation (src/data_processing/2B_cdnow_elog2cbs_full.py:70-101) ----------------------------- */
/* out[i] = (scale * v[i] - mean) / sd with mean and sd (pandas: ddof = 1) of the scaled column: `first_sales_scaled`
 * (scale 1e-3, :68-80) and `age_scaled` (scale 1, :84-86).  mean_out / sd_out nullable. */
int clv_standardize(int device, int64_t n, const double* v, double scale, double* out, double* mean_out, double* sd_out);
/* out[i] = table[codes[i]] (NaN outside the table): `gender_binary` = gender.map({"M": 1, "F": 0}) (:97-100) on the
 * category codes. */
int clv_recode(int device, int64_t n, const int32_t* codes, const double* table, int n_table, double* out);

/* ---- test hook: the level-1 variates of MH step `step` (two Student-t3 proposals, accept uniform) that customers
 * 0..n-1 of chain 0 consume in sweep `sweep`, as the sweep kernel generates them in rng_mode fast / strict. */
int clv_debug_variates(int device, uint64_t seed, uint32_t sweep, int32_t step, int rng_mode, int64_t n, double* t3_l,
                       double* t3_m, double* u_acc);

/* ---- test hook: the customer-sharded path on ONE device.  `shards` are n_shards (2..16) handles on the same device holding
 * contiguous customer ranges of one problem (same model, chains, seed, n_global; gid_offset = start of the range; initialised
 * with the GLOBAL statistics; no communicator).  Advances all of them n_sweeps sweeps in lockstep; the per-sweep all-reduce
 * of the level-2 statistics runs the production peer-mailbox protocol with the ranks emulated as the blocks of one
 * cooperative launch (kernels of separate launches must not wait on each other on one GPU).  The shards' states must
 * then equal the unsharded handle's bit for bit -- the single-GPU proof of "results do not depend on the GPU count". */
int clv_debug_lockstep_advance(clv_sampler** shards, int n_shards, int64_t n_sweeps);

/* ---- test hook: the parallel host memcpy behind the staged transfers (no device involved): copies `bytes` from src to
 * dst on the library's copy threads, returns the number of threads that took part. */
int clv_debug_host_copy(void* dst, const void* src, int64_t bytes);
/* Test hook: the first-touch threads clv_run starts on the caller's level-1 array (one `lock or byte, 0` per page: takes the
 * write fault, leaves the data as it is), here on any buffer; returns when every page has been touched. */
int clv_debug_first_touch(void* buf, int64_t bytes, int threads);

/* ---- micro-benchmarks used by bench.py for the issue-rate roofline -------------------------- */
/* Measures, on `device`, sustained warp-instruction throughput of dependent-free loops of
 * FFMA, IMAD (32-bit), MUFU.EX2 and DFMA.  out[4] = giga thread-ops/s for each. */
int clv_measure_issue_peaks(int device, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* CLV_B200_H */
