// clv_cbs.cuh — event log -> customer-by-sufficient-statistic table (SURVEY 8f row f-3;
// reference: src/models/utils/elog2cbs2param.py:33-94, a pandas groupby pipeline).
// Device pipeline: identity permutation (k_iota), stable radix sort by (cust, day) [CUB, two passes on signed keys],
// head flags, exclusive scan -> customer index, then ONE THREAD PER CUSTOMER walks its (few) sorted events
// sequentially: same-day events merge (sales summed, elog2cbs2param.py:62), calibration / hold-out statistics
// accumulate in event order (deterministic, no atomics); customers without a calibration purchase are dropped by a
// scan of the keep flags + scatter (k_cbs_compact), so only the kept rows travel to the host.
// Covariate standardisation of src/data_processing/2B_cdnow_elog2cbs_full.py:62-105: the first purchase amount (first
// row of the customer in INPUT order, as pandas' groupby.first) comes out of the same walk; z-scores (pandas std,
// ddof = 1) by a deterministic two-pass reduction (k_col_sum / k_zscore); category recoding by a small table.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace clv {

struct CbsOut {
  long long* cust;      // customer id
  int* x;               // repeat transactions in calibration (distinct days - 1)        :76
  double* t_x;          // recency: last calibration purchase - first, in `unit` days      :77
  double* litt;         // sum of log inter-transaction times (calibration)                :78
  double* sales;        // calibration sales                                               :79
  double* sales_x;      // calibration sales excluding the first day                       :80
  int* first_day;       // day number of the first purchase                                :81
  double* T_cal;        // (T_cal - first) / unit                                          :84
  double* T_star;       // (T_tot - first) / unit - T_cal                                  :89
  int* x_star;          // distinct purchase days in (T_cal, T_tot]                        :91
  double* sales_star;   // hold-out sales
  double* first_sales;  // sales of the customer's first event in INPUT order (2B_cdnow_elog2cbs_full.py:62-68, groupby.first)
  int* keep;            // 1 if the customer has a calibration purchase (others are dropped, as the groupby does)
};

__global__ void k_iota(unsigned* p, long long n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) p[e] = (unsigned)e;
}

template <typename T>
__global__ void k_gather(T* dst, const T* src, const unsigned* perm, long long n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dst[e] = src[perm[e]];
}

__global__ void k_cbs_heads(const long long* cust, long long n, int* head) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    head[e] = (e == 0 || cust[e] != cust[e - 1]) ? 1 : 0;
}

// starts[c] = first event of customer c (head positions scattered by the exclusive scan of the head flags)
__global__ void k_cbs_starts(const int* head, const int* idx, long long n, long long* starts) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    if (head[e]) starts[idx[e]] = e;
}

// perm[e] = input position of sorted event e; sales_in = the sales column in input order (nullable)
__global__ void k_cbs_customers(const long long* cust, const int* day, const double* sales, const unsigned* perm,
                                const double* sales_in, const long long* starts,
                                long long n_events, long long n_cust, int T_cal_day, int T_tot_day, double unit, CbsOut o) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cust; c += (long long)gridDim.x * blockDim.x) {
    const long long e0 = starts[c], e1 = (c + 1 < n_cust) ? starts[c + 1] : n_events;
    const int first = day[e0];                      // sorted: the customer's first purchase day
    int n_cal = 0, n_val = 0, last_cal = first, prev_day = first;
    double s_cal = 0.0, s_first = 0.0, s_val = 0.0, litt = 0.0;
    long long e = e0;
    unsigned first_in = 0xffffffffu;                // smallest input position among the customer's events
    while (e < e1) {
      const int d = day[e];
      double s = 0.0;
      for (; e < e1 && day[e] == d; ++e) {          // same (cust, date): one transaction, sales summed
        s += sales ? sales[e] : 1.0;
        first_in = min(first_in, perm[e]);
      }
      if (d <= T_cal_day) {
        if (n_cal == 0) s_first = s;
        else litt += log((double)(d - prev_day) / unit);                   // itt > 0 always for distinct days
        ++n_cal;
        s_cal += s;
        last_cal = d;
      } else if (d <= T_tot_day) {
        ++n_val;
        s_val += s;
      }
      prev_day = d;
    }
    o.cust[c] = cust[e0];
    o.keep[c] = n_cal > 0;
    o.x[c] = n_cal > 0 ? n_cal - 1 : 0;
    o.t_x[c] = (double)(last_cal - first) / unit;
    o.litt[c] = litt;
    o.sales[c] = s_cal;
    o.sales_x[c] = s_cal - s_first;
    o.first_day[c] = first;
    o.T_cal[c] = (double)(T_cal_day - first) / unit;
    o.T_star[c] = (double)(T_tot_day - first) / unit - o.T_cal[c];
    o.x_star[c] = n_val;
    o.sales_star[c] = s_val;
    o.first_sales[c] = sales_in ? sales_in[first_in] : 1.0;
  }
}

// Rows with keep != 0 move to position pos[c] (exclusive scan of keep) of the compact table.
__global__ void k_cbs_compact(CbsOut in, const int* pos, long long n_cust, CbsOut out) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cust; c += (long long)gridDim.x * blockDim.x) {
    if (!in.keep[c]) continue;
    const int m = pos[c];
    out.cust[m] = in.cust[c]; out.x[m] = in.x[c]; out.t_x[m] = in.t_x[c]; out.litt[m] = in.litt[c];
    out.sales[m] = in.sales[c]; out.sales_x[m] = in.sales_x[c]; out.first_day[m] = in.first_day[c];
    out.T_cal[m] = in.T_cal[c]; out.T_star[m] = in.T_star[c]; out.x_star[m] = in.x_star[c];
    out.sales_star[m] = in.sales_star[c]; out.first_sales[m] = in.first_sales[c];
  }
}

// ---- column standardisation ---------------------------------------------------------------------------------------
// Deterministic sum of f(v[i]) over a column: per-thread partials in a fixed grid-stride order, block tree, one partial
// per block; a second launch with one block folds the partials.  pass 0: v ; pass 1: (v - mean)^2.
constexpr int COLSUM_BLOCKS = 256, COLSUM_THREADS = 256;
__global__ void __launch_bounds__(COLSUM_THREADS) k_col_sum(const double* v, long long n, int pass, const double* mean, double* partial) {
  __shared__ double sh[COLSUM_THREADS];
  const double m = pass ? *mean : 0.0;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = v[i] - m;
    acc += pass ? d * d : d;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = COLSUM_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// out[0] = (sum of partials) / denom (pass 0: the mean; pass 1 with sqrt: the standard deviation)
__global__ void __launch_bounds__(COLSUM_THREADS) k_col_fold(const double* partial, int nparts, double denom, int take_sqrt, double* out) {
  __shared__ double sh[COLSUM_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += COLSUM_THREADS) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = COLSUM_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { const double r = sh[0] / denom; *out = take_sqrt ? sqrt(r) : r; }
}
__global__ void k_zscore(const double* v, long long n, const double* mean, const double* sd, double scale, double* out) {
  const double m = *mean, s = *sd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (v[i] * scale - m) / s;
}
// category codes -> values through a small table (e.g. gender {F, M} -> {0, 1}); codes outside the table give NaN (pandas map)
__global__ void k_recode(const int* code, long long n, const double* lut, int nlut, double* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = code[i];
    out[i] = (c >= 0 && c < nlut) ? lut[c] : __longlong_as_double(0x7ff8000000000000ll);
  }
}

}  // namespace clv
