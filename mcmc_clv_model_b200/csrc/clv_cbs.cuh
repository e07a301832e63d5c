// clv_cbs.cuh — event log -> customer-by-sufficient-statistic table (SURVEY 8f row f-3;
// reference: src/models/utils/elog2cbs2param.py:33-94, a pandas groupby pipeline).
// Device pipeline: stable radix sort by (cust, day) [CUB, two passes], head flags, exclusive scan -> customer index,
// then ONE THREAD PER CUSTOMER walks its (few) sorted events sequentially: same-day events merge (sales summed,
// elog2cbs2param.py:62), calibration / hold-out statistics accumulate in event order (deterministic, no atomics).
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
  int* keep;            // 1 if the customer has a calibration purchase (others are dropped, as the groupby does)
};

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

__global__ void k_cbs_customers(const long long* cust, const int* day, const double* sales, const long long* starts,
                                long long n_events, long long n_cust, int T_cal_day, int T_tot_day, double unit, CbsOut o) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cust; c += (long long)gridDim.x * blockDim.x) {
    const long long e0 = starts[c], e1 = (c + 1 < n_cust) ? starts[c + 1] : n_events;
    const int first = day[e0];                      // sorted: the customer's first purchase day
    int n_cal = 0, n_val = 0, last_cal = first, prev_day = first;
    double s_cal = 0.0, s_first = 0.0, s_val = 0.0, litt = 0.0;
    long long e = e0;
    while (e < e1) {
      const int d = day[e];
      double s = 0.0;
      for (; e < e1 && day[e] == d; ++e) s += sales ? sales[e] : 1.0;     // same (cust, date): one transaction, sales summed
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
  }
}

}  // namespace clv
