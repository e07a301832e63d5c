"""Event log -> customer-by-sufficient-statistic (CBS) table on the device: host mirror of the reference's
`elog2cbs` (src/models/utils/elog2cbs2param.py:33-94) -- same arguments, same returned columns."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_UNIT_DAYS = {"D": 1.0, "day": 1.0, "days": 1.0, "W": 7.0, "week": 7.0, "weeks": 7.0}


def elog2cbs(elog, units="week", T_cal=None, T_tot=None, device=0, with_first_sales=False):
    """Convert an event log (columns cust, date[, sales]) into one row per customer: cust, x, t_x, litt, sales, sales_x,
    first, T_cal[, T_star, x_star, sales_star].  Dates are used at day resolution; `cust` must be integer-valued.
    with_first_sales=True appends `first_sales`, the sales of each customer's first row in the log (the
    groupby("cust")["sales"].first() of src/data_processing/2B_cdnow_elog2cbs_full.py:62-68), from the same device pass."""
    import pandas as pd
    if not isinstance(elog, pd.DataFrame):
        raise ValueError("elog must be a pandas DataFrame")
    if "cust" not in elog.columns or "date" not in elog.columns:
        raise ValueError("elog must contain 'cust' and 'date' columns")
    if elog.empty:
        return pd.DataFrame(columns=["cust", "x", "t.x", "litt", "first", "T.cal"])
    if units not in _UNIT_DAYS:
        raise ValueError(f"units must be one of {sorted(_UNIT_DAYS)}")
    if "sales" in elog.columns and not pd.api.types.is_numeric_dtype(elog["sales"]):
        raise ValueError("'sales' column must be numeric")
    dates = pd.to_datetime(elog["date"])
    epoch = pd.Timestamp("1970-01-01")
    day = ((dates - epoch) // pd.Timedelta(days=1)).to_numpy().astype(np.int32)
    t_cal = dates.max() if T_cal is None else pd.to_datetime(T_cal)
    t_tot = dates.max() if T_tot is None else pd.to_datetime(T_tot)
    has_holdout = t_cal < t_tot
    cal_day, tot_day = int((t_cal - epoch) // pd.Timedelta(days=1)), int((t_tot - epoch) // pd.Timedelta(days=1))
    cust = np.ascontiguousarray(elog["cust"].to_numpy(), dtype=np.int64)
    sales = np.ascontiguousarray(elog["sales"].to_numpy(), dtype=np.float64) if "sales" in elog.columns else None
    n = cust.size
    nc = np.unique(cust).size
    out = dict(cust=np.empty(nc, np.int64), x=np.empty(nc, np.int32), t_x=np.empty(nc), litt=np.empty(nc), sales=np.empty(nc),
               sales_x=np.empty(nc), first=np.empty(nc, np.int32), T_cal=np.empty(nc), T_star=np.empty(nc),
               x_star=np.empty(nc, np.int32), sales_star=np.empty(nc), first_sales=np.empty(nc))
    m = C.c_int64()
    i32 = lambda a: a.ctypes.data_as(L.c_int32_p)  # noqa: E731
    L.check(L.load().clv_elog2cbs(int(device), n, cust.ctypes.data_as(L.c_int64_p), i32(np.ascontiguousarray(day)), L.dptr(sales),
                                  cal_day, tot_day, _UNIT_DAYS[units], C.byref(m), out["cust"].ctypes.data_as(L.c_int64_p),
                                  i32(out["x"]), L.dptr(out["t_x"]), L.dptr(out["litt"]), L.dptr(out["sales"]),
                                  L.dptr(out["sales_x"]), i32(out["first"]), L.dptr(out["T_cal"]), L.dptr(out["T_star"]),
                                  i32(out["x_star"]), L.dptr(out["sales_star"]), L.dptr(out["first_sales"])))
    k = m.value
    df = pd.DataFrame({"cust": out["cust"][:k], "x": out["x"][:k].astype(np.int64), "t_x": out["t_x"][:k], "litt": out["litt"][:k],
                       "sales": out["sales"][:k], "sales_x": out["sales_x"][:k],
                       "first": epoch + pd.to_timedelta(out["first"][:k].astype(np.int64), unit="D"), "T_cal": out["T_cal"][:k]})
    if has_holdout:
        df["T_star"] = out["T_star"][:k]
        df["x_star"] = out["x_star"][:k].astype(float)
        df["sales_star"] = out["sales_star"][:k]
    if with_first_sales:
        df["first_sales"] = out["first_sales"][:k]
    return df


def standardize(values, scale=1.0, device=0):
    """(scale * v - mean) / std with pandas' std (ddof = 1) of the scaled column, on the device
    (2B_cdnow_elog2cbs_full.py:70-86).  Returns (z, mean, std)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.empty_like(v)
    m, s = C.c_double(), C.c_double()
    L.check(L.load().clv_standardize(int(device), v.size, L.dptr(v), float(scale), L.dptr(out), C.byref(m), C.byref(s)))
    return out, m.value, s.value


def add_covariates(cbs, elog, customers, device=0):
    """The covariate columns the reference's data-processing script adds to the full-CDNOW CBS
    (src/data_processing/2B_cdnow_elog2cbs_full.py:44-105): `customers` (columns cust, age, gender) is left-merged on
    cust; first_sales_scaled = z-score of 1e-3 x the customer's first purchase amount; age_scaled = z-score of age;
    gender_binary = {"M": 1, "F": 0}.  The arithmetic (first purchase per customer, means, standard deviations, z-scores,
    recoding) runs on the device; pandas only aligns the rows."""
    import pandas as pd
    out = cbs.merge(customers, on="cust", how="left")
    if "first_sales" in out.columns:
        fs = out["first_sales"].to_numpy(float)
    else:
        first = elog2cbs(elog, units="D", device=device, with_first_sales=True)[["cust", "first_sales"]]
        fs = out[["cust"]].merge(first, on="cust", how="left")["first_sales"].to_numpy(float)
    out["first_sales_scaled"] = standardize(fs, 1e-3, device)[0]
    out["age_scaled"] = standardize(out["age"].to_numpy(float), 1.0, device)[0]
    codes = pd.Categorical(out["gender"], categories=["F", "M"]).codes.astype(np.int32)        # -1 for anything else
    g = np.empty(len(out))
    L.check(L.load().clv_recode(int(device), codes.size, codes.ctypes.data_as(L.c_int32_p), L.dptr(np.array([0.0, 1.0])), 2, L.dptr(g)))
    out["gender_binary"] = g
    return out.drop(columns=[c for c in ("gender", "zone", "state", "age_category", "first_sales") if c in out.columns])
