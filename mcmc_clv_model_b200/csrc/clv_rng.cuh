// clv_rng.cuh — counter-based random numbers for the sampler.
//
// Philox4x32-10 keyed by the 64-bit seed, counter (customer_gid, sweep, slot, domain | global_chain << 4):
// a draw depends only on WHICH customer/chain/sweep it belongs to, never on the thread, block,
// shard or GPU that computes it.  The key being the same for every chain, the ten round keys are
// launch constants: the host passes them as kernel parameters and the hot loop reads them straight
// from the constant bank (no per-round key arithmetic).  oracle/philox_np.py restates this contract in NumPy; the STRICT
// transforms below (fp64) are what it reproduces, the FAST ones (fp32 through the SFU: MUFU.LG2 /
// MUFU.SIN / MUFU.COS / MUFU.RSQ) draw from the same laws.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace clv {

enum : uint32_t { DOM_SAMPLER = 0, DOM_LEVEL2 = 1, DOM_FORECAST = 2, DOM_GENERATOR = 3 };

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ inline PhiloxKey seed_key(uint64_t seed) { return PhiloxKey{(uint32_t)seed, (uint32_t)(seed >> 32)}; }
// fourth counter word: domain in the low 4 bits, global chain index above
__host__ __device__ inline uint32_t dom_word(uint32_t domain, uint32_t global_chain) { return domain | (global_chain << 4); }

// the ten round keys of a seed (k0 + r W0, k1 + r W1), for kernels that take them as parameters
struct PhiloxRoundKeys {
  uint32_t k[20];
};
__host__ __device__ inline PhiloxRoundKeys round_keys(uint64_t seed) {
  PhiloxRoundKeys rk;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    rk.k[2 * r] = k0;
    rk.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return rk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one IMAD.WIDE.U32 per product gives both halves
    uint32_t lo0, hi0, lo1, hi1;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(M0));
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(M1));
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// same function with the round keys precomputed (kernel parameters => constant-bank operands)
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxRoundKeys& rk) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t lo0, hi0, lo1, hi1;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(M0));
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(M1));
    const uint32_t n0 = hi1 ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// ---- uniforms ---------------------------------------------------------------------------------
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  uint64_t m = ((uint64_t)(a >> 5) << 26) + (uint64_t)(b >> 6);
  return ((double)m + 0.5) * 0x1.0p-53;
}
__device__ __forceinline__ double u32d(uint32_t a) { return ((double)a + 0.5) * 0x1.0p-32; }
// fp32 image of u32d (24 significant bits), only used to screen MH decisions
__device__ __forceinline__ float u32f(uint32_t a) { return fmaf((float)a, 0x1.0p-32f, 0x1.0p-33f); }
// (k + 0.5) 2^-24, k = top 24 bits: exact in fp32.  Converting the whole word with round-toward-zero keeps its top 24
// SIGNIFICANT bits -- exactly (a >> 8) << 8 for words >= 2^24, and up to 8 further low bits for the 2^-8 of the words
// below that (a difference < 2^-32 in the uniform) -- so the shift is folded into the conversion: one I2F.RZ + one FFMA.
__device__ __forceinline__ float u24f(uint32_t a) { return fmaf(__uint2float_rz(a), 0x1.0p-32f, 0x1.0p-25f); }

// the same uniform in fp64 (STRICT transforms)
__device__ __forceinline__ double u24d(uint32_t a) { return ((double)(a >> 8) + 0.5) * 0x1.0p-24; }

// ---- word layout of the Metropolis steps (sampler domain) ----------------------------------------------------------
// A step consumes ONE Philox block (slot 1 + step): words (x, y) = (V, angle) for the log-lambda proposal, (z, w) for
// log mu; each transform uses the TOP 24 bits of its word (STRICT: exactly a >> 8; FAST: the top 24 significant bits,
// which for a word below 2^24 -- one in 256 -- reach into the low byte and move the variate by < 2^-32).  The accept
// uniform is assembled from the LOW bytes of the four words.  Slot 0 is z/tau, slot 1+2S the eta normal (tri).
__device__ __forceinline__ uint32_t low_bytes(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  // three PRMT: byte 0 of w0..w3 -> bytes 0..3
  return __byte_perm(__byte_perm(w0, w1, 0x0040), __byte_perm(w2, w3, 0x0040), 0x5410);
}

// SFU primitives without the denormal / IEEE-rounding fix-up code the default intrinsics carry
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_ftz(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_ftz(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_ftz(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- Student t(3) without rejection, from TWO uniforms ---------------------------------------------------------------
// t = N0 / sqrt((N1^2 + C)/3) with (N0, N1) = R (cos th, sin th) a Box-Muller pair, R^2 = 2 E1, and C = 2 E3 an
// independent chi-square(2):  t / sqrt(3) = cos(th) / sqrt(sin^2(th) + E3 / E1).  The ratio of two independent
// Exp(1) variables has CDF r / (1 + r), i.e. E3 / E1 = V / (1 - V) with V uniform -- no logarithm is needed, and a
// variate costs one angle and one V:   t / sqrt(3) = cos(th) / sqrt(sin^2(th) + V / (1 - V)).
__device__ __forceinline__ double t3_strict(uint32_t rv, uint32_t rb) {
  const double v = u24d(rv), ang = 6.283185307179586476925286766559 * u24d(rb);
  double s, c;
  sincos(ang, &s, &c);
  return 1.7320508075688772 * c / sqrt(s * s + v / (1.0 - v));
}

// FAST variant (fp32, two SFU operations: cos, rsqrt), returns t / sqrt(3) (the caller folds sqrt(3) into the proposal
// scale).  With w = 1 - V:  cos(th) / sqrt(sin^2 + V / w) = cos(th) w rsqrt((sin^2 w + V) w),  sin^2 = 1 - cos^2.
// V is rounded toward zero so that it stays below 1 (w >= 2^-24, computed exactly); the map is odd in cos(th), so the
// proposal stays exactly symmetric whatever the rounding.
__device__ __forceinline__ float fma_rz(float a, float b, float c) { float y; asm("fma.rz.ftz.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float t3_fast(uint32_t rv, uint32_t rb) {
  const float v = fma_rz(__uint2float_rz(rv), 0x1.0p-32f, 0x1.0p-25f), w = 1.0f - v;
  // angle 2 pi (k + 0.5) 2^-24 straight from the integer
  const float ang = fmaf(__uint2float_rz(rb), 6.2831853071795865f * 0x1.0p-32f, 6.2831853071795865f * 0x1.0p-25f);
  const float c = cos_ftz(ang);
  const float s2 = fmaf(-c, c, 1.0f);
  return (c * w) * rsqrt_ftz(fmaf(s2, w, v) * w);
}

// two standard normals from four words (53-bit uniforms, fp64): cos / sin branch
__device__ __forceinline__ void normal_pair_u53(uint4 r, double* nc, double* ns) {
  double ua = u53(r.x, r.y), ub = u53(r.z, r.w);
  double rad = sqrt(-2.0 * log(ua));
  double s, c;
  sincos(6.283185307179586476925286766559 * ub, &s, &c);
  *nc = rad * c;
  *ns = rad * s;
}

// ---- level-2 domain (one thread per chain; always fp64) -----------------------------------------
__device__ inline double level2_normal(PhiloxKey key, uint32_t c3, uint32_t sweep, uint32_t idx) {
  double c, s;
  normal_pair_u53(philox4x32_10(idx, sweep, 0u, c3, key), &c, &s);
  return c;
}

// chi2(df) = 2 Gamma(df/2, 1), Marsaglia-Tsang (df >= 2)
__device__ inline double level2_chi2(PhiloxKey key, uint32_t c3, uint32_t sweep, uint32_t idx, double df) {
  double a = 0.5 * df;
  double d = a - 1.0 / 3.0;
  double c = 1.0 / sqrt(9.0 * d);
  for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
    double x, unused;
    normal_pair_u53(philox4x32_10(idx, sweep, 2u * attempt, c3, key), &x, &unused);
    uint4 rb = philox4x32_10(idx, sweep, 2u * attempt + 1u, c3, key);
    double u = u53(rb.x, rb.y);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return 2.0 * d * v;
  }
  return 2.0 * d;  // unreachable in practice (acceptance > 0.95 per attempt)
}

}  // namespace clv
