// znll_device.cuh -- hand-written sm_100a device code of the fused unbinned-NLL kernel.
//
// One kernel instantiation per model tree: the tree is a C++ type (e.g. Sum<CBall<0>,Expo<0>>),
// every node contributes
//   * a forward evaluation per event (normalised value P or ln P),
//   * a backward sweep that adds  dL/d(slot)  into per-thread register accumulators
//     (L = sum_i w_i*log(P_i + 1e-307); reverse mode over the tiny tree, analytic leaf derivatives),
// and the kernel wraps this in a persistent grid-stride loop over double2-vectorised SoA columns,
// followed by a warp-shuffle / shared-memory block reduction and a deterministic fixed-order
// cross-block reduce executed by the last block to finish (ticket counter).
//
// The same header is compiled ahead of time by nvcc (znll_catalog.cu) and, for tree shapes that
// are not in the catalogue, at run time by NVRTC (znll_host.cu) -- so it only depends on CUDA
// built-ins, no standard headers.
//
// Per-call constants (normalisation integrals etc.) are computed in fp64 on the host
// (znll_host.cu: fill_consts) and arrive in the kernel parameter block (constant bank).
// Layout contract of consts / accumulators: pre-order over the tree, a node's own entries first.
//
// Reference formulas (zfit 0.28.0):  Gauss dist_tfp.py:108-120;  Exponential basic.py:111-126;
// CrystalBall physics.py:31-45;  Chebyshev polynomials.py:32-43,164-175,338-347;
// SumPDF functor.py:197-206;  ProductPDF functor.py:408-433;  NLL loss.py:73-142.
#pragma once

#define ZNLL_MAXCOL 8
#define ZNLL_BLOCK 256
#ifndef ZNLL_MIN_BLOCKS
#define ZNLL_MIN_BLOCKS 3  // resident CTAs per SM the register allocation is tuned for
#endif
#define ZNLL_LOG_EPS 1e-307

namespace znll {

// ----------------------------------------------------------------------------------------------
// fp64 elementary functions tuned for this kernel (the FP64 pipe is the roofline):
//   * polynomial coefficients live in the constant bank, so every DFMA takes them as an operand
//     (libdevice materialises each one with two UMOVs -> more issue slots than DP work);
//   * log uses a 128-entry (1/c, log c) table in shared memory and a degree-7 log1p polynomial:
//     13 DP instructions instead of ~35 (no division); exp: 15 DP instructions;
//   * reciprocal: MUFU.RCP64H seed + two Newton steps (4 DFMA).
// Accuracy (tests/test_math_gpu.py): exp, log <= 1 ulp-level relative error (measured ~2e-16),
// far inside the 1e-12 NLL tolerance; arguments outside the fast range (|x| >= 700 for exp;
// non-positive, subnormal, inf or NaN for log / rcp) take the libdevice path.
// ----------------------------------------------------------------------------------------------
static __constant__ double kExpC[10] = {
    0.5000000000000001,     0.16666666666666669,    0.04166666666662413,   0.008333333333330062,
    0.0013888888917213717,  0.00019841269863053618, 2.4801521295954376e-05, 2.7557268459997064e-06,
    2.7620088445409746e-07, 2.510038549551032e-08};
static __constant__ double kLogC[6] = {-0.5,
                                       0.33333333333333337,
                                       -0.24999999998288297,
                                       0.19999999998478485,
                                       -0.16666959217649738,
                                       0.14285974331115564};
static __constant__ double kMathC[8] = {
    1.4426950408889634,        // 0 log2(e)
    6755399441055744.0,        // 1 1.5 * 2^52
    -6.93147180369123816490e-01,  // 2 -ln2_hi
    -1.90821492927058770002e-10,  // 3 -ln2_lo
    0x1.62e42fefa3800p-1,      // 4 ln2_hi (42 significant bits: k*ln2_hi is exact)
    0x1.ef35793c76730p-45,     // 5 ln2_lo
    4503601774854144.0,        // 6 2^52 + 2^31
    0.0};

struct Ctx {
  const double2* logtab;  // shared memory: (1/c_i, log c_i) for 128 intervals of [0.6875, 1.375)
};

__device__ __forceinline__ double fast_exp(double x) {
  const int hx = __double2hiint(x) & 0x7fffffff;
  if (hx >= 0x4085E000) return exp(x);  // |x| >= 700, inf, NaN: rare
  const double t = fma(x, kMathC[0], kMathC[1]);
  const int n = __double2loint(t);
  const double nd = t - kMathC[1];
  double r = fma(nd, kMathC[2], x);
  r = fma(nd, kMathC[3], r);
  double p = kExpC[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

__device__ __forceinline__ double fast_rcp(double x) {
  const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
  if (ex - 1u >= 0x7fdu) return 1.0 / x;  // zero, subnormal, inf, NaN
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ double fast_log(double x, const Ctx& cx) {
  const int hi = __double2hiint(x);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(x);  // x <= 0, subnormal, inf, NaN
  const int tmp = hi - 0x3fe60000;
  const int i = (tmp >> 13) & 127;
  const int k = tmp >> 20;
  const double z = __hiloint2double(hi - (tmp & 0xfff00000), __double2loint(x));
  const double2 tc = cx.logtab[i];
  const double r = fma(z, tc.x, -1.0);
  const double kd = __hiloint2double(0x43300000, k ^ 0x80000000) - kMathC[6];
  double s = kLogC[5];
#pragma unroll
  for (int j = 4; j >= 0; --j) s = fma(s, r, kLogC[j]);
  const double w = fma(kd, kMathC[4], tc.y);
  const double t2 = fma(r * r, s, kd * kMathC[5]);
  return w + (r + t2);
}

// one entry per thread < 128, then __syncthreads() by the caller
__device__ __forceinline__ void init_logtab(double2* tab) {
  if (threadIdx.x < 128) {
    const int ha = 0x3fe60000 + ((int)threadIdx.x << 13);
    const double a = __hiloint2double(ha, 0), b = __hiloint2double(ha + (1 << 13), 0);
    const double invc = 1.0 / (0.5 * (a + b));
    tab[threadIdx.x] = make_double2(invc, -log(invc));
  }
}

// ----------------------------------------------------------------------------------------------
// leaves.  c = const block, x = this event's column values, acc = per-thread accumulators.
// CO / AO are compile-time offsets into c / acc, so every access is a fixed register or a
// constant-bank operand.
// ----------------------------------------------------------------------------------------------

// Gauss: consts [1/sigma, -mu/sigma, -(0.5*ln(2pi)+ln sigma) - ln I];  slots [mu, sigma]
template <int COL>
struct Gauss {
  static constexpr int NS = 2, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  double t, P;

  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool&) {
    t = fma(x[COL], c[CO + 0], c[CO + 1]);  // x/sigma - mu/sigma (tfp Normal._log_prob)
    return fma(-0.5 * t, t, c[CO + 2]);
  }
  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    bool neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg));
    return P;
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__ c, double* acc) {
    const double gi = g * c[CO + 0];
    acc[AO + 0] = fma(gi, t, acc[AO + 0]);                  // d/dmu    =  t/sigma
    acc[AO + 1] = fma(gi, fma(t, t, -1.0), acc[AO + 1]);    // d/dsigma = (t^2-1)/sigma
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    bwd_log<CO, AO>(a * P, c, acc);
  }
};

// Exponential: consts [lambda, shift, -ln I];  slots [lambda]
template <int COL>
struct Expo {
  static constexpr int NS = 1, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  double xs, P;

  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool&) {
    xs = x[COL] - c[CO + 1];
    return fma(c[CO + 0], xs, c[CO + 2]);
  }
  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    bool neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg));
    return P;
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__, double* acc) {
    acc[AO + 0] = fma(g, xs, acc[AO + 0]);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    bwd_log<CO, AO>(a * P, c, acc);
  }
};

// CrystalBall: slots [mu, sigma, alpha, n]
// consts [0 mu, 1 sign(alpha)/sigma, 2 1/sigma, 3 |alpha|, 4 n, 5 B=n/|alpha|-|alpha|,
//         6 ln A - ln I, 7 -ln I, 8 ln(n/|alpha|)+1, 9 1/|alpha|, 10 -n/|alpha|-|alpha|,
//         11 n/alpha^2+1, 12 sign(alpha)]
template <int COL>
struct CBall {
  static constexpr int NS = 4, NC = 13, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  double t, dt, lu, P;  // dt = d ln v / d t ; lu = ln(B-t) on the tail
  bool tail;

  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool&) {
    t = (x[COL] - c[CO + 0]) * c[CO + 1];
    tail = t < -c[CO + 3];
    double lp;
    if (tail) {  // A*(B-t)^-n  ==  exp(ln A - n*ln(B-t))
      const double u = c[CO + 5] - t;
      lu = fast_log(u, cx);
      dt = c[CO + 4] * fast_rcp(u);
      lp = fma(-c[CO + 4], lu, c[CO + 6]);
    } else {  // exp(-t^2/2)
      lu = 0.0;
      dt = -t;
      lp = fma(-0.5 * t, t, c[CO + 7]);
    }
    return lp;
  }
  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    bool neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg));
    return P;
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__ c, double* acc) {
    const double gd = g * dt;
    acc[AO + 0] = fma(-gd, c[CO + 1], acc[AO + 0]);          // dt/dmu    = -sign/sigma
    acc[AO + 1] = fma(-gd * t, c[CO + 2], acc[AO + 1]);      // dt/dsigma = -t/sigma
    if (tail) {
      // d ln v/d|alpha| = -n/a - a + (n/u)(n/a^2+1) ; d ln v/dn = ln(n/a)+1 - ln u - (n/u)/a
      acc[AO + 2] = fma(g * c[CO + 12], fma(dt, c[CO + 11], c[CO + 10]), acc[AO + 2]);
      acc[AO + 3] = fma(g, (c[CO + 8] - lu) - dt * c[CO + 9], acc[AO + 3]);
    }
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    bwd_log<CO, AO>(a * P, c, acc);
  }
};

// Chebyshev of degree DEG: slots [c_0 .. c_DEG]
// consts [0 2/(hi-lo), 1 -(lo+hi)/(hi-lo), 2 1/I, 3.. c_k/I]
template <int COL, int DEG>
struct Cheb {
  static constexpr int NS = DEG + 1, NC = 3 + DEG + 1, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = false;
  double T[DEG + 1];
  double P;

  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    const double z = fma(x[COL], c[CO + 0], c[CO + 1]);
    T[0] = 1.0;
    double p = c[CO + 3];
    if constexpr (DEG >= 1) {
      T[1] = z;
      p = fma(c[CO + 4], z, p);
    }
    const double z2 = z + z;
#pragma unroll
    for (int k = 2; k <= DEG; ++k) {
      T[k] = fma(z2, T[k - 1], -T[k - 2]);
      p = fma(c[CO + 3 + k], T[k], p);
    }
    P = p;
    return p;
  }
  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool& neg) {
    const double p = fwd_val<CO>(cx, c, x);
    neg ^= (p < 0.0);
    return fast_log(fabs(p), cx);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    const double ai = a * c[CO + 2];
    acc[AO + 0] += ai;
#pragma unroll
    for (int k = 1; k <= DEG; ++k) acc[AO + k] = fma(ai, T[k], acc[AO + k]);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__ c, double* acc) {
    bwd_val<CO, AO>(g * fast_rcp(P), c, acc);
  }
};

// ----------------------------------------------------------------------------------------------
// functors
// ----------------------------------------------------------------------------------------------
template <class... Cs>
struct Kids;
template <>
struct Kids<> {
  static constexpr int NS = 0, NC = 0, COLMASK = 0, N = 0;
};
template <class H, class... T>
struct Kids<H, T...> {
  H h;
  Kids<T...> t;
  static constexpr int NS = H::NS + Kids<T...>::NS;
  static constexpr int NC = H::NC + Kids<T...>::NC;
  static constexpr int COLMASK = H::COLMASK | Kids<T...>::COLMASK;
  static constexpr int N = 1 + sizeof...(T);
};

// SumPDF: consts [f_0..f_{K-1}]; slots [f_0..f_{K-1}]; then the children
template <class... Cs>
struct Sum {
  static constexpr int K = sizeof...(Cs);
  static constexpr int NS = K + Kids<Cs...>::NS, NC = K + Kids<Cs...>::NC;
  static constexpr int COLMASK = Kids<Cs...>::COLMASK;
  static constexpr bool LOGNAT = false;
  Kids<Cs...> kids;
  double p[K];
  double P;

  template <int CO, int I, int CK, class KD>
  __device__ __forceinline__ void fwd_rec(KD& kd, const Ctx& cx, const double* __restrict__ c, const double* x, double& tot) {
    if constexpr (KD::N > 0) {
      p[I] = kd.h.template fwd_val<CK>(cx, c, x);
      tot = fma(c[CO + I], p[I], tot);
      fwd_rec<CO, I + 1, CK + decltype(kd.h)::NC>(kd.t, cx, c, x, tot);
    }
  }
  template <int CO, int AO, int I, int CK, int AK, class KD>
  __device__ __forceinline__ void bwd_rec(KD& kd, double a, const double* __restrict__ c, double* acc) {
    if constexpr (KD::N > 0) {
      acc[AO + I] = fma(a, p[I], acc[AO + I]);
      kd.h.template bwd_val<CK, AK>(a * c[CO + I], c, acc);
      bwd_rec<CO, AO, I + 1, CK + decltype(kd.h)::NC, AK + decltype(kd.h)::NS>(kd.t, a, c, acc);
    }
  }
  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    double tot = 0.0;
    fwd_rec<CO, 0, CO + K>(kids, cx, c, x, tot);
    P = tot;
    return tot;
  }
  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool& neg) {
    const double v = fwd_val<CO>(cx, c, x);
    neg ^= (v < 0.0);
    return fast_log(fabs(v), cx);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    bwd_rec<CO, AO, 0, CO + K, AO + K>(kids, a, c, acc);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__ c, double* acc) {
    bwd_val<CO, AO>(g * fast_rcp(P), c, acc);
  }
};

// ProductPDF over disjoint observables: no consts, no slots of its own
template <class... Cs>
struct Prod {
  static constexpr int NS = Kids<Cs...>::NS, NC = Kids<Cs...>::NC;
  static constexpr int COLMASK = Kids<Cs...>::COLMASK;
  static constexpr bool LOGNAT = true;
  Kids<Cs...> kids;
  double P;

  template <int CK, class KD>
  __device__ __forceinline__ void fwd_rec(KD& kd, const Ctx& cx, const double* __restrict__ c, const double* x,
                                          double& tot, bool& neg) {
    if constexpr (KD::N > 0) {
      tot += kd.h.template fwd_log<CK>(cx, c, x, neg);
      fwd_rec<CK + decltype(kd.h)::NC>(kd.t, cx, c, x, tot, neg);
    }
  }
  template <int CK, int AK, class KD>
  __device__ __forceinline__ void bwd_rec(KD& kd, double g, const double* __restrict__ c, double* acc) {
    if constexpr (KD::N > 0) {
      kd.h.template bwd_log<CK, AK>(g, c, acc);
      bwd_rec<CK + decltype(kd.h)::NC, AK + decltype(kd.h)::NS>(kd.t, g, c, acc);
    }
  }
  template <int CO>
  __device__ __forceinline__ double fwd_log(const Ctx& cx, const double* __restrict__ c, const double* x, bool& neg) {
    double tot = 0.0;
    fwd_rec<CO>(kids, cx, c, x, tot, neg);
    return tot;
  }
  template <int CO>
  __device__ __forceinline__ double fwd_val(const Ctx& cx, const double* __restrict__ c, const double* x) {
    bool neg = false;
    const double lp = fwd_log<CO>(cx, c, x, neg);
    P = neg ? -fast_exp(lp) : fast_exp(lp);
    return P;
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_log(double g, const double* __restrict__ c, double* acc) {
    bwd_rec<CO, AO>(kids, g, c, acc);
  }
  template <int CO, int AO>
  __device__ __forceinline__ void bwd_val(double a, const double* __restrict__ c, double* acc) {
    bwd_log<CO, AO>(a * P, c, acc);
  }
};

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
template <int NC>
struct KArgs {
  const double* cols[ZNLL_MAXCOL];
  const double* w;
  long long n;
  double log_offset;
  double* partials;        // [gridDim.x][NACC]
  unsigned int* ticket;    // zero before launch; reset by the last block
  double* out;             // [NACC]
  double c[NC > 0 ? NC : 1];
};

// accumulators: [0] running sum of w*l - offset, [1] its Kahan compensation, [2..] slots.
// Returns this event's term  w*l - offset  (the caller adds two of them, then one Kahan step).
template <class M, bool WEIGHTED, bool GRAD>
__device__ __forceinline__ double event(const Ctx& cx, const double* __restrict__ c, const double* x, double w,
                                        double off, double* acc) {
  M m;
  double l, g;
  if constexpr (M::LOGNAT) {
    bool neg = false;
    const double lp = m.template fwd_log<0>(cx, c, x, neg);
    l = lp;
    g = 1.0;
    if (lp < -690.0 || neg) {  // p + 1e-307 matters only here (loss.py:117)
      const double p = neg ? -exp(lp) : exp(lp);
      const double pp = p + ZNLL_LOG_EPS;
      l = log(pp);
      g = p / pp;
    }
    if constexpr (WEIGHTED) g *= w;
    if constexpr (GRAD) m.template bwd_log<0, 2>(g, c, acc);
  } else {
    const double P = m.template fwd_val<0>(cx, c, x);
    const double pp = P + ZNLL_LOG_EPS;
    l = fast_log(pp, cx);
    if constexpr (GRAD) {
      g = fast_rcp(pp);
      if constexpr (WEIGHTED) g *= w;
      m.template bwd_val<0, 2>(g, c, acc);
    }
  }
  if constexpr (WEIGHTED)
    return fma(w, l, -off);  // weights multiply first, then the unweighted offset (loss.py:134-137)
  else
    return l - off;
}

__device__ __forceinline__ void kahan_add(double* acc, double term) {
  const double y = term - acc[1];
  const double s = acc[0] + y;
  acc[1] = (s - acc[0]) - y;
  acc[0] = s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// error-free merge of two compensated sums (s - c is the value)
__device__ __forceinline__ void kahan_merge(double& s, double& c, double s2, double c2) {
  const double t = s + s2;
  const double bp = t - s;
  const double e = (s - (t - bp)) + (s2 - bp);  // TwoSum error of s + s2
  s = t;
  c = (c + c2) - e;
}

template <int NACC>
__device__ __forceinline__ void block_reduce_and_finish(double* acc, double* partials, unsigned int* ticket,
                                                        double* out) {
  __shared__ double sm[ZNLL_BLOCK / 32][NACC];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // value pair: error-free warp merge
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double s2 = __shfl_down_sync(0xffffffffu, acc[0], o);
    const double c2 = __shfl_down_sync(0xffffffffu, acc[1], o);
    kahan_merge(acc[0], acc[1], s2, c2);
  }
#pragma unroll
  for (int j = 2; j < NACC; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) sm[warp][j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = sm[0][0], cc = sm[0][1];
    for (int wv = 1; wv < ZNLL_BLOCK / 32; ++wv) kahan_merge(s, cc, sm[wv][0], sm[wv][1]);
    partials[(size_t)blockIdx.x * NACC + 0] = s;
    partials[(size_t)blockIdx.x * NACC + 1] = cc;
  } else if (threadIdx.x >= 2 && threadIdx.x < NACC) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < ZNLL_BLOCK / 32; ++wv) s += sm[wv][threadIdx.x];
    partials[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // deterministic finish: the order of the sum over blocks depends only on gridDim.x.
  // warp w handles accumulators w, w+8, ...; lanes stride over blocks, then a fixed shuffle tree.
  const int nb = gridDim.x;
  for (int j = warp; j < NACC; j += ZNLL_BLOCK / 32) {
    if (j == 1) continue;  // handled together with j == 0
    if (j == 0) {
      double s = 0.0, cc = 0.0;
      for (int b = lane; b < nb; b += 32)
        kahan_merge(s, cc, __ldcg(&partials[(size_t)b * NACC + 0]), __ldcg(&partials[(size_t)b * NACC + 1]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double s2 = __shfl_down_sync(0xffffffffu, s, o);
        const double c2 = __shfl_down_sync(0xffffffffu, cc, o);
        kahan_merge(s, cc, s2, c2);
      }
      if (lane == 0) {
        out[0] = s;
        out[1] = cc;
      }
    } else {
      double s = 0.0;
      for (int b = lane; b < nb; b += 32) s += __ldcg(&partials[(size_t)b * NACC + j]);
      s = warp_sum(s);
      if (lane == 0) out[j] = s;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch on this stream
}

template <class M, bool WEIGHTED>
struct EventPair {
  double xa[ZNLL_MAXCOL], xb[ZNLL_MAXCOL];
  double wa, wb;
  template <int NC>
  __device__ __forceinline__ void load(const KArgs<NC>& args, long long i) {
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k) {
      if ((M::COLMASK >> k) & 1) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(args.cols[k]) + i);
        xa[k] = v.x;
        xb[k] = v.y;
      }
    }
    wa = wb = 1.0;
    if constexpr (WEIGHTED) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(args.w) + i);
      wa = v.x;
      wb = v.y;
    }
  }
};

template <class M, bool WEIGHTED, bool GRAD>
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS) nll_kernel(const __grid_constant__ KArgs<M::NC> args) {
  constexpr int NACC = 2 + M::NS;
  __shared__ double2 logtab[128];
  init_logtab(logtab);
  __syncthreads();
  Ctx cx;
  cx.logtab = logtab;
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  const double off = args.log_offset;
  const long long nvec = args.n >> 1;
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x;
  EventPair<M, WEIGHTED> cur, nxt;
  if (i < nvec) nxt.load(args, i);
  while (i < nvec) {
    cur = nxt;
    i += stride;
    if (i < nvec) nxt.load(args, i);  // software prefetch: the next pair is in flight while this one computes
    const double ta = event<M, WEIGHTED, GRAD>(cx, c, cur.xa, cur.wa, off, acc);
    const double tb = event<M, WEIGHTED, GRAD>(cx, c, cur.xb, cur.wb, off, acc);
    kahan_add(acc, ta + tb);
  }
  if ((args.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail event
    double xa[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) xa[k] = args.cols[k][args.n - 1];
    double wa = 1.0;
    if constexpr (WEIGHTED) wa = args.w[args.n - 1];
    kahan_add(acc, event<M, WEIGHTED, GRAD>(cx, c, xa, wa, off, acc));
  }
  block_reduce_and_finish<NACC>(acc, args.partials, args.ticket, args.out);
}

// per-event normalised pdf values (BasePDF.pdf, core/basepdf.py:398-429): out[i] = P(x_i); args.partials = out
template <class M>
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS) pdf_kernel(const __grid_constant__ KArgs<M::NC> args) {
  __shared__ double2 logtab[128];
  init_logtab(logtab);
  __syncthreads();
  Ctx cx;
  cx.logtab = logtab;
  const double* __restrict__ c = args.c;
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  for (long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x; i < args.n; i += stride) {
    double x[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) x[k] = __ldg(args.cols[k] + i);
    M m;
    double p;
    if constexpr (M::LOGNAT) {
      bool neg = false;
      const double lp = m.template fwd_log<0>(cx, c, x, neg);
      p = neg ? -exp(lp) : exp(lp);
    } else {
      p = m.template fwd_val<0>(cx, c, x);
    }
    args.partials[i] = p;
  }
}

// sum of weights with the same deterministic reduction (Data.samplesize, core/data.py:268-275)
template <int UNUSED>
__global__ void __launch_bounds__(ZNLL_BLOCK, 2)
sumw_kernel(const double* __restrict__ w, long long n, double* partials, unsigned int* ticket, double* out) {
  double acc[2] = {0.0, 0.0};
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  for (long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x; i < n; i += stride) {
    const double y = w[i] - acc[1];
    const double s = acc[0] + y;
    acc[1] = (s - acc[0]) - y;
    acc[0] = s;
  }
  block_reduce_and_finish<2>(acc, partials, ticket, out);
}

}  // namespace znll
