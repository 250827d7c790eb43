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
// The FP64 pipe is the roofline, and its enemy is the dependent DFMA chain (Horner polynomials):
// all node code is therefore written once over a value type V that is either `double` or `D2`
// -- the two events a thread loads with one double2 -- so that every source-level operation
// emits two independent instructions back to back (2-way ILP inside each warp) and every rare-path
// branch (out-of-range exp/log arguments, CrystalBall tail) is taken once per pair.
//
// The same header is compiled ahead of time by nvcc (znll_catalog_*.cu) and, for tree shapes that
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
#define ZNLL_MIN_BLOCKS 2  // resident CTAs per SM the register allocation is tuned for
#endif
#define ZNLL_LOG_EPS 1e-307

namespace znll {

// ----------------------------------------------------------------------------------------------
// two-lane value type
// ----------------------------------------------------------------------------------------------
struct D2 {
  double a, b;
};
struct M2 {
  bool a, b;
};

__device__ __forceinline__ double la(double x) { return x; }
__device__ __forceinline__ double lb(double x) { return x; }
__device__ __forceinline__ double la(D2 x) { return x.a; }
__device__ __forceinline__ double lb(D2 x) { return x.b; }
__device__ __forceinline__ bool la(bool x) { return x; }
__device__ __forceinline__ bool lb(bool x) { return x; }
__device__ __forceinline__ bool la(M2 x) { return x.a; }
__device__ __forceinline__ bool lb(M2 x) { return x.b; }

template <class V>
struct Lanes;
template <>
struct Lanes<double> {
  typedef bool Mask;
  static __device__ __forceinline__ double make(double a, double) { return a; }
  static __device__ __forceinline__ bool mask(bool a, bool) { return a; }
  static constexpr int N = 1;
};
template <>
struct Lanes<D2> {
  typedef M2 Mask;
  static __device__ __forceinline__ D2 make(double a, double b) {
    D2 r;
    r.a = a;
    r.b = b;
    return r;
  }
  static __device__ __forceinline__ M2 mask(bool a, bool b) {
    M2 r;
    r.a = a;
    r.b = b;
    return r;
  }
  static constexpr int N = 2;
};

// result type of a mixed scalar / two-lane expression
template <class A, class B = double, class C = double>
struct Wide {
  typedef double T;
};
template <class B, class C>
struct Wide<D2, B, C> {
  typedef D2 T;
};
template <class C>
struct Wide<double, D2, C> {
  typedef D2 T;
};
template <>
struct Wide<double, double, D2> {
  typedef D2 T;
};

#define ZNLL_DEV __device__ __forceinline__

template <class A, class B, class C>
ZNLL_DEV typename Wide<A, B, C>::T zfma(A x, B y, C z) {
  typedef typename Wide<A, B, C>::T R;
  return Lanes<R>::make(fma(la(x), la(y), la(z)), fma(lb(x), lb(y), lb(z)));
}
template <class A, class B>
ZNLL_DEV typename Wide<A, B>::T zmul(A x, B y) {
  typedef typename Wide<A, B>::T R;
  return Lanes<R>::make(la(x) * la(y), lb(x) * lb(y));
}
template <class A, class B>
ZNLL_DEV typename Wide<A, B>::T zadd(A x, B y) {
  typedef typename Wide<A, B>::T R;
  return Lanes<R>::make(la(x) + la(y), lb(x) + lb(y));
}
template <class A, class B>
ZNLL_DEV typename Wide<A, B>::T zsub(A x, B y) {
  typedef typename Wide<A, B>::T R;
  return Lanes<R>::make(la(x) - la(y), lb(x) - lb(y));
}
template <class V>
ZNLL_DEV V zneg(V x) {
  return Lanes<V>::make(-la(x), -lb(x));
}
template <class V>
ZNLL_DEV V zabs(V x) {
  return Lanes<V>::make(fabs(la(x)), fabs(lb(x)));
}
template <class V>
ZNLL_DEV typename Lanes<V>::Mask zlt(V x, double y) {
  return Lanes<V>::mask(la(x) < y, lb(x) < y);
}
// x < 0 read off the sign bit (true for -0.0 and negative NaNs as well): one integer compare instead of a DSETP
template <class V>
ZNLL_DEV typename Lanes<V>::Mask zsignbit(V x) {
  return Lanes<V>::mask(__double2hiint(la(x)) < 0, __double2hiint(lb(x)) < 0);
}
template <class V>
ZNLL_DEV bool zany_signbit(V x) {
  return (__double2hiint(la(x)) | __double2hiint(lb(x))) < 0;
}
template <class M>
ZNLL_DEV bool zany(M m) {
  return la(m) | lb(m);  // not ||: both lanes are known, a second branch would cost more than the OR
}
template <class M>
ZNLL_DEV bool zall(M m) {
  return la(m) && lb(m);
}
ZNLL_DEV bool mxor(bool p, bool q) { return p != q; }
ZNLL_DEV M2 mxor(M2 p, M2 q) {
  M2 r;
  r.a = p.a != q.a;
  r.b = p.b != q.b;
  return r;
}
template <class M, class V>
ZNLL_DEV V zsel(M m, V x, V y) {
  return Lanes<V>::make(la(m) ? la(x) : la(y), lb(m) ? lb(x) : lb(y));
}
template <class V>
ZNLL_DEV V zbc(double x) {
  return Lanes<V>::make(x, x);
}
// acc += sum over lanes of g*d
ZNLL_DEV void zacc(double& acc, double g, double d) { acc = fma(g, d, acc); }
ZNLL_DEV void zacc(double& acc, D2 g, D2 d) { acc = fma(g.b, d.b, fma(g.a, d.a, acc)); }
ZNLL_DEV void zacc(double& acc, D2 g, double d) { acc = fma(g.a + g.b, d, acc); }
// per-lane accumulators (covariance pass: the two events of a pair keep separate score vectors)
ZNLL_DEV void zacc(D2& acc, D2 g, D2 d) {
  acc.a = fma(g.a, d.a, acc.a);
  acc.b = fma(g.b, d.b, acc.b);
}
ZNLL_DEV void zacc(D2& acc, D2 g, double d) {
  acc.a = fma(g.a, d, acc.a);
  acc.b = fma(g.b, d, acc.b);
}
ZNLL_DEV void zacc1(D2& acc, D2 g) {
  acc.a += g.a;
  acc.b += g.b;
}
ZNLL_DEV void zacc1(double& acc, double g) { acc += g; }
ZNLL_DEV void zacc1(double& acc, D2 g) { acc += g.a + g.b; }

// ----------------------------------------------------------------------------------------------
// fp64 elementary functions tuned for this kernel:
//   * polynomial coefficients live in the constant bank, so every DFMA takes them as an operand
//     (libdevice materialises each one with two UMOVs -> more issue slots than DP work);
//   * log uses a 128-entry (1/c, log c) table in shared memory and a degree-7 log1p polynomial:
//     13 DP instructions instead of ~35 (no division); exp: 15 DP instructions;
//   * reciprocal: MUFU.RCP64H seed + two Newton steps (4 DFMA);
//   * both lanes of a D2 run interleaved; the libdevice path (|x| >= 700 for exp; non-positive,
//     subnormal, inf or NaN for log / rcp) is taken once per pair.
// Accuracy is tested on the GPU against libm (tests/test_math_gpu.py).
// ----------------------------------------------------------------------------------------------
static __constant__ double kExpC[10] = {
    0.5000000000000001,     0.16666666666666669,    0.04166666666662413,    0.008333333333330062,
    0.0013888888917213717,  0.00019841269863053618, 2.4801521295954376e-05, 2.7557268459997064e-06,
    2.7620088445409746e-07, 2.510038549551032e-08};
static __constant__ double kLogC[6] = {-0.5,
                                       0.33333333333333337,
                                       -0.24999999998288297,
                                       0.19999999998478485,
                                       -0.16666959217649738,
                                       0.14285974331115564};
static __constant__ double kMathC[8] = {
    1.4426950408889634,           // 0 log2(e)
    6755399441055744.0,           // 1 1.5 * 2^52
    -6.93147180369123816490e-01,  // 2 -ln2_hi
    -1.90821492927058770002e-10,  // 3 -ln2_lo
    0x1.62e42fefa3800p-1,         // 4 ln2_hi (42 significant bits: k*ln2_hi is exact)
    0x1.ef35793c76730p-45,        // 5 ln2_lo
    4503601774854144.0,           // 6 2^52 + 2^31
    0.0};

// ZNLL_EXP_TABLE = 1 selects a table-driven exp (see exp_core); 0 the table-free degree-11 polynomial.  The modelled gain
// of the table is under 2 % of an event pair's issue cycles and it adds a shared-memory look-up with data-dependent
// bank conflicts, so it stays off until a GPU measurement says otherwise.
#ifndef ZNLL_EXP_TABLE
#define ZNLL_EXP_TABLE 0
#endif
#if ZNLL_EXP_TABLE
// 2^(j/128), j = 0..127, correctly rounded (generated with mpmath); copied into shared memory by init_tabs
static __constant__ double kExp2Tab[128] = {
    0x1.0000000000000p+0, 0x1.0163da9fb3335p+0, 0x1.02c9a3e778061p+0, 0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0, 0x1.0706b29ddf6dep+0, 0x1.0874518759bc8p+0, 0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0, 0x1.0cc922b7247f7p+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0, 0x1.12abdc06c31ccp+0, 0x1.1429aaea92de0p+0, 0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0, 0x1.18af9388c8deap+0, 0x1.1a35beb6fcb75p+0, 0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0, 0x1.1ed5022fcd91dp+0, 0x1.2063b88628cd6p+0, 0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0, 0x1.251ce4fb2a63fp+0, 0x1.26b4565e27cddp+0, 0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0, 0x1.2b87fd0dad990p+0, 0x1.2d285a6e4030bp+0, 0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0, 0x1.32170fc4cd831p+0, 0x1.33c08b26416ffp+0, 0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0, 0x1.38cae6d05d866p+0, 0x1.3a7db34e59ff7p+0, 0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0, 0x1.3fa4504ac801cp+0, 0x1.4160a21f72e2ap+0, 0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0, 0x1.46a41ed1d0057p+0, 0x1.486a2b5c13cd0p+0, 0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0, 0x1.4dcb299fddd0dp+0, 0x1.4f9b2769d2ca7p+0, 0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0, 0x1.551a4ca5d920fp+0, 0x1.56f4736b527dap+0, 0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0, 0x1.5c9268a5946b7p+0, 0x1.5e76f15ad2148p+0, 0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0, 0x1.6434634ccc320p+0, 0x1.6623882552225p+0, 0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6c012750bdabfp+0, 0x1.6dfb23c651a2fp+0, 0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0, 0x1.73f9a48a58174p+0, 0x1.75feb564267c9p+0, 0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0, 0x1.7c1ed0130c132p+0, 0x1.7e2f336cf4e62p+0, 0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0, 0x1.8471a4623c7adp+0, 0x1.868d99b4492edp+0, 0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8cf3216b5448cp+0, 0x1.8f1ae99157736p+0, 0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0, 0x1.95a44cbc8520fp+0, 0x1.97d829fde4e50p+0, 0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0, 0x1.9e86319e32323p+0, 0x1.a0c667b5de565p+0, 0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0, 0x1.a799e1330b358p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0, 0x1.b0e07298db666p+0, 0x1.b33a2b84f15fbp+0, 0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.ba5b030a1064ap+0, 0x1.bcc1e904bc1d2p+0, 0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c67f12e57d14bp+0, 0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0, 0x1.cdf0b555dc3fap+0, 0x1.d072d4a07897cp+0, 0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0, 0x1.d80e316c98398p+0, 0x1.da9e603db3285p+0, 0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e264614f5a129p+0, 0x1.e502ee78b3ff6p+0, 0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0, 0x1.ecf482d8e67f1p+0, 0x1.efa1bee615a27p+0, 0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0, 0x1.f7bfdad9cbe14p+0, 0x1.fa7c1819e90d8p+0, 0x1.fd3c22b8f71f1p+0,
};
static __constant__ double kExpT[8] = {
    0x1.71547652b82fep+7,    // 0 128/ln2
    6755399441055744.0,      // 1 1.5 * 2^52
    -0x1.62e42fef80000p-8,   // 2 -(ln2/128)_hi: 35 significant bits, n * hi is exact for |n| < 2^18
    -0x1.1cf79abc9e3b4p-43,  // 3 -(ln2/128)_lo
    0x1.1111111111111p-7,    // 4 1/120
    0x1.5555555555555p-5,    // 5 1/24
    0x1.5555555555555p-3,    // 6 1/6
    0.0};
#endif

// per-block tables of the elementary functions, in shared memory
struct MathTabs {
  double2 log[128];  // (1/c_i, log c_i) for 128 intervals of [0.6875, 1.375)
#if ZNLL_EXP_TABLE
  double exp2[128];  // 2^(j/128)
#endif
};
struct Ctx {
  const double2* logtab;
  const double* exptab;
};

ZNLL_DEV bool exp_in_range(double x) { return (__double2hiint(x) & 0x7fffffff) < 0x4085E000; }  // |x| < 700
ZNLL_DEV bool log_in_range(double x) { return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u; }
// x < -690 read off the high word (the sign bit set and the magnitude above 690; -inf and negative NaNs included)
ZNLL_DEV bool below_m690(double x) { return (unsigned)__double2hiint(x) > 0xC0859000u; }
ZNLL_DEV bool rcp_in_range(double x) { return (((unsigned)__double2hiint(x) >> 20) & 0x7ffu) - 1u < 0x7fdu; }
ZNLL_DEV double scale2n(double p, int n) { return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p)); }
ZNLL_DEV double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

#if ZNLL_EXP_TABLE
// exp(x) = 2^k * 2^(j/128) * e^r with n = 128 k + j = nearest(x * 128/ln2) and |r| <= ln2/256: a table look-up in shared
// memory and a degree-5 Taylor polynomial (truncation 6e-19) -- 10 DP instructions instead of the 17 of the table-free
// version, for one LDS and three integer instructions more.  Total error about 1 ulp (table 0.5 + final FMA 0.5).
ZNLL_DEV double exp_tab(int n, const Ctx& cx) { return cx.exptab[n & 127]; }
template <class V>
ZNLL_DEV V exp_core(V x, const Ctx& cx) {  // |x| < 700
  const V t = zfma(x, kExpT[0], kExpT[1]);
  const int na = __double2loint(la(t)), nb = __double2loint(lb(t));
  const V T = Lanes<V>::make(exp_tab(na, cx), exp_tab(nb, cx));
  const V nd = zsub(t, kExpT[1]);
  V r = zfma(nd, kExpT[2], x);
  r = zfma(nd, kExpT[3], r);
  V q = zfma(r, kExpT[4], kExpT[5]);
  q = zfma(q, r, kExpT[6]);
  q = zfma(q, r, 0.5);
  const V p = zfma(zmul(r, r), q, r);  // e^r - 1
  const V v = zfma(T, p, T);
  return Lanes<V>::make(scale2n(la(v), na >> 7), scale2n(lb(v), nb >> 7));
}
#else
template <class V>
ZNLL_DEV V exp_core(V x, const Ctx&) {  // |x| < 700
  const V t = zfma(x, kMathC[0], kMathC[1]);
  const V nd = zsub(t, kMathC[1]);
  V r = zfma(nd, kMathC[2], x);
  r = zfma(nd, kMathC[3], r);
  V p = zfma(r, kExpC[9], kExpC[8]);
#pragma unroll
  for (int k = 7; k >= 0; --k) p = zfma(p, r, kExpC[k]);
  p = zfma(p, r, 1.0);
  p = zfma(p, r, 1.0);
  return Lanes<V>::make(scale2n(la(p), __double2loint(la(t))), scale2n(lb(p), __double2loint(lb(t))));
}
#endif
template <class V>
ZNLL_DEV V fast_exp(V x, const Ctx& cx) {
  if (!(exp_in_range(la(x)) && exp_in_range(lb(x)))) return Lanes<V>::make(exp(la(x)), exp(lb(x)));  // rare
  return exp_core(x, cx);
}
// exp(-|d|): the range test reads d itself, so -|d| never has to exist in a register (the DFMAs of the core take it
// through their operand modifiers)
template <class V>
ZNLL_DEV V fast_exp_negabs(V d, const Ctx& cx) {
  if (!(exp_in_range(la(d)) && exp_in_range(lb(d)))) return Lanes<V>::make(exp(-fabs(la(d))), exp(-fabs(lb(d))));  // rare
  return exp_core(zneg(zabs(d)), cx);
}

template <class V>
ZNLL_DEV V rcp_core(V x) {  // x normal, 1/x normal
  V y = Lanes<V>::make(rcp_seed(la(x)), rcp_seed(lb(x)));
  const V nx = zneg(x);
  V e = zfma(nx, y, 1.0);
  y = zfma(y, e, y);
  e = zfma(nx, y, 1.0);
  return zfma(y, e, y);
}
template <class V>
ZNLL_DEV V fast_rcp(V x) {
  if (!(rcp_in_range(la(x)) && rcp_in_range(lb(x)))) return Lanes<V>::make(1.0 / la(x), 1.0 / lb(x));  // rare
  return rcp_core(x);
}

struct LogSplit {
  double z, kd;
  double2 tc;
};
ZNLL_DEV LogSplit log_split(double x, const Ctx& cx) {
  LogSplit s;
  const int hi = __double2hiint(x);
  const int tmp = hi - 0x3fe60000;
  s.z = __hiloint2double(hi - (tmp & 0xfff00000), __double2loint(x));
  s.tc = cx.logtab[(tmp >> 13) & 127];
  s.kd = __hiloint2double(0x43300000, (tmp >> 20) ^ 0x80000000);
  return s;
}

template <class V>
ZNLL_DEV V log_core(V x, const Ctx& cx) {  // x positive, normal, finite
  const LogSplit sa = log_split(la(x), cx), sb = log_split(lb(x), cx);
  const V z = Lanes<V>::make(sa.z, sb.z), invc = Lanes<V>::make(sa.tc.x, sb.tc.x), logc = Lanes<V>::make(sa.tc.y, sb.tc.y);
  const V kd = zsub(Lanes<V>::make(sa.kd, sb.kd), kMathC[6]);
  const V r = zfma(z, invc, -1.0);
  V s = zfma(r, kLogC[5], kLogC[4]);
#pragma unroll
  for (int j = 3; j >= 0; --j) s = zfma(s, r, kLogC[j]);
  const V w = zfma(kd, kMathC[4], logc);
  const V t2 = zfma(zmul(r, r), s, zmul(kd, kMathC[5]));
  return zadd(w, zadd(r, t2));
}
template <class V>
ZNLL_DEV V fast_log(V x, const Ctx& cx) {
  if (!(log_in_range(la(x)) && log_in_range(lb(x)))) return Lanes<V>::make(log(la(x)), log(lb(x)));  // rare
  return log_core(x, cx);
}

// log and reciprocal of the SAME argument behind ONE range test (the step kernels are issue-bound: a test is ~4 integer
// instructions per lane plus a branch and its convergence barrier).  Fast path: x positive and normal with a normal
// reciprocal (biased exponent 1..2045); everything else -- zero, negative, subnormal, huge, inf, NaN -- is answered by
// libdevice once per pair, so the IEEE results (log of a negative number is NaN, 1/0 is inf) are unchanged.
ZNLL_DEV bool pos_normal(double x) { return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fd00000u; }
template <class V>
ZNLL_DEV void fast_log_rcp(V x, const Ctx& cx, V& lg, V& rc) {
  if (!(pos_normal(la(x)) && pos_normal(lb(x)))) {  // rare
    lg = Lanes<V>::make(log(la(x)), log(lb(x)));
    rc = Lanes<V>::make(1.0 / la(x), 1.0 / lb(x));
    return;
  }
  lg = log_core(x, cx);
  rc = rcp_core(x);
}
// lg = log|x|, rc = 1/x, neg = (x < 0): for values that are positive in every sane fit (a polynomial pdf, a sum of
// fractions times pdfs) but whose sign must be tracked when they are not
template <class V>
ZNLL_DEV void fast_logabs_rcp(V x, const Ctx& cx, V& lg, V& rc, typename Lanes<V>::Mask& neg) {
  if (!(pos_normal(la(x)) && pos_normal(lb(x)))) {  // rare
    lg = Lanes<V>::make(log(fabs(la(x))), log(fabs(lb(x))));
    rc = Lanes<V>::make(1.0 / la(x), 1.0 / lb(x));
    neg = zlt(x, 0.0);
    return;
  }
  lg = log_core(x, cx);
  rc = rcp_core(x);
  neg = Lanes<V>::mask(false, false);
}

// one entry per thread < 128, then __syncthreads() by the caller
ZNLL_DEV void init_tabs(MathTabs& tabs) {
  if (threadIdx.x < 128) {
    const int ha = 0x3fe60000 + ((int)threadIdx.x << 13);
    const double a = __hiloint2double(ha, 0), b = __hiloint2double(ha + (1 << 13), 0);
    const double invc = 1.0 / (0.5 * (a + b));
    tabs.log[threadIdx.x] = make_double2(invc, -log(invc));
#if ZNLL_EXP_TABLE
    tabs.exp2[threadIdx.x] = kExp2Tab[threadIdx.x];
#endif
  }
}
ZNLL_DEV Ctx ctx_of(const MathTabs& tabs) {
  Ctx cx;
  cx.logtab = tabs.log;
#if ZNLL_EXP_TABLE
  cx.exptab = tabs.exp2;
#else
  cx.exptab = nullptr;
#endif
  return cx;
}

// ----------------------------------------------------------------------------------------------
// leaves.  c = const block, x = this event's (pair's) column values, acc = per-thread accumulators.
// CO / AO are compile-time offsets into c / acc, so every access is a fixed register or a
// constant-bank operand.  `neg` tracks the sign of a product with non-positive factors.
// ----------------------------------------------------------------------------------------------

// Gauss: consts [1/sigma, -mu/sigma, -(0.5*ln(2pi)+ln sigma) - ln I];  slots [mu, sigma]
template <int COL, class V>
struct GaussN {
  typedef typename Lanes<V>::Mask Mask;
  V t, P;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx&, const double* __restrict__ c, const V* x, Mask&) {
    t = zfma(x[COL], c[CO + 0], c[CO + 1]);  // x/sigma - mu/sigma (tfp Normal._log_prob)
    return zfma(zmul(t, -0.5), t, c[CO + 2]);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const V gi = zmul(g, c[CO + 0]);
    zacc(acc[AO + 0], gi, t);                  // d/dmu    =  t/sigma
    zacc(acc[AO + 1], gi, zfma(t, t, -1.0));   // d/dsigma = (t^2-1)/sigma
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct Gauss {
  static constexpr int NS = 2, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = GaussN<COL, V>;
};

// TruncatedGauss (models/dist_tfp.py:357-417, tfp TruncatedNormal): the Gauss shape, zero outside [low, high]; the
// truncation's own normaliser cancels against the integral over the norm range, which the host takes over the
// intersection of the two intervals.  consts [1/sigma, -mu/sigma, -(0.5*ln(2pi)+ln sigma) - ln I, low, high];
// slots [mu, sigma, low, high] (low / high only move the normalisation: no per-event derivative).
template <int COL, class V>
struct TGaussN {
  typedef typename Lanes<V>::Mask Mask;
  V t, P;
  Mask out;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx&, const double* __restrict__ c, const V* x, Mask&) {
    t = zfma(x[COL], c[CO + 0], c[CO + 1]);
    const double low = c[CO + 3], high = c[CO + 4];
    out = Lanes<V>::mask(la(x[COL]) < low || la(x[COL]) > high, lb(x[COL]) < low || lb(x[COL]) > high);
    // a huge negative log stands for "probability zero": exp() gives 0, sums and differences stay finite
    return zsel(out, zbc<V>(-1e300), zfma(zmul(t, -0.5), t, c[CO + 2]));
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const V gi = zsel(out, zbc<V>(0.0), zmul(g, c[CO + 0]));
    zacc(acc[AO + 0], gi, t);
    zacc(acc[AO + 1], gi, zfma(t, t, -1.0));
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct TGauss {
  static constexpr int NS = 4, NC = 5, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = TGaussN<COL, V>;
};

// Exponential: consts [lambda, shift, -ln I];  slots [lambda]
template <int COL, class V>
struct ExpoN {
  typedef typename Lanes<V>::Mask Mask;
  V xs, P;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx&, const double* __restrict__ c, const V* x, Mask&) {
    xs = zsub(x[COL], c[CO + 1]);
    return zfma(xs, c[CO + 0], c[CO + 2]);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__, A* acc) {
    zacc(acc[AO + 0], g, xs);
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct Expo {
  static constexpr int NS = 1, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = ExpoN<COL, V>;
};

// CrystalBall: slots [mu, sigma, alpha, n]
// consts [0 mu, 1 sign(alpha)/sigma, 2 1/sigma, 3 |alpha|, 4 n, 5 B=n/|alpha|-|alpha|,
//         6 ln A - ln I, 7 -ln I, 8 ln(n/|alpha|)+1, 9 1/|alpha|, 10 -n/|alpha|-|alpha|,
//         11 n/alpha^2+1, 12 sign(alpha)]
template <int COL, class V>
struct CBallN {
  typedef typename Lanes<V>::Mask Mask;
  V t, dt, lu, P;  // dt = d ln v / d t ; lu = ln(B-t) on the tail
  Mask tail;
  bool any_tail;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask&) {
    t = zmul(zsub(x[COL], c[CO + 0]), c[CO + 1]);
    // t < -|alpha|  <=>  t + |alpha| < 0, and that sum has the exact sign (it is exact when the two are within a factor
    // of two of each other, far from zero otherwise; t == -|alpha| gives +0): integer sign tests replace three DSETPs
    const V ta = zadd(t, c[CO + 3]);
    tail = zsignbit(ta);
    any_tail = zany_signbit(ta);
    const V lp_core = zfma(zmul(t, -0.5), t, c[CO + 7]);  // exp(-t^2/2)
    if (!any_tail) {                                       // whole pair on the Gaussian core (events are sorted)
      dt = zneg(t);
      lu = zbc<V>(0.0);
      return lp_core;
    }
    // A*(B-t)^-n == exp(ln A - n*ln(B-t)); a lane on the core gets the safe value u = 2 (z.safe_where)
    const V u = zsel(tail, zsub(c[CO + 5], t), zbc<V>(2.0));
    V ru;
    fast_log_rcp(u, cx, lu, ru);  // one range test for both
    const V dtt = zmul(ru, c[CO + 4]);
    const V lp_tail = zfma(lu, -c[CO + 4], c[CO + 6]);
    dt = zsel(tail, dtt, zneg(t));
    return zsel(tail, lp_tail, lp_core);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const V gd = zmul(g, dt);
    zacc(acc[AO + 0], zneg(gd), c[CO + 1]);             // dt/dmu    = -sign/sigma
    zacc(acc[AO + 1], zmul(zneg(gd), t), c[CO + 2]);    // dt/dsigma = -t/sigma
    if (any_tail) {
      // d ln v/d|alpha| = -n/a - a + (n/u)(n/a^2+1) ; d ln v/dn = ln(n/a)+1 - ln u - (n/u)/a ; 0 on the core
      const V gt = zsel(tail, g, zbc<V>(0.0));
      zacc(acc[AO + 2], zmul(gt, c[CO + 12]), zfma(dt, c[CO + 11], c[CO + 10]));
      zacc(acc[AO + 3], gt, zsub(zsub(c[CO + 8], lu), zmul(dt, c[CO + 9])));
    }
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct CBall {
  static constexpr int NS = 4, NC = 13, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = CBallN<COL, V>;
};

// GaussExpTail (models/physics.py:608-623, 727-805): Gaussian core, exponential tail  exp(a^2/2 + a t)  for t < -|alpha|.
// slots [mu, sigma, alpha];  consts [0 mu, 1 sign(alpha)/sigma, 2 1/sigma, 3 |alpha|, 4 a^2/2 - ln I, 5 -ln I, 6 sign(alpha)]
template <int COL, class V>
struct GETailN {
  typedef typename Lanes<V>::Mask Mask;
  V t, dt, P;
  Mask tail;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx&, const double* __restrict__ c, const V* x, Mask&) {
    t = zmul(zsub(x[COL], c[CO + 0]), c[CO + 1]);
    tail = zlt(t, -c[CO + 3]);
    dt = zsel(tail, zbc<V>(c[CO + 3]), zneg(t));  // d ln v / d t
    return zsel(tail, zfma(t, c[CO + 3], c[CO + 4]), zfma(zmul(t, -0.5), t, c[CO + 5]));
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const V gd = zneg(zmul(g, dt));
    zacc(acc[AO + 0], gd, c[CO + 1]);
    zacc(acc[AO + 1], zmul(gd, t), c[CO + 2]);
    zacc(acc[AO + 2], zsel(tail, g, zbc<V>(0.0)), zmul(zadd(t, c[CO + 3]), c[CO + 6]));  // d ln v/d|alpha| = a + t
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct GETail {
  static constexpr int NS = 3, NC = 7, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = GETailN<COL, V>;
};

// GeneralizedGaussExpTail (models/physics.py:625-634, 819-923): x < mu -> GaussExpTail(mu, sigmal, alphal), else
// GaussExpTail(mu, sigmar, -alphar).  slots [mu, sigmal, alphal, sigmar, alphar]
// consts [0 mu, 1 -ln I, per side (L at 2, R at 7): +0 sign(alpha_eff)/sigma, +1 1/sigma, +2 |alpha|, +3 a^2/2 - ln I,
//         +4 sign(alpha slot)]
template <int COL, class V>
struct GGETailN {
  typedef typename Lanes<V>::Mask Mask;
  V t, dt, P;
  Mask left, tail;

  template <int K>
  ZNLL_DEV V side(const double* __restrict__ c) const {
    return zsel(left, zbc<V>(c[2 + K]), zbc<V>(c[7 + K]));
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx&, const double* __restrict__ c, const V* x, Mask&) {
    const double* cc = c + CO;
    const V dx = zsub(x[COL], cc[0]);
    left = zlt(dx, 0.0);
    t = zmul(dx, side<0>(cc));
    const V a = side<2>(cc);
    tail = Lanes<V>::mask(la(t) < -la(a), lb(t) < -lb(a));
    dt = zsel(tail, a, zneg(t));
    return zsel(tail, zfma(t, a, side<3>(cc)), zfma(zmul(t, -0.5), t, cc[1]));
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const double* cc = c + CO;
    const V zero = zbc<V>(0.0);
    const V gd = zneg(zmul(g, dt));
    zacc(acc[AO + 0], gd, side<0>(cc));
    const V gs = zmul(zmul(gd, t), side<1>(cc));
    zacc1(acc[AO + 1], zsel(left, gs, zero));
    zacc1(acc[AO + 3], zsel(left, zero, gs));
    const V ga = zmul(zmul(zsel(tail, g, zero), side<4>(cc)), zadd(t, side<2>(cc)));
    zacc1(acc[AO + 2], zsel(left, ga, zero));
    zacc1(acc[AO + 4], zsel(left, zero, ga));
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct GGETail {
  static constexpr int NS = 5, NC = 12, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = GGETailN<COL, V>;
};

// Orthogonal-polynomial sums of degree DEG (models/polynomials.py:46-175): slots [c_0 .. c_DEG]
// consts [0 2/(hi-lo), 1 -(lo+hi)/(hi-lo), 2 1/I, 3.. c_k/I]
// KIND 3: Bernstein basis B_{k,DEG}(t), t = (x - lo)/(hi - lo) (rescale_zero_one + de_casteljau, :872-905).
// KIND 0: Chebyshev T_n (T_{n+1} = 2zT_n - T_{n-1}, polynomials.py:338-343); 1: Legendre
// ((n+1)P_{n+1} = (2n+1)zP_n - nP_{n-1}, :181-190); 2: Chebyshev of the second kind U_n (U_1 = 2z, :487-560)
template <int COL, int DEG, int KIND, class V>
struct PolyN {
  typedef typename Lanes<V>::Mask Mask;
  V T[DEG + 1];
  V P, rP;  // rP = 1/P, produced together with log|P| by fwd_log

  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx&, const double* __restrict__ c, const V* x) {
    const V z = zfma(x[COL], c[CO + 0], c[CO + 1]);
    if constexpr (KIND == 3) {  // Bernstein basis on t = z in [0, 1]: B_k = C(DEG, k) t^k (1 - t)^(DEG - k)
      const V u = zfma(z, -1.0, 1.0);
      V tp[DEG + 1], up[DEG + 1];
      tp[0] = up[0] = zbc<V>(1.0);
#pragma unroll
      for (int k = 1; k <= DEG; ++k) {
        tp[k] = zmul(tp[k - 1], z);
        up[k] = zmul(up[k - 1], u);
      }
      V p = zbc<V>(0.0);
      double binom = 1.0;
#pragma unroll
      for (int k = 0; k <= DEG; ++k) {
        T[k] = zmul(zmul(tp[k], up[DEG - k]), binom);
        p = zfma(T[k], c[CO + 3 + k], p);
        binom = binom * (double)(DEG - k) / (double)(k + 1);
      }
      P = p;
      return p;
    }
    T[0] = zbc<V>(1.0);
    V p = zbc<V>(c[CO + 3]);
    const V z2 = zadd(z, z);
    if constexpr (DEG >= 1) {
      T[1] = KIND == 2 ? z2 : z;
      p = zfma(T[1], c[CO + 4], p);
    }
#pragma unroll
    for (int k = 2; k <= DEG; ++k) {
      if constexpr (KIND == 1) {
        const double n = (double)(k - 1);
        T[k] = zsub(zmul(zmul(z, T[k - 1]), (2.0 * n + 1.0) / (n + 1.0)), zmul(T[k - 2], n / (n + 1.0)));
      } else {
        T[k] = zfma(z2, T[k - 1], zneg(T[k - 2]));
      }
      p = zfma(T[k], c[CO + 3 + k], p);
    }
    P = p;
    return p;
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    const V p = fwd_val<CO>(cx, c, x);
    V lg;
    Mask ng;
    fast_logabs_rcp(p, cx, lg, rP, ng);
    neg = mxor(neg, ng);
    return lg;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    const V ai = zmul(a, c[CO + 2]);
    if constexpr (KIND == 3)
      zacc(acc[AO + 0], ai, T[0]);
    else
      zacc1(acc[AO + 0], ai);
#pragma unroll
    for (int k = 1; k <= DEG; ++k) zacc(acc[AO + k], ai, T[k]);
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    bwd_val<CO, AO>(zmul(g, rP), c, acc);
  }
};
template <int COL, int DEG, int KIND>
struct Poly {
  static constexpr int NS = DEG + 1, NC = 3 + DEG + 1, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = false;
  template <class V>
  using Node = PolyN<COL, DEG, KIND, V>;
};
template <int COL, int DEG>
using Cheb = Poly<COL, DEG, 0>;
template <int COL, int DEG>
using Legd = Poly<COL, DEG, 1>;
template <int COL, int DEG>
using Chb2 = Poly<COL, DEG, 2>;
template <int COL, int DEG>
using Bern = Poly<COL, DEG, 3>;  // Bernstein (models/polynomials.py:872-1028); consts [0], [1] map x to t in [0, 1]

// Generalized (two-sided) CrystalBall (models/physics.py:47-80, 186-232): x < mu -> CrystalBall(mu, sigmal, alphal, nl),
// else CrystalBall(mu, sigmar, -alphar, nr); DoubleCB is the same node with sigmal = sigmar = sigma.
// slots [mu, sigmal, alphal, nl, sigmar, alphar, nr]
// consts [0 mu, 1 -ln I, then per side (L at 2, R at 13):
//         +0 sign(alpha_eff)/sigma, +1 1/sigma, +2 |alpha|, +3 n, +4 B, +5 ln A - ln I, +6 ln(n/a)+1, +7 1/a,
//         +8 -n/a-a, +9 n/a^2+1, +10 sign(alpha slot)]
template <int COL, class V>
struct GCBallN {
  typedef typename Lanes<V>::Mask Mask;
  V t, dt, lu, P;
  Mask left, tail;
  bool any_tail;

  template <int K>
  ZNLL_DEV V side(const double* __restrict__ c) const {  // per-lane selection of a side constant
    return zsel(left, zbc<V>(c[2 + K]), zbc<V>(c[13 + K]));
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask&) {
    const double* cc = c + CO;
    const V dx = zsub(x[COL], cc[0]);
    left = zlt(dx, 0.0);
    t = zmul(dx, side<0>(cc));
    const V a = side<2>(cc);
    tail = Lanes<V>::mask(la(t) < -la(a), lb(t) < -lb(a));
    any_tail = zany(tail);
    const V lp_core = zfma(zmul(t, -0.5), t, cc[1]);
    if (!any_tail) {
      dt = zneg(t);
      lu = zbc<V>(0.0);
      return lp_core;
    }
    const V u = zsel(tail, zsub(side<4>(cc), t), zbc<V>(2.0));
    V ru;
    fast_log_rcp(u, cx, lu, ru);
    const V n = side<3>(cc);
    const V dtt = zmul(ru, n);
    const V lp_tail = zfma(lu, zneg(n), side<5>(cc));
    dt = zsel(tail, dtt, zneg(t));
    return zsel(tail, lp_tail, lp_core);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const double* cc = c + CO;
    const V zero = zbc<V>(0.0);
    const V gd = zneg(zmul(g, dt));
    zacc(acc[AO + 0], gd, side<0>(cc));  // dt/dmu = -sign/sigma
    const V gs = zmul(zmul(gd, t), side<1>(cc));  // dt/dsigma = -t/sigma, to the side's own sigma
    zacc1(acc[AO + 1], zsel(left, gs, zero));
    zacc1(acc[AO + 4], zsel(left, zero, gs));
    if (any_tail) {
      const V gt = zsel(tail, g, zero);
      const V ga = zmul(zmul(gt, side<10>(cc)), zfma(dt, side<9>(cc), side<8>(cc)));
      const V gn = zmul(gt, zsub(zsub(side<6>(cc), lu), zmul(dt, side<7>(cc))));
      zacc1(acc[AO + 2], zsel(left, ga, zero));
      zacc1(acc[AO + 5], zsel(left, zero, ga));
      zacc1(acc[AO + 3], zsel(left, gn, zero));
      zacc1(acc[AO + 6], zsel(left, zero, gn));
    }
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct GCBall {
  static constexpr int NS = 7, NC = 24, COLMASK = 1 << COL;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = GCBallN<COL, V>;
};

// ----------------------------------------------------------------------------------------------
// functors
// ----------------------------------------------------------------------------------------------
template <class... Cs>
struct Meta;
template <>
struct Meta<> {
  static constexpr int NS = 0, NC = 0, COLMASK = 0;
};
template <class H, class... T>
struct Meta<H, T...> {
  static constexpr int NS = H::NS + Meta<T...>::NS;
  static constexpr int NC = H::NC + Meta<T...>::NC;
  static constexpr int COLMASK = H::COLMASK | Meta<T...>::COLMASK;
};

template <class V, class... Cs>
struct Kids;
template <class V>
struct Kids<V> {
  static constexpr int N = 0;
};
template <class V, class H, class... T>
struct Kids<V, H, T...> {
  typedef H Head;
  typename H::template Node<V> h;
  Kids<V, T...> t;
  static constexpr int N = 1 + sizeof...(T);
};

// SumPDF: consts [f_0..f_{K-1}]; slots [f_0..f_{K-1}]; then the children
template <class V, class... Cs>
struct SumN {
  typedef typename Lanes<V>::Mask Mask;
  static constexpr int K = sizeof...(Cs);
  Kids<V, Cs...> kids;
  V p[K];
  V P, rP;  // rP = 1/P, produced together with log|P| by fwd_log

  template <int CO, int I, int CK, class KD>
  ZNLL_DEV void fwd_rec(KD& kd, const Ctx& cx, const double* __restrict__ c, const V* x, V& tot) {
    if constexpr (KD::N > 0) {
      p[I] = kd.h.template fwd_val<CK>(cx, c, x);
      tot = zfma(p[I], c[CO + I], tot);
      fwd_rec<CO, I + 1, CK + KD::Head::NC>(kd.t, cx, c, x, tot);
    }
  }
  template <int CO, int AO, int I, int CK, int AK, class KD, class A>
  ZNLL_DEV void bwd_rec(KD& kd, V a, const double* __restrict__ c, A* acc) {
    if constexpr (KD::N > 0) {
      zacc(acc[AO + I], a, p[I]);
      kd.h.template bwd_val<CK, AK>(zmul(a, c[CO + I]), c, acc);
      bwd_rec<CO, AO, I + 1, CK + KD::Head::NC, AK + KD::Head::NS>(kd.t, a, c, acc);
    }
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    V tot = zbc<V>(0.0);
    fwd_rec<CO, 0, CO + K>(kids, cx, c, x, tot);
    P = tot;
    return tot;
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    const V v = fwd_val<CO>(cx, c, x);
    V lg;
    Mask ng;
    fast_logabs_rcp(v, cx, lg, rP, ng);
    neg = mxor(neg, ng);
    return lg;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_rec<CO, AO, 0, CO + K, AO + K>(kids, a, c, acc);
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    bwd_val<CO, AO>(zmul(g, rP), c, acc);
  }
};
// SumPDF of exactly two log-native children (the signal + background case), in log-sum-exp form:
//   ln P = m + ln S,  m = max(l0, l1),  S = f0*e^(l0-m) + f1*e^(l1-m)
// One of the two exponentials is exp(0) = 1, so a single exp is evaluated per event instead of two, it can
// neither overflow nor lose the event to underflow, and the Sum becomes log-native itself (no exp at all
// when it sits under a ProductPDF).  d ln P/d f_i = e_i/S;  the adjoint of child i is g*f_i*e_i/S.
template <class V, class C0, class C1>
struct SumLseN {
  typedef typename Lanes<V>::Mask Mask;
  typename C0::template Node<V> k0;
  typename C1::template Node<V> k1;
  V q0, q1;  // e_i / S
  V P;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    Mask n0 = Lanes<V>::mask(false, false), n1 = n0;
    const V l0 = k0.template fwd_log<CO + 2>(cx, c, x, n0);
    const V l1 = k1.template fwd_log<CO + 2 + C0::NC>(cx, c, x, n1);
    const V d = zsub(l0, l1);
    const Mask lt = zlt(d, 0.0);  // l0 < l1
    const V e = fast_exp_negabs(d, cx);
    const V one = zbc<V>(1.0);
    V e0 = zsel(lt, e, one), e1 = zsel(lt, one, e);
    e0 = zsel(n0, zneg(e0), e0);  // children that are themselves products with a negative factor
    e1 = zsel(n1, zneg(e1), e1);
    const V S = zfma(e0, c[CO + 0], zmul(e1, c[CO + 1]));
    V lS, rS;
    Mask ng;
    fast_logabs_rcp(S, cx, lS, rS, ng);  // one range test for both; S < 0 (a negative fraction) is the rare path
    neg = mxor(neg, ng);
    q0 = zmul(e0, rS);
    q1 = zmul(e1, rS);
    return zadd(zsel(lt, l1, l0), lS);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg = Lanes<V>::mask(false, false);
    const V e = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    P = zsel(neg, zneg(e), e);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    const V g0 = zmul(g, q0), g1 = zmul(g, q1);
    zacc1(acc[AO + 0], g0);
    zacc1(acc[AO + 1], g1);
    k0.template bwd_log<CO + 2, AO + 2>(zmul(g0, c[CO + 0]), c, acc);
    k1.template bwd_log<CO + 2 + C0::NC, AO + 2 + C0::NS>(zmul(g1, c[CO + 1]), c, acc);
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};

template <bool B, class T, class F>
struct If {
  typedef T type;
};
template <class T, class F>
struct If<false, T, F> {
  typedef F type;
};
template <class V, class... Cs>
struct SumPick {
  typedef SumN<V, Cs...> type;
};
template <class V, class C0, class C1>
struct SumPick<V, C0, C1> {
  typedef typename If<C0::LOGNAT && C1::LOGNAT, SumLseN<V, C0, C1>, SumN<V, C0, C1>>::type type;
};
template <class... Cs>
struct AllLog {
  static constexpr bool value = false;
};
template <class C0, class C1>
struct AllLog<C0, C1> {
  static constexpr bool value = C0::LOGNAT && C1::LOGNAT;
};

template <class... Cs>
struct Sum {
  static constexpr int K = sizeof...(Cs);
  static constexpr int NS = K + Meta<Cs...>::NS, NC = K + Meta<Cs...>::NC;
  static constexpr int COLMASK = Meta<Cs...>::COLMASK;
  static constexpr bool LOGNAT = AllLog<Cs...>::value;
  template <class V>
  using Node = typename SumPick<V, Cs...>::type;
};

// ProductPDF over disjoint observables: no consts, no slots of its own
template <class V, class... Cs>
struct ProdN {
  typedef typename Lanes<V>::Mask Mask;
  Kids<V, Cs...> kids;
  V P;

  template <int CK, class KD>
  ZNLL_DEV void fwd_rec(KD& kd, const Ctx& cx, const double* __restrict__ c, const V* x, V& tot, Mask& neg) {
    if constexpr (KD::N > 0) {
      tot = zadd(tot, kd.h.template fwd_log<CK>(cx, c, x, neg));
      fwd_rec<CK + KD::Head::NC>(kd.t, cx, c, x, tot, neg);
    }
  }
  template <int CK, int AK, class KD, class A>
  ZNLL_DEV void bwd_rec(KD& kd, V g, const double* __restrict__ c, A* acc) {
    if constexpr (KD::N > 0) {
      kd.h.template bwd_log<CK, AK>(g, c, acc);
      bwd_rec<CK + KD::Head::NC, AK + KD::Head::NS>(kd.t, g, c, acc);
    }
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    V tot = zbc<V>(0.0);
    fwd_rec<CO>(kids, cx, c, x, tot, neg);
    return tot;
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg = Lanes<V>::mask(false, false);
    const V e = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    P = zsel(neg, zneg(e), e);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    bwd_rec<CO, AO>(kids, g, c, acc);
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <class... Cs>
struct Prod {
  static constexpr int NS = Meta<Cs...>::NS, NC = Meta<Cs...>::NC;
  static constexpr int COLMASK = Meta<Cs...>::COLMASK;
  static constexpr bool LOGNAT = true;
  template <class V>
  using Node = ProdN<V, Cs...>;
};

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
// Cross-GPU exchange fused into the last block of the (last channel's) kernel: every rank pushes its raw
// accumulators into all peers' exchange buffers with plain stores over NVLink (peer-mapped symmetric memory),
// publishes a sequence flag with release semantics at system scope, waits for the peers' flags, and sums the
// world's contributions in fixed rank order -- bit-identical on every rank, no NCCL call on the step path.
// Buffer layout per rank: [128 B: u64 flag per source rank][2 parities][world][total] doubles.
#define ZNLL_MAXPEERS 8
#define ZNLL_XFLAG_BYTES 128
struct Xchg {
  double* buf[ZNLL_MAXPEERS];  // buf[r]: rank r's exchange buffer as mapped into this process
  double* acc;                 // this rank's accumulators of ALL channels (plan-wide), reduced in place
  unsigned long long seq;      // evaluation counter (> 0), identical on all ranks
  int rank, world, total, pad;
  double* hout;                // host-mapped pinned copy of acc (zero-copy result path), or null
  unsigned long long* hflag;   // host-mapped flag: set to seq after hout is complete
};

// Result hand-off without a D2H copy or a stream synchronisation: the last block writes the final accumulators
// straight into host-mapped pinned memory and then publishes the evaluation's sequence number; the host spins on it.
ZNLL_DEV void host_publish(const Xchg& xc) {
  __syncthreads();
  if (threadIdx.x < xc.total) xc.hout[threadIdx.x] = __ldcg(xc.acc + threadIdx.x);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(xc.hflag), "l"(xc.seq) : "memory");
}

ZNLL_DEV void peer_exchange(const Xchg& xc) {
  __shared__ int timed_out;
  const int t = threadIdx.x;
  if (t == 0) timed_out = 0;
  __syncthreads();  // all of this kernel's out[] stores are done (and earlier channels' kernels have completed)
  const int par = (int)(xc.seq & 1ull);
  if (t < xc.total) {
    const double v = __ldcg(xc.acc + t);
    for (int p = 0; p < xc.world; ++p) {
      double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(xc.buf[p]) + ZNLL_XFLAG_BYTES) +
                    ((size_t)(par * xc.world + xc.rank) * xc.total + t);
      *reinterpret_cast<volatile double*>(dst) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < xc.world) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(xc.buf[t]) + xc.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(xc.seq) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(xc.buf[xc.rank]) + t;
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    } while (v < xc.seq && (clock64() - t0) < 60000000000ll);  // ~30 s: a dead peer must not hang the GPU
    if (v < xc.seq) timed_out = 1;
  }
  __syncthreads();
  if (t < xc.total) {
    const double* src = reinterpret_cast<const double*>(reinterpret_cast<const char*>(xc.buf[xc.rank]) + ZNLL_XFLAG_BYTES) +
                        (size_t)par * xc.world * xc.total + t;
    double s = 0.0;
    for (int r = 0; r < xc.world; ++r) s += __ldcg(src + (size_t)r * xc.total);  // fixed rank order
    xc.acc[t] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

template <int NC>
struct KArgs {
  const double* cols[ZNLL_MAXCOL];
  const double* w;
  long long n;
  double log_offset;
  double* partials;        // [gridDim.x][NACC]
  unsigned int* ticket;    // zero before launch; reset by the last block
  double* out;             // [NACC]
  Xchg xc;                 // cross-GPU exchange (world == 0: none)
  double c[NC > 0 ? NC : 1];
};

// accumulators: [0] running sum of w*l - offset, [1] its Kahan compensation, [2..] slots.
// Returns the term(s)  w*l - offset  of this event (pair).
template <class M, class V, bool WEIGHTED, bool GRAD, class A>
ZNLL_DEV V event(const Ctx& cx, const double* __restrict__ c, const V* x, V w, double off, A* acc) {
  typedef typename Lanes<V>::Mask Mask;
  typename M::template Node<V> m;
  V l;
  if constexpr (M::LOGNAT) {
    Mask neg = Lanes<V>::mask(false, false);
    const V lp = m.template fwd_log<0>(cx, c, x, neg);
    l = lp;
    if (below_m690(la(lp)) | below_m690(lb(lp)) | zany(neg)) {  // p + 1e-307 matters only here (loss.py:117); rare
      double ll[2] = {la(lp), lb(lp)}, gg[2] = {1.0, 1.0};
      const bool ng[2] = {la(neg), lb(neg)};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (ll[k] < -690.0 || ng[k]) {
          const double p = ng[k] ? -exp(ll[k]) : exp(ll[k]);
          const double pp = p + ZNLL_LOG_EPS;
          ll[k] = log(pp);
          gg[k] = p / pp;
        }
      }
      l = Lanes<V>::make(ll[0], ll[1]);
      if constexpr (GRAD) {
        V g = Lanes<V>::make(gg[0], gg[1]);
        if constexpr (WEIGHTED) g = zmul(g, w);
        m.template bwd_log<0, 2>(g, c, acc);
      }
    } else if constexpr (GRAD) {
      // the adjoint of ln P is exactly 1 (or the weight): a compile-time constant the reverse sweep folds away
      if constexpr (WEIGHTED)
        m.template bwd_log<0, 2>(w, c, acc);
      else
        m.template bwd_log<0, 2>(zbc<V>(1.0), c, acc);
    }
  } else {
    const V P = m.template fwd_val<0>(cx, c, x);
    const V pp = zadd(P, ZNLL_LOG_EPS);
    if constexpr (GRAD) {
      V g;
      fast_log_rcp(pp, cx, l, g);  // one range test for both
      if constexpr (WEIGHTED) g = zmul(g, w);
      m.template bwd_val<0, 2>(g, c, acc);
    } else {
      l = fast_log(pp, cx);
    }
  }
  if constexpr (WEIGHTED)
    return zfma(w, l, -off);  // weights multiply first, then the unweighted offset (loss.py:134-137)
  else
    return zsub(l, off);
}

ZNLL_DEV void kahan_add(double* acc, double term) {
  const double y = term - acc[1];
  const double s = acc[0] + y;
  acc[1] = (s - acc[0]) - y;
  acc[0] = s;
}

ZNLL_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// error-free merge of two compensated sums (s - c is the value)
ZNLL_DEV void kahan_merge(double& s, double& c, double s2, double c2) {
  const double t = s + s2;
  const double bp = t - s;
  const double e = (s - (t - bp)) + (s2 - bp);  // TwoSum error of s + s2
  s = t;
  c = (c + c2) - e;
}

template <int NACC>
ZNLL_DEV void block_reduce_and_finish(double* acc, double* partials, unsigned int* ticket, double* out, const Xchg* xc) {
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
  if (xc != nullptr && xc->world > 1) peer_exchange(*xc);
  if (xc != nullptr && xc->hout != nullptr) host_publish(*xc);
}

template <class M, bool WEIGHTED>
struct EventPair {
  D2 x[ZNLL_MAXCOL];
  D2 w;
  template <int NC>
  ZNLL_DEV void load(const KArgs<NC>& args, long long i) {
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k) {
      if ((M::COLMASK >> k) & 1) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(args.cols[k]) + i);
        x[k].a = v.x;
        x[k].b = v.y;
      }
    }
    w.a = w.b = 1.0;
    if constexpr (WEIGHTED) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(args.w) + i);
      w.a = v.x;
      w.b = v.y;
    }
  }
};

#ifndef ZNLL_PINGPONG
#define ZNLL_PINGPONG 1
#endif
__host__ __device__ constexpr int zpopc(unsigned m) {
  int n = 0;
  while (m) {
    n += (int)(m & 1u);
    m >>= 1;
  }
  return n;
}
template <class M, bool WEIGHTED, bool GRAD>
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS) nll_kernel(const __grid_constant__ KArgs<M::NC> args) {
  constexpr int NACC = 2 + M::NS;
  __shared__ MathTabs tabs;
  init_tabs(tabs);
  __syncthreads();
  const Ctx cx = ctx_of(tabs);
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  const double off = args.log_offset;
  const long long nvec = args.n >> 1;
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x;
  if constexpr ((ZNLL_PINGPONG != 0) && (zpopc((unsigned)M::COLMASK) + (WEIGHTED ? 1 : 0) >= 3)) {
    // three or more 8-byte streams per event: two register buffers used alternately -- the loop body exists twice and no
    // buffer is ever copied (the copy is 12+ moves per pair there; measured -10 % on cfg 4, while one- and two-stream
    // trees lose 1 % to the doubled loop body and keep the copying loop below)
    EventPair<M, WEIGHTED> b0, b1;
    if (i < nvec) b0.load(args, i);
    while (i < nvec) {
      i += stride;
      if (i < nvec) b1.load(args, i);
      {
        const D2 t = event<M, D2, WEIGHTED, GRAD>(cx, c, b0.x, b0.w, off, acc);
        kahan_add(acc, t.a + t.b);
      }
      if (i >= nvec) break;
      i += stride;
      if (i < nvec) b0.load(args, i);
      {
        const D2 t = event<M, D2, WEIGHTED, GRAD>(cx, c, b1.x, b1.w, off, acc);
        kahan_add(acc, t.a + t.b);
      }
    }
  } else {
    EventPair<M, WEIGHTED> cur, nxt;
    if (i < nvec) nxt.load(args, i);
    while (i < nvec) {
      cur = nxt;
      i += stride;
      if (i < nvec) nxt.load(args, i);  // software prefetch: the next pair is in flight while this one computes
      const D2 t = event<M, D2, WEIGHTED, GRAD>(cx, c, cur.x, cur.w, off, acc);
      kahan_add(acc, t.a + t.b);
    }
  }
  if ((args.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail event: scalar instantiation
    double xa[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) xa[k] = args.cols[k][args.n - 1];
    double wa = 1.0;
    if constexpr (WEIGHTED) wa = args.w[args.n - 1];
    kahan_add(acc, event<M, double, WEIGHTED, GRAD>(cx, c, xa, wa, off, acc));
  }
  block_reduce_and_finish<NACC>(acc, args.partials, args.ticket, args.out, &args.xc);
}

// Weighted-covariance pass (reference: covariance_with_weights,
// src/zfit/minimizers/errors.py:333-464, C = sum_i w_i^2 grad(log p_i) grad(log p_i)^T): besides the usual
// accumulators it sums the outer product of the per-event vector e_i = [c_i0 .. c_i,NS-1, w_i] with
// c_ij = w_i * d l_i / d slot_j (normalisation constants held fixed; the host applies the same linear map as for
// the gradient).  Runs once or twice per fit (metric seed / weighted covariance), not on the step path.
// out: [0,1] value pair, [2, 2+NS) slot sums, then the upper triangle of sum e e^T, row-major, (NS+1)(NS+2)/2 entries.
template <class M, bool WEIGHTED>
__global__ void __launch_bounds__(ZNLL_BLOCK, 1) cov_kernel(const __grid_constant__ KArgs<M::NC> args) {
  constexpr int NS = M::NS, NE = NS + 1, NQ = NE * (NE + 1) / 2, NACC = 2 + NS + NQ;
  __shared__ MathTabs tabs;
  init_tabs(tabs);
  __syncthreads();
  const Ctx cx = ctx_of(tabs);
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  const long long nvec = args.n >> 1;
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x;
  // two events per thread in lock-step like the step kernel; the pair keeps per-lane score vectors e.a / e.b
  EventPair<M, WEIGHTED> cur, nxt;
  if (i < nvec) nxt.load(args, i);
  while (i < nvec) {
    cur = nxt;
    i += stride;
    if (i < nvec) nxt.load(args, i);
    D2 e[2 + NE];
#pragma unroll
    for (int j = 0; j < 2 + NE; ++j) e[j] = zbc<D2>(0.0);
    const D2 t = event<M, D2, WEIGHTED, true>(cx, c, cur.x, cur.w, args.log_offset, e);
    kahan_add(acc, t.a + t.b);
    e[2 + NS] = cur.w;
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[2 + j] += e[2 + j].a + e[2 + j].b;
    int q = 2 + NS;
#pragma unroll
    for (int a = 0; a < NE; ++a)
#pragma unroll
      for (int b = a; b < NE; ++b) {
        acc[q] = fma(e[2 + a].b, e[2 + b].b, fma(e[2 + a].a, e[2 + b].a, acc[q]));
        ++q;
      }
  }
  if ((args.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail event
    double x[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) x[k] = args.cols[k][args.n - 1];
    double w = 1.0;
    if constexpr (WEIGHTED) w = args.w[args.n - 1];
    double e[2 + NE];
#pragma unroll
    for (int j = 0; j < 2 + NE; ++j) e[j] = 0.0;
    kahan_add(acc, event<M, double, WEIGHTED, true>(cx, c, x, w, args.log_offset, e));
    e[2 + NS] = w;
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[2 + j] += e[2 + j];
    int q = 2 + NS;
#pragma unroll
    for (int a = 0; a < NE; ++a)
#pragma unroll
      for (int b = a; b < NE; ++b) {
        acc[q] = fma(e[2 + a], e[2 + b], acc[q]);
        ++q;
      }
  }
  block_reduce_and_finish<NACC>(acc, args.partials, args.ticket, args.out, nullptr);
}

// per-event normalised pdf values (BasePDF.pdf, core/basepdf.py:398-429): out[i] = P(x_i); args.partials = out
template <class M>
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS) pdf_kernel(const __grid_constant__ KArgs<M::NC> args) {
  __shared__ MathTabs tabs;
  init_tabs(tabs);
  __syncthreads();
  const Ctx cx = ctx_of(tabs);
  const double* __restrict__ c = args.c;
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  for (long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x; i < args.n; i += stride) {
    double x[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) x[k] = __ldg(args.cols[k] + i);
    typename M::template Node<double> m;
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
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS)
sumw_kernel(const double* __restrict__ w, long long n, double* partials, unsigned int* ticket, double* out) {
  double acc[2] = {0.0, 0.0};
  const long long stride = (long long)gridDim.x * ZNLL_BLOCK;
  for (long long i = (long long)blockIdx.x * ZNLL_BLOCK + threadIdx.x; i < n; i += stride) kahan_add(acc, w[i]);
  block_reduce_and_finish<2>(acc, partials, ticket, out, nullptr);
}

}  // namespace znll
