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
#ifndef ZNLL_BLOCK
#define ZNLL_BLOCK 256  // threads per CTA (tuning experiments: -DZNLL_BLOCK=192 with three resident CTAs)
#endif
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
// index of (a, b), a <= b, in the row-major upper triangle of an n x n symmetric matrix
__host__ __device__ constexpr int tri_index(int a, int b, int n) { return a * n - a * (a - 1) / 2 + (b - a); }
ZNLL_DEV double zlaneprod(double x) { return x; }
ZNLL_DEV double zlaneprod(D2 x) { return x.a * x.b; }
// a value the compiler must treat as new (blocks common-subexpression hoisting across a branch)
ZNLL_DEV double opaque(double x) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+d"(x));
#endif
  return x;
}
ZNLL_DEV D2 opaque(D2 x) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+d"(x.a), "+d"(x.b));
#endif
  return x;
}
template <class V>
ZNLL_DEV V zbc(double x) {
  return Lanes<V>::make(x, x);
}
// acc += sum over lanes of g*d
ZNLL_DEV void zacc(double& acc, double g, double d) { acc = fma(g, d, acc); }
ZNLL_DEV void zacc(double& acc, D2 g, D2 d) { acc = fma(g.b, d.b, fma(g.a, d.a, acc)); }
ZNLL_DEV void zacc(double& acc, D2 g, double d) { acc = fma(g.a + g.b, d, acc); }
// sum over lanes of l (or w*l) minus the pair's offset
ZNLL_DEV double zsum_minus(double l, double o) { return l - o; }
ZNLL_DEV double zsum_minus(D2 l, double o) { return (l.a + l.b) - o; }
ZNLL_DEV double zdot_minus(double w, double l, double o) { return fma(w, l, -o); }
ZNLL_DEV double zdot_minus(D2 w, D2 l, double o) { return fma(w.b, l.b, fma(w.a, l.a, -o)); }
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
// fp64 elementary functions tuned for this kernel.  The step kernels are bound by the FP64 pipe's issue rate, so the
// aim is the smallest number of DP instructions, then of instructions at all:
//   * both exp and log are table-driven (tables in shared memory, filled per block by init_tabs): ZNLL_TAB_N entries
//     each, so that the remaining polynomial is short -- with 512 entries exp costs 9 DP instructions (degree-4 Taylor)
//     and log 11 (degree-5 log1p), against 17 / 35+ for table-free libdevice-style code;
//   * polynomial coefficients live in the constant bank, so every DFMA takes them as an operand (libdevice
//     materialises each one with two UMOVs);
//   * reciprocal: MUFU.RCP64H seed + two Newton steps (4 DFMA);
//   * log and reciprocal of the same argument share ONE range test (fast_log_rcp), and both lanes of a D2 run
//     interleaved; the libdevice path (|x| >= 700 for exp; non-positive, subnormal, huge, inf or NaN for log / rcp) is
//     taken once per pair, so IEEE semantics for those arguments are libdevice's.
// Accuracy is tested against libm on the GPU (tests/test_math_gpu.py) and, compiled for the host, on the CPU
// (tests/test_hostsim_cpu.py): exp and rcp within 1 ulp, log within 4e-16 relative away from 0.
// ----------------------------------------------------------------------------------------------
#ifndef ZNLL_TAB_BITS
#define ZNLL_TAB_BITS 9  // 9: 512-entry tables; 7: 128-entry tables with polynomials two (log) / one (exp) terms longer
#endif
#define ZNLL_TAB_N (1 << ZNLL_TAB_BITS)

// 2^(j/512), j = 0..511, correctly rounded (generated with mpmath); copied into shared memory by init_tabs
static __constant__ double kExp2Tab[512] = {
    0x1.0000000000000p+0, 0x1.0058c86da1c0ap+0, 0x1.00b1afa5abcbfp+0, 0x1.010ab5b2cbd11p+0,
    0x1.0163da9fb3335p+0, 0x1.01bd1e77170b4p+0, 0x1.02168143b0281p+0, 0x1.027003103b10ep+0,
    0x1.02c9a3e778061p+0, 0x1.032363d42b027p+0, 0x1.037d42e11bbccp+0, 0x1.03d7411915a8ap+0,
    0x1.04315e86e7f85p+0, 0x1.048b9b35659d8p+0, 0x1.04e5f72f654b1p+0, 0x1.0540727fc1762p+0,
    0x1.059b0d3158574p+0, 0x1.05f5c74f0bec2p+0, 0x1.0650a0e3c1f89p+0, 0x1.06ab99fa6407cp+0,
    0x1.0706b29ddf6dep+0, 0x1.0761ead925493p+0, 0x1.07bd42b72a836p+0, 0x1.0818ba42e7d30p+0,
    0x1.0874518759bc8p+0, 0x1.08d0088f8093fp+0, 0x1.092bdf66607e0p+0, 0x1.0987d61701716p+0,
    0x1.09e3ecac6f383p+0, 0x1.0a402331b9715p+0, 0x1.0a9c79b1f3919p+0, 0x1.0af8f03834e52p+0,
    0x1.0b5586cf9890fp+0, 0x1.0bb23d833d93fp+0, 0x1.0c0f145e46c85p+0, 0x1.0c6c0b6bdae53p+0,
    0x1.0cc922b7247f7p+0, 0x1.0d265a4b520bap+0, 0x1.0d83b23395decp+0, 0x1.0de12a7b26300p+0,
    0x1.0e3ec32d3d1a2p+0, 0x1.0e9c7c55189c6p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0f58503328e6dp+0,
    0x1.0fb66affed31bp+0, 0x1.1014a66f951cep+0, 0x1.1073028d7233ep+0, 0x1.10d17f64d9ef1p+0,
    0x1.11301d0125b51p+0, 0x1.118edb6db2dc1p+0, 0x1.11edbab5e2ab6p+0, 0x1.124cbae51a5c8p+0,
    0x1.12abdc06c31ccp+0, 0x1.130b1e264a0e9p+0, 0x1.136a814f204abp+0, 0x1.13ca058cbae1ep+0,
    0x1.1429aaea92de0p+0, 0x1.1489717425438p+0, 0x1.14e95934f312ep+0, 0x1.154962388149ep+0,
    0x1.15a98c8a58e51p+0, 0x1.1609d83606e12p+0, 0x1.166a45471c3c2p+0, 0x1.16cad3c92df73p+0,
    0x1.172b83c7d517bp+0, 0x1.178c554eaea89p+0, 0x1.17ed48695bbc0p+0, 0x1.184e5d23816c9p+0,
    0x1.18af9388c8deap+0, 0x1.1910eba4df41fp+0, 0x1.1972658375d2fp+0, 0x1.19d4013041dc2p+0,
    0x1.1a35beb6fcb75p+0, 0x1.1a979e2363cf8p+0, 0x1.1af99f8138a1cp+0, 0x1.1b5bc2dc40bf0p+0,
    0x1.1bbe084045cd4p+0, 0x1.1c206fb91588fp+0, 0x1.1c82f95281c6bp+0, 0x1.1ce5a51860746p+0,
    0x1.1d4873168b9aap+0, 0x1.1dab6358e15e8p+0, 0x1.1e0e75eb44027p+0, 0x1.1e71aad999e82p+0,
    0x1.1ed5022fcd91dp+0, 0x1.1f387bf9cda38p+0, 0x1.1f9c18438ce4dp+0, 0x1.1fffd7190241ep+0,
    0x1.2063b88628cd6p+0, 0x1.20c7bc96ffc18p+0, 0x1.212be3578a819p+0, 0x1.21902cd3d09b9p+0,
    0x1.21f49917ddc96p+0, 0x1.2259282fc1f27p+0, 0x1.22bdda27912d1p+0, 0x1.2322af0b63bffp+0,
    0x1.2387a6e756238p+0, 0x1.23ecc1c78903ap+0, 0x1.2451ffb82140ap+0, 0x1.24b760c547f15p+0,
    0x1.251ce4fb2a63fp+0, 0x1.25828c65fa1ffp+0, 0x1.25e85711ece75p+0, 0x1.264e450b3cb82p+0,
    0x1.26b4565e27cddp+0, 0x1.271a8b16f0a30p+0, 0x1.2780e341ddf29p+0, 0x1.27e75eeb3ab98p+0,
    0x1.284dfe1f56381p+0, 0x1.28b4c0ea83f36p+0, 0x1.291ba7591bb70p+0, 0x1.2982b17779965p+0,
    0x1.29e9df51fdee1p+0, 0x1.2a5130f50d65cp+0, 0x1.2ab8a66d10f13p+0, 0x1.2b203fc675d1fp+0,
    0x1.2b87fd0dad990p+0, 0x1.2befde4f2e280p+0, 0x1.2c57e39771b2fp+0, 0x1.2cc00cf2f6c18p+0,
    0x1.2d285a6e4030bp+0, 0x1.2d90cc15d5346p+0, 0x1.2df961f641589p+0, 0x1.2e621c1c14833p+0,
    0x1.2ecafa93e2f56p+0, 0x1.2f33fd6a454d2p+0, 0x1.2f9d24abd886bp+0, 0x1.300670653dfe4p+0,
    0x1.306fe0a31b715p+0, 0x1.30d975721b004p+0, 0x1.31432edeeb2fdp+0, 0x1.31ad0cf63eeacp+0,
    0x1.32170fc4cd831p+0, 0x1.3281375752b40p+0, 0x1.32eb83ba8ea32p+0, 0x1.3355f4fb45e20p+0,
    0x1.33c08b26416ffp+0, 0x1.342b46484ebb4p+0, 0x1.3496266e3fa2dp+0, 0x1.35012ba4ea77dp+0,
    0x1.356c55f929ff1p+0, 0x1.35d7a577dd72bp+0, 0x1.36431a2de883bp+0, 0x1.36aeb428335b4p+0,
    0x1.371a7373aa9cbp+0, 0x1.3786581d3f669p+0, 0x1.37f26231e754ap+0, 0x1.385e91be9c811p+0,
    0x1.38cae6d05d866p+0, 0x1.393761742d808p+0, 0x1.39a401b7140efp+0, 0x1.3a10c7a61d55bp+0,
    0x1.3a7db34e59ff7p+0, 0x1.3aeac4bcdf3eap+0, 0x1.3b57fbfec6cf4p+0, 0x1.3bc559212ef89p+0,
    0x1.3c32dc313a8e5p+0, 0x1.3ca0853c10f28p+0, 0x1.3d0e544ede173p+0, 0x1.3d7c4976d27fap+0,
    0x1.3dea64c123422p+0, 0x1.3e58a63b0a09bp+0, 0x1.3ec70df1c5175p+0, 0x1.3f359bf29743fp+0,
    0x1.3fa4504ac801cp+0, 0x1.40132b07a35dfp+0, 0x1.40822c367a024p+0, 0x1.40f153e4a136ap+0,
    0x1.4160a21f72e2ap+0, 0x1.41d016f44d8f5p+0, 0x1.423fb2709468ap+0, 0x1.42af74a1af3f1p+0,
    0x1.431f5d950a897p+0, 0x1.438f6d5817663p+0, 0x1.43ffa3f84b9d4p+0, 0x1.4470018321a1ap+0,
    0x1.44e086061892dp+0, 0x1.4551318eb43ecp+0, 0x1.45c2042a7d232p+0, 0x1.4632fde7006f4p+0,
    0x1.46a41ed1d0057p+0, 0x1.471566f8827d0p+0, 0x1.4786d668b3237p+0, 0x1.47f86d3001fe5p+0,
    0x1.486a2b5c13cd0p+0, 0x1.48dc10fa920a1p+0, 0x1.494e1e192aed2p+0, 0x1.49c052c5916c4p+0,
    0x1.4a32af0d7d3dep+0, 0x1.4aa532feaada6p+0, 0x1.4b17dea6db7d7p+0, 0x1.4b8ab213d5283p+0,
    0x1.4bfdad5362a27p+0, 0x1.4c70d073537cap+0, 0x1.4ce41b817c114p+0, 0x1.4d578e8bb586bp+0,
    0x1.4dcb299fddd0dp+0, 0x1.4e3eeccbd7b2ap+0, 0x1.4eb2d81d8abffp+0, 0x1.4f26eba2e35f0p+0,
    0x1.4f9b2769d2ca7p+0, 0x1.500f8b804f127p+0, 0x1.508417f4531eep+0, 0x1.50f8ccd3deb0dp+0,
    0x1.516daa2cf6642p+0, 0x1.51e2b00da3b14p+0, 0x1.5257de83f4eefp+0, 0x1.52cd359dfd53dp+0,
    0x1.5342b569d4f82p+0, 0x1.53b85df598d78p+0, 0x1.542e2f4f6ad27p+0, 0x1.54a4298571b06p+0,
    0x1.551a4ca5d920fp+0, 0x1.559098bed1bdfp+0, 0x1.56070dde910d2p+0, 0x1.567dac1351819p+0,
    0x1.56f4736b527dap+0, 0x1.576b63f4d854cp+0, 0x1.57e27dbe2c4cfp+0, 0x1.5859c0d59ca07p+0,
    0x1.58d12d497c7fdp+0, 0x1.5948c32824135p+0, 0x1.59c0827ff07ccp+0, 0x1.5a386b5f43d92p+0,
    0x1.5ab07dd485429p+0, 0x1.5b28b9ee20d1ep+0, 0x1.5ba11fba87a03p+0, 0x1.5c19af482fc8fp+0,
    0x1.5c9268a5946b7p+0, 0x1.5d0b4be135accp+0, 0x1.5d84590998b93p+0, 0x1.5dfd902d47c65p+0,
    0x1.5e76f15ad2148p+0, 0x1.5ef07ca0cbf0fp+0, 0x1.5f6a320dceb71p+0, 0x1.5fe411b078d26p+0,
    0x1.605e1b976dc09p+0, 0x1.60d84fd15612ap+0, 0x1.6152ae6cdf6f4p+0, 0x1.61cd3778bc944p+0,
    0x1.6247eb03a5585p+0, 0x1.62c2c91c56acdp+0, 0x1.633dd1d1929fdp+0, 0x1.63b90532205d8p+0,
    0x1.6434634ccc320p+0, 0x1.64afec30678b7p+0, 0x1.652b9febc8fb7p+0, 0x1.65a77e8dcc390p+0,
    0x1.6623882552225p+0, 0x1.669fbcc140be7p+0, 0x1.671c1c70833f6p+0, 0x1.6798a7420a036p+0,
    0x1.68155d44ca973p+0, 0x1.68923e87bfb7ap+0, 0x1.690f4b19e9538p+0, 0x1.698c830a4c8d4p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6a877541ee718p+0, 0x1.6b052fa75173ep+0, 0x1.6b8315a736c75p+0,
    0x1.6c012750bdabfp+0, 0x1.6c7f64b30aa09p+0, 0x1.6cfdcddd47645p+0, 0x1.6d7c62dea2f8ap+0,
    0x1.6dfb23c651a2fp+0, 0x1.6e7a10a38cee8p+0, 0x1.6ef9298593ae5p+0, 0x1.6f786e7ba9fefp+0,
    0x1.6ff7df9519484p+0, 0x1.70777ce1303f6p+0, 0x1.70f7466f42e87p+0, 0x1.71773c4eaa988p+0,
    0x1.71f75e8ec5f74p+0, 0x1.7277ad3ef9011p+0, 0x1.72f8286ead08ap+0, 0x1.7378d02d50b8fp+0,
    0x1.73f9a48a58174p+0, 0x1.747aa5953c849p+0, 0x1.74fbd35d7cbfdp+0, 0x1.757d2df29ce7cp+0,
    0x1.75feb564267c9p+0, 0x1.768069c1a861dp+0, 0x1.77024b1ab6e09p+0, 0x1.7784597eeba8fp+0,
    0x1.780694fde5d3fp+0, 0x1.7888fda749e5dp+0, 0x1.790b938ac1cf6p+0, 0x1.798e56b7fcf03p+0,
    0x1.7a11473eb0187p+0, 0x1.7a94652e958aap+0, 0x1.7b17b0976cfdbp+0, 0x1.7b9b2988fb9ecp+0,
    0x1.7c1ed0130c132p+0, 0x1.7ca2a4456e7a3p+0, 0x1.7d26a62ff86f0p+0, 0x1.7daad5e2850acp+0,
    0x1.7e2f336cf4e62p+0, 0x1.7eb3bedf2e1b9p+0, 0x1.7f3878491c491p+0, 0x1.7fbd5fbab091fp+0,
    0x1.80427543e1a12p+0, 0x1.80c7b8f4abaa9p+0, 0x1.814d2add106d9p+0, 0x1.81d2cb0d1736ap+0,
    0x1.82589994cce13p+0, 0x1.82de968443d9ap+0, 0x1.8364c1eb941f7p+0, 0x1.83eb1bdadb46dp+0,
    0x1.8471a4623c7adp+0, 0x1.84f85b91e07f1p+0, 0x1.857f4179f5b21p+0, 0x1.8606562ab00ecp+0,
    0x1.868d99b4492edp+0, 0x1.87150c27004c2p+0, 0x1.879cad931a436p+0, 0x1.88247e08e1957p+0,
    0x1.88ac7d98a6699p+0, 0x1.8934ac52be8f7p+0, 0x1.89bd0a478580fp+0, 0x1.8a4597875c644p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8b574029db01ep+0, 0x1.8be05bad61778p+0, 0x1.8c69a6bdb5598p+0,
    0x1.8cf3216b5448cp+0, 0x1.8d7ccbc6c19e6p+0, 0x1.8e06a5e0866d9p+0, 0x1.8e90afc931857p+0,
    0x1.8f1ae99157736p+0, 0x1.8fa553499284bp+0, 0x1.902fed0282c8ap+0, 0x1.90bab6ccce12cp+0,
    0x1.9145b0b91ffc6p+0, 0x1.91d0dad829e70p+0, 0x1.925c353aa2fe2p+0, 0x1.92e7bff148396p+0,
    0x1.93737b0cdc5e5p+0, 0x1.93ff669e2802bp+0, 0x1.948b82b5f98e5p+0, 0x1.9517cf65253d1p+0,
    0x1.95a44cbc8520fp+0, 0x1.9630faccf9243p+0, 0x1.96bdd9a7670b3p+0, 0x1.974ae95cba768p+0,
    0x1.97d829fde4e50p+0, 0x1.98659b9bddb5bp+0, 0x1.98f33e47a22a2p+0, 0x1.9981121235681p+0,
    0x1.9a0f170ca07bap+0, 0x1.9a9d4d47f2598p+0, 0x1.9b2bb4d53fe0dp+0, 0x1.9bba4dc5a3dd3p+0,
    0x1.9c49182a3f090p+0, 0x1.9cd81414380f2p+0, 0x1.9d674194bb8d5p+0, 0x1.9df6a0bcfc15ep+0,
    0x1.9e86319e32323p+0, 0x1.9f15f4499c647p+0, 0x1.9fa5e8d07f29ep+0, 0x1.a0360f4424fcbp+0,
    0x1.a0c667b5de565p+0, 0x1.a156f23701b15p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a2789dacfe68cp+0,
    0x1.a309bec4a2d33p+0, 0x1.a39b1231475f7p+0, 0x1.a42c980460ad8p+0, 0x1.a4be504f696b1p+0,
    0x1.a5503b23e255dp+0, 0x1.a5e25893523d4p+0, 0x1.a674a8af46052p+0, 0x1.a7072b8950a73p+0,
    0x1.a799e1330b358p+0, 0x1.a82cc9be14dcap+0, 0x1.a8bfe53c12e59p+0, 0x1.a95333beb0b7ep+0,
    0x1.a9e6b5579fdbfp+0, 0x1.aa7a6a1897fd2p+0, 0x1.ab0e521356ebap+0, 0x1.aba26d59a09eep+0,
    0x1.ac36bbfd3f37ap+0, 0x1.accb3e100301ep+0, 0x1.ad5ff3a3c2774p+0, 0x1.adf4dcca5a413p+0,
    0x1.ae89f995ad3adp+0, 0x1.af1f4a17a4735p+0, 0x1.afb4ce622f2ffp+0, 0x1.b04a868742ee4p+0,
    0x1.b0e07298db666p+0, 0x1.b17692a8fa8cdp+0, 0x1.b20ce6c9a8952p+0, 0x1.b2a36f0cf3f3ap+0,
    0x1.b33a2b84f15fbp+0, 0x1.b3d11c43bbd62p+0, 0x1.b468415b749b1p+0, 0x1.b4ff9ade433c6p+0,
    0x1.b59728de5593ap+0, 0x1.b62eeb6ddfc87p+0, 0x1.b6c6e29f1c52ap+0, 0x1.b75f0e844bfc6p+0,
    0x1.b7f76f2fb5e47p+0, 0x1.b89004b3a7804p+0, 0x1.b928cf22749e4p+0, 0x1.b9c1ce8e77680p+0,
    0x1.ba5b030a1064ap+0, 0x1.baf46ca7a67a7p+0, 0x1.bb8e0b79a6f1fp+0, 0x1.bc27df9285775p+0,
    0x1.bcc1e904bc1d2p+0, 0x1.bd5c27e2cb5e5p+0, 0x1.bdf69c3f3a207p+0, 0x1.be91462c95b60p+0,
    0x1.bf2c25bd71e09p+0, 0x1.bfc73b0468d30p+0, 0x1.c06286141b33dp+0, 0x1.c0fe06ff301f4p+0,
    0x1.c199bdd85529cp+0, 0x1.c235aab23e61ep+0, 0x1.c2d1cd9fa652cp+0, 0x1.c36e26b34e065p+0,
    0x1.c40ab5fffd07ap+0, 0x1.c4a77b9881650p+0, 0x1.c544778fafb22p+0, 0x1.c5e1a9f8630adp+0,
    0x1.c67f12e57d14bp+0, 0x1.c71cb269e601fp+0, 0x1.c7ba88988c933p+0, 0x1.c8589584661a1p+0,
    0x1.c8f6d9406e7b5p+0, 0x1.c99553dfa8313p+0, 0x1.ca3405751c4dbp+0, 0x1.cad2ee13da7cbp+0,
    0x1.cb720dcef9069p+0, 0x1.cc1164b994d23p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cd50b869d8f0fp+0,
    0x1.cdf0b555dc3fap+0, 0x1.ce90e9be12cb9p+0, 0x1.cf3155b5bab74p+0, 0x1.cfd1f95018d17p+0,
    0x1.d072d4a07897cp+0, 0x1.d113e7ba2c38cp+0, 0x1.d1b532b08c968p+0, 0x1.d256b596f948cp+0,
    0x1.d2f87080d89f2p+0, 0x1.d39a638197a3cp+0, 0x1.d43c8eacaa1d6p+0, 0x1.d4def2158a91fp+0,
    0x1.d5818dcfba487p+0, 0x1.d62461eec14bep+0, 0x1.d6c76e862e6d3p+0, 0x1.d76ab3a99745bp+0,
    0x1.d80e316c98398p+0, 0x1.d8b1e7e2d479dp+0, 0x1.d955d71ff6075p+0, 0x1.d9f9ff37adb4ap+0,
    0x1.da9e603db3285p+0, 0x1.db42fa45c4dfdp+0, 0x1.dbe7cd63a8315p+0, 0x1.dc8cd9ab294e4p+0,
    0x1.dd321f301b460p+0, 0x1.ddd79e065807dp+0, 0x1.de7d5641c0658p+0, 0x1.df2347f63c159p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e06fd81a2ece1p+0, 0x1.e11676b197d17p+0, 0x1.e1bd4f11f8220p+0,
    0x1.e264614f5a129p+0, 0x1.e30bad7dcee90p+0, 0x1.e3b333b16ee12p+0, 0x1.e45af3fe592e8p+0,
    0x1.e502ee78b3ff6p+0, 0x1.e5ab2334ac7eep+0, 0x1.e653924676d76p+0, 0x1.e6fc3bc24e350p+0,
    0x1.e7a51fbc74c83p+0, 0x1.e84e3e4933c7ep+0, 0x1.e8f7977cdb740p+0, 0x1.e9a12b6bc3181p+0,
    0x1.ea4afa2a490dap+0, 0x1.eaf503ccd2be5p+0, 0x1.eb9f4867cca6ep+0, 0x1.ec49c80faa594p+0,
    0x1.ecf482d8e67f1p+0, 0x1.ed9f78d802dc2p+0, 0x1.ee4aaa2188510p+0, 0x1.eef616ca06dd6p+0,
    0x1.efa1bee615a27p+0, 0x1.f04da28a52e59p+0, 0x1.f0f9c1cb6412ap+0, 0x1.f1a61cbdf5be7p+0,
    0x1.f252b376bba97p+0, 0x1.f2ff860a70c22p+0, 0x1.f3ac948dd7274p+0, 0x1.f459df15b82acp+0,
    0x1.f50765b6e4540p+0, 0x1.f5b5288633625p+0, 0x1.f6632798844f8p+0, 0x1.f7116302bd526p+0,
    0x1.f7bfdad9cbe14p+0, 0x1.f86e8f32a4b45p+0, 0x1.f91d802243c89p+0, 0x1.f9ccadbdac61dp+0,
    0x1.fa7c1819e90d8p+0, 0x1.fb2bbf4c0ba54p+0, 0x1.fbdba3692d514p+0, 0x1.fc8bc4866e8adp+0,
    0x1.fd3c22b8f71f1p+0, 0x1.fdecbe15f6314p+0, 0x1.fe9d96b2a23d9p+0, 0x1.ff4eaca4391b6p+0,
};
#if ZNLL_TAB_BITS == 9
static __constant__ double kExpT[8] = {
    0x1.71547652b82fep+9,     // 0 512/ln2
    6755399441055744.0,       // 1 1.5 * 2^52
    -0x1.62e42fef00000p-10,   // 2 -(ln2/512)_hi: 33 significant bits, n * hi is exact for |n| < 2^20
    -0x1.473de6af278edp-43,   // 3 -(ln2/512)_lo
    0.0,                      // 4 (1/120: not needed, |r| <= ln2/1024)
    0x1.5555555555555p-5,     // 5 1/24
    0x1.5555555555555p-3,     // 6 1/6
    0.0};
// log1p(r) = r + r^2 * s(r), |r| <= 2^-10: Taylor through r^5 (truncation 1.4e-19)
static __constant__ double kLogC[4] = {-0.5, 0x1.5555555555555p-2, -0.25, 0x1.999999999999ap-3};
#elif ZNLL_TAB_BITS == 7
static __constant__ double kExpT[8] = {
    0x1.71547652b82fep+7,     // 0 128/ln2
    6755399441055744.0,       // 1 1.5 * 2^52
    -0x1.62e42fef80000p-8,    // 2 -(ln2/128)_hi: 35 significant bits
    -0x1.1cf79abc9e3b4p-43,   // 3 -(ln2/128)_lo
    0x1.1111111111111p-7,     // 4 1/120
    0x1.5555555555555p-5,     // 5 1/24
    0x1.5555555555555p-3,     // 6 1/6
    0.0};
// log1p(r) = r + r^2 * s(r), |r| <= 2^-8: minimax fit of degree 5 in s
static __constant__ double kLogC[6] = {-0.5,
                                       0.33333333333333337,
                                       -0.24999999998288297,
                                       0.19999999998478485,
                                       -0.16666959217649738,
                                       0.14285974331115564};
#else
#error "ZNLL_TAB_BITS must be 7 or 9"
#endif
static __constant__ double kMathC[4] = {
    0x1.62e42fefa39efp-1,  // 0 ln2 (its rounding error, 5.5e-17 relative, is relative to k ln2 and so to the result)
    0.0,
    4503601774854144.0,    // 2 2^52 + 2^31
    0.0};

// per-block tables of the elementary functions, in shared memory
struct MathTabs {
  double2 log[ZNLL_TAB_N];  // (1/c_i, log c_i) for ZNLL_TAB_N intervals of [0.6875, 1.375)
  double exp2[ZNLL_TAB_N];  // 2^(j/ZNLL_TAB_N)
};
#define ZNLL_HAS_OPTIMISTIC 1  // (tests/hostsim keys on this to drive the pair loop the way nll_kernel does)
#define ZNLL_HAS_SPLIT 1       // (... and on this for the running-product form of the pair loop)
// Evaluation context of one thread: where the tables are, and HOW out-of-range arguments are handled.
//   careful = true : every elementary function tests its argument and answers an out-of-range one (|x| >= 700 for exp;
//                    non-positive, subnormal, huge, inf or NaN for log / reciprocal) through libdevice on the spot --
//                    IEEE semantics, one branch per call site.
//   careful = false: the OPTIMISTIC sweep of the step kernels.  The table-driven cores run unconditionally and each
//                    call site only ORs its range test into `bad` (predicate logic, no branch, no reconvergence
//                    barrier); event() looks at `bad` ONCE after the forward sweep and, if it is set, evaluates the
//                    pair again with a careful context.  A pair that takes no rare path thereby crosses a single test.
// `careful` is a compile-time constant wherever a Ctx is built (everything is inlined), so the mode costs nothing.
struct Ctx {
#if defined(__CUDA_ARCH__)
  // byte addresses of the tables in the shared window, held in ordinary registers: naming the __shared__ object at
  // every look-up makes ptxas rebuild its cluster-window address each time (S2UR CgaCtaId + UMOV + ULEA per site)
  unsigned logtab, exptab;
#else
  const double2* logtab;
  const double* exptab;
#endif
  bool careful;
  // optimistic sweep: running maxima of the range-test operands, one per KIND of test; any_flag() compares each with its
  // bound once per pair.  (An accumulated 0/1 flag cost a compare, a select and an OR per call site.)
  mutable unsigned chk_abs;   // |x| < 600        : high word with the sign cleared            (exp arguments, l0 - l1)
  mutable unsigned chk_mult;  // 2^-200 <= x < 2^200 : high word - 0x33700000 (factors of the running product)
  mutable unsigned chk_pos;   // 2^-958 <= x < 2^1022: high word - 0x04100000 (arguments of log / reciprocal)
  mutable unsigned chk_add;   // x >= -600         : high word as it is                         (log-density terms)
};
ZNLL_DEV double tab_exp2(const Ctx& cx, int n) {  // 2^((n mod N)/N)
#if defined(__CUDA_ARCH__)
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(cx.exptab + (((unsigned)n & (ZNLL_TAB_N - 1)) << 3)));
  return v;
#else
  return cx.exptab[n & (ZNLL_TAB_N - 1)];
#endif
}
ZNLL_DEV double2 tab_log(const Ctx& cx, int j) {  // (1/c_j, log c_j), j already reduced to [0, N)
#if defined(__CUDA_ARCH__)
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(cx.logtab + ((unsigned)j << 4)));
  return v;
#else
  return cx.logtab[j];
#endif
}
ZNLL_DEV unsigned umax2(unsigned a, unsigned b) { return a > b ? a : b; }
ZNLL_DEV unsigned umax3(unsigned a, unsigned b, unsigned c) { return umax2(a, umax2(b, c)); }
ZNLL_DEV unsigned hi_of(double x) { return (unsigned)__double2hiint(x); }
ZNLL_DEV void reset_flags(const Ctx& cx) { cx.chk_abs = cx.chk_mult = cx.chk_pos = cx.chk_add = 0u; }
ZNLL_DEV bool any_flag(const Ctx& cx) {
  return (cx.chk_abs >= 0x4082C000u) | (cx.chk_mult >= 0x19000000u) | (cx.chk_pos >= 0x7bc00000u) | (cx.chk_add > 0xC082C000u);
}

template <class V>
ZNLL_DEV void flag_abs600(const Ctx& cx, V x) {
  cx.chk_abs = umax3(cx.chk_abs, hi_of(la(x)) & 0x7fffffffu, hi_of(lb(x)) & 0x7fffffffu);
}
template <class V>
ZNLL_DEV void flag_mult(const Ctx& cx, V x) {
  cx.chk_mult = umax3(cx.chk_mult, hi_of(la(x)) - 0x33700000u, hi_of(lb(x)) - 0x33700000u);
}
template <class V>
ZNLL_DEV void flag_pos(const Ctx& cx, V x) {
  cx.chk_pos = umax3(cx.chk_pos, hi_of(la(x)) - 0x04100000u, hi_of(lb(x)) - 0x04100000u);
}
template <class V>
ZNLL_DEV void flag_add(const Ctx& cx, V x) {
  cx.chk_add = umax3(cx.chk_add, hi_of(la(x)), hi_of(lb(x)));
}
ZNLL_DEV bool exp_in_range(double x) { return (__double2hiint(x) & 0x7fffffff) < 0x4085E000; }  // |x| < 700
ZNLL_DEV bool log_in_range(double x) { return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fc00000u; }  // normal, < 2^1022
// x < -690 read off the high word (the sign bit set and the magnitude above 690; -inf and negative NaNs included)
ZNLL_DEV bool below_m690(double x) { return (unsigned)__double2hiint(x) > 0xC0859000u; }
ZNLL_DEV bool rcp_in_range(double x) { return (((unsigned)__double2hiint(x) >> 20) & 0x7ffu) - 1u < 0x7fdu; }
ZNLL_DEV double scale2n(double p, int n) { return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p)); }
// p * 2^(n >> ZNLL_TAB_BITS) from the exp core's rounded integer n: one shift, one multiply-add on the high word
ZNLL_DEV double scale2n_tab(double p, int n) {
#if defined(__CUDA_ARCH__)
  int hi;
  asm("{ .reg .s32 k; shr.s32 k, %1, %2; mad.lo.s32 %0, k, 1048576, %3; }" : "=r"(hi) : "r"(n), "n"(ZNLL_TAB_BITS), "r"(__double2hiint(p)));
  return __hiloint2double(hi, __double2loint(p));
#else
  return scale2n(p, n >> ZNLL_TAB_BITS);
#endif
}
ZNLL_DEV double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

// exp(x) = 2^k * 2^(j/N) * e^r with n = N k + j = nearest(x * N/ln2) and |r| <= ln2/(2N): a table look-up and a short
// Taylor polynomial (truncation below 1.3e-18).  Total error about 1 ulp (table 0.5 + final FMA 0.5).
template <class V>
ZNLL_DEV V exp_core(V x, const Ctx& cx) {  // |x| < 700
  const V t = zfma(x, kExpT[0], kExpT[1]);
  const int na = __double2loint(la(t)), nb = __double2loint(lb(t));
  const V T = Lanes<V>::make(tab_exp2(cx, na), tab_exp2(cx, nb));
  const V nd = zsub(t, kExpT[1]);
  V r = zfma(nd, kExpT[2], x);
  r = zfma(nd, kExpT[3], r);
#if ZNLL_TAB_BITS == 9
  V q = zfma(r, kExpT[5], kExpT[6]);
#else
  V q = zfma(r, kExpT[4], kExpT[5]);
  q = zfma(q, r, kExpT[6]);
#endif
  q = zfma(q, r, 0.5);
  const V p = zfma(zmul(r, q), r, r);  // e^r - 1 = r + (r q) r: the DFMA reads two distinct registers, not three
  const V v = zfma(T, p, T);
  return Lanes<V>::make(scale2n_tab(la(v), na), scale2n_tab(lb(v), nb));
}
template <class V>
ZNLL_DEV V fast_exp(V x, const Ctx& cx) {
  if (!cx.careful)
    flag_abs600(cx, x);
  else if (!(exp_in_range(la(x)) && exp_in_range(lb(x))))
    return Lanes<V>::make(exp(la(x)), exp(lb(x)));  // rare
  return exp_core(x, cx);
}
// exp(-|d|): the range test reads d itself, so -|d| never has to exist in a register (the DFMAs of the core take it
// through their operand modifiers)
template <class V>
ZNLL_DEV V fast_exp_negabs(V d, const Ctx& cx) {
  if (!cx.careful)
    flag_abs600(cx, d);
  else if (!(exp_in_range(la(d)) && exp_in_range(lb(d))))
    return Lanes<V>::make(exp(-fabs(la(d))), exp(-fabs(lb(d))));  // rare
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
// reciprocals of the two lanes of a pair from ONE reciprocal (Montgomery's trick): 1/a = b * 1/(ab), 1/b = a * 1/(ab).
// `ab` is the lanes' product, which the split evaluation forms anyway for the running product; 6 DP instructions and one
// MUFU instead of 8 and two.  Callers guarantee a, b in [2^-200, 2^200), so ab and its reciprocal are normal.
ZNLL_DEV double rcp_lanes(double x, double) { return rcp_core(x); }
ZNLL_DEV D2 rcp_lanes(D2 x, double ab) {
  const double inv = rcp_core(ab);
  D2 r;
  r.a = x.b * inv;
  r.b = x.a * inv;
  return r;
}
template <class V>
ZNLL_DEV V fast_rcp(V x) {  // always careful (not on the step kernels' hot path)
  if (!(rcp_in_range(la(x)) && rcp_in_range(lb(x)))) return Lanes<V>::make(1.0 / la(x), 1.0 / lb(x));  // rare
  return rcp_core(x);
}

// log(x) = k ln2 + log c + log1p(r), x = 2^k z, z in [0.6875, 1.375), c the midpoint of z's table interval, r = z/c - 1.
// r is formed as x * (2^-k / c) - 1: the power of two goes into the exponent of the table's 1/c (one integer subtract on
// its high word), so z itself never has to be assembled in a register pair.  2^-k / c stays normal for x < 2^1022, which
// is what the range tests of the callers admit.
struct LogSplit {
  double sinvc, logc, kd;  // 2^-k / c, log c, k + 2^52 + 2^31
};
ZNLL_DEV LogSplit log_split(double x, const Ctx& cx) {
  LogSplit s;
  const int tmp = __double2hiint(x) - 0x3fe60000;
  const double2 tc = tab_log(cx, (tmp >> (20 - ZNLL_TAB_BITS)) & (ZNLL_TAB_N - 1));
  s.sinvc = __hiloint2double(__double2hiint(tc.x) - (tmp & 0xfff00000), __double2loint(tc.x));
  s.logc = tc.y;
  s.kd = __hiloint2double(0x43300000, (tmp >> 20) ^ 0x80000000);
  return s;
}
template <class V>
ZNLL_DEV V log_core(V x, const Ctx& cx) {  // x positive, normal, below 2^1022
  const LogSplit sa = log_split(la(x), cx), sb = log_split(lb(x), cx);
  const V sinvc = Lanes<V>::make(sa.sinvc, sb.sinvc), logc = Lanes<V>::make(sa.logc, sb.logc);
  const V kd = zsub(Lanes<V>::make(sa.kd, sb.kd), kMathC[2]);
  const V r = zfma(x, sinvc, -1.0);
#if ZNLL_TAB_BITS == 9
  V s = zfma(r, kLogC[3], kLogC[2]);
#pragma unroll
  for (int j = 1; j >= 0; --j) s = zfma(s, r, kLogC[j]);
#else
  V s = zfma(r, kLogC[5], kLogC[4]);
#pragma unroll
  for (int j = 3; j >= 0; --j) s = zfma(s, r, kLogC[j]);
#endif
  const V w = zfma(kd, kMathC[0], logc);
  return zadd(w, zfma(zmul(r, s), r, r));  // log1p(r) = r + (r s) r (two distinct register operands)
}
template <class V>
ZNLL_DEV V fast_log(V x, const Ctx& cx) {
  if (!cx.careful)
    flag_pos(cx, x);
  else if (!(log_in_range(la(x)) && log_in_range(lb(x))))
    return Lanes<V>::make(log(la(x)), log(lb(x)));  // rare
  return log_core(x, cx);
}

// log and reciprocal of the SAME argument behind ONE range test.  Fast path: x positive and normal with a normal
// reciprocal (biased exponent 1..2045); everything else -- zero, negative, subnormal, huge, inf, NaN -- is answered by
// libdevice once per pair, so the IEEE results (log of a negative number is NaN, 1/0 is inf) are unchanged.
ZNLL_DEV bool pos_normal(double x) { return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fc00000u; }  // [2^-1022, 2^1022)
template <class V>
ZNLL_DEV void fast_log_rcp(V x, const Ctx& cx, V& lg, V& rc) {
  if (!cx.careful) {
    flag_pos(cx, x);
  } else if (!(pos_normal(la(x)) && pos_normal(lb(x)))) {  // rare
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
  if (!cx.careful) {
    flag_pos(cx, x);  // a negative x is flagged as well: the optimistic sweep never sees a sign
  } else if (!(pos_normal(la(x)) && pos_normal(lb(x)))) {  // rare
    lg = Lanes<V>::make(log(fabs(la(x))), log(fabs(lb(x))));
    rc = Lanes<V>::make(1.0 / la(x), 1.0 / lb(x));
    neg = zlt(x, 0.0);
    return;
  }
  lg = log_core(x, cx);
  rc = rcp_core(x);
  neg = Lanes<V>::mask(false, false);
}
// log(p + 1e-307) and its derivative 1/(p + 1e-307) (loss.py:117).  For p >= 2^-958 the sum IS p, bit for bit
// (1e-307 / p < 2^-54), so the common path neither adds nor tests twice: one test for "p positive, at least 2^-958, and
// 1/p normal"; a tiny, zero, negative or non-finite p takes the literal formula through libdevice.
ZNLL_DEV bool pos_not_tiny(double x) { return (unsigned)(__double2hiint(x) - 0x04100000) < 0x7bc00000u; }  // [2^-958, 2^1022)
template <class V, bool WANT_RCP>
ZNLL_DEV void log_eps_rcp(V p, const Ctx& cx, V& lg, V& rc) {
  if (!cx.careful) {
    flag_pos(cx, p);
  } else if (!(pos_not_tiny(la(p)) && pos_not_tiny(lb(p)))) {  // rare
    const double pa = la(p) + ZNLL_LOG_EPS, pb = lb(p) + ZNLL_LOG_EPS;
    lg = Lanes<V>::make(log(pa), log(pb));
    rc = Lanes<V>::make(1.0 / pa, 1.0 / pb);
    return;
  }
  lg = log_core(p, cx);
  if (WANT_RCP) rc = rcp_core(p);
}

// a factor of the running product of the split evaluation: positive, within [2^-200, 2^200)
ZNLL_DEV bool mult_in_range(double x) { return (unsigned)(__double2hiint(x) - 0x33700000) < 0x19000000u; }
// the running product itself: positive, within [2^-512, 2^512) -- one more factor pair (each within 2^+-400) cannot overflow it
ZNLL_DEV bool prod_in_range(double x) { return (unsigned)(__double2hiint(x) - 0x1FF00000) < 0x40000000u; }
ZNLL_DEV bool below_m600(double x) { return (unsigned)__double2hiint(x) > 0xC082C000u; }  // x < -600 (incl. -inf, negative NaN)

// every thread fills its share of the tables, then __syncthreads() by the caller
ZNLL_DEV void init_tabs(MathTabs& tabs) {
  for (int j = threadIdx.x; j < ZNLL_TAB_N; j += ZNLL_BLOCK) {
    const int ha = 0x3fe60000 + (j << (20 - ZNLL_TAB_BITS));
    const double a = __hiloint2double(ha, 0), b = __hiloint2double(ha + (1 << (20 - ZNLL_TAB_BITS)), 0);
    const double invc = 1.0 / (0.5 * (a + b));
    tabs.log[j] = make_double2(invc, -log(invc));
    tabs.exp2[j] = kExp2Tab[j << (9 - ZNLL_TAB_BITS)];
  }
}
// `careful` selects the mode (see Ctx).  On the device the table addresses come out of a volatile asm statement, placed
// after the block barrier that follows init_tabs: ptxas can then neither rematerialise them at every look-up nor move a
// table load above the barrier (every load depends on them).
ZNLL_DEV Ctx ctx_of(const MathTabs& tabs, bool careful = true) {
  Ctx cx;
#if defined(__CUDA_ARCH__)
  unsigned base;
  asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(base) : "l"(&tabs));
  cx.logtab = base;
  cx.exptab = base + (unsigned)(sizeof(double2) * ZNLL_TAB_N);
#else
  cx.logtab = tabs.log;
  cx.exptab = tabs.exp2;
#endif
  cx.careful = careful;
  reset_flags(cx);
  return cx;
}
ZNLL_DEV Ctx with_mode(const Ctx& cx, bool careful) {
  Ctx r = cx;
  r.careful = careful;
  reset_flags(r);
  return r;
}
// A tree that is evaluated in log space from leaves that need neither exp nor log (a Gauss, an Exponential, products of
// them: cfg 1) never touches the tables; its kernels skip filling them (the rare p + 1e-307 path uses libdevice).
template <class M>
struct NeedTabs {
  static constexpr bool value = !M::LOGNAT || M::TABS;
};

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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  // Hessian pass: h += r * (d2 u / d slot_a d slot_b) / u over the upper triangle (mu mu, mu sigma, sigma sigma) of the
  // leaf's slots, u the leaf's unnormalised density: ((t^2-1), (t^3-3t), (t^4-5t^2+2)) / sigma^2
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const V t2 = zmul(t, t);
    const V rs = zmul(r, c[CO + 0] * c[CO + 0]);
    zacc(h[0], rs, zadd(t2, -1.0));
    zacc(h[1], rs, zmul(t, zadd(t2, -3.0)));
    zacc(h[2], rs, zfma(t2, zadd(t2, -5.0), 2.0));
  }
};
template <int COL>
struct Gauss {
  static constexpr int NS = 2, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;  // HESS: the exact one-pass Hessian knows this node
  static constexpr int ROOT = 0;
  static constexpr int NH = 3;                     // second-order accumulators: upper triangle over the node's slots
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = false;  // does log-space evaluation look up the exp / log tables?
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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  // Hessian pass: the Gauss block on (mu, sigma) for events inside [low, high]; low / high only move the normalisation.
  // h: upper triangle over the four slots.
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const V t2 = zmul(t, t);
    const V rs = zsel(out, zbc<V>(0.0), zmul(r, c[CO + 0] * c[CO + 0]));
    zacc(h[tri_index(0, 0, 4)], rs, zadd(t2, -1.0));
    zacc(h[tri_index(0, 1, 4)], rs, zmul(t, zadd(t2, -3.0)));
    zacc(h[tri_index(1, 1, 4)], rs, zfma(t2, zadd(t2, -5.0), 2.0));
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct TGauss {
  static constexpr int NS = 4, NC = 5, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 10;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = false;
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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__, double* h) const {  // u = exp(lambda xs): u'' / u = xs^2
    zacc(h[0], r, zmul(xs, xs));
  }
};
template <int COL>
struct Expo {
  static constexpr int NS = 1, NC = 3, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 1;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = false;
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
  V t, mdt, lu, P;  // mdt = -d ln v / d t (t itself on the Gaussian core); lu = ln(B-t) on the tail
  Mask tail;
  bool any_tail;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask&) {
    t = zmul(zsub(x[COL], c[CO + 0]), c[CO + 1]);
    // t < -|alpha|  <=>  t + |alpha| < 0, and that sum has the exact sign (it is exact when the two are within a factor
    // of two of each other, far from zero otherwise; t == -|alpha| gives +0): integer sign tests replace three DSETPs
    const V ta = zadd(t, c[CO + 3]);
    any_tail = zany_signbit(ta);
    const V lp_core = zfma(zmul(t, -0.5), t, c[CO + 7]);  // exp(-t^2/2)
    if (!any_tail) return lp_core;  // whole pair on the Gaussian core (events are bucketed): mdt, lu, tail stay unset
    tail = zsignbit(ta);
    // A*(B-t)^-n == exp(ln A - n*ln(B-t)); a lane on the core gets the safe value u = 2 (z.safe_where)
    const V u = zsel(tail, zsub(c[CO + 5], t), zbc<V>(2.0));
    V ru;
    fast_log_rcp(u, cx, lu, ru);  // one range test for both
    const V mdtt = zmul(ru, -c[CO + 4]);  // -n/u
    const V lp_tail = zfma(lu, -c[CO + 4], c[CO + 6]);
    mdt = zsel(tail, mdtt, t);
    return zsel(tail, lp_tail, lp_core);
  }
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg;
    P = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    return P;
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_log(V g, const double* __restrict__ c, A* acc) {
    if (!any_tail) {  // core pair: -d ln v/dt is t itself, nothing flows to alpha and n
      const V gm = zmul(g, t);
      zacc(acc[AO + 0], gm, c[CO + 1]);            // dt/dmu    = -sign/sigma
      zacc(acc[AO + 1], zmul(gm, t), c[CO + 2]);   // dt/dsigma = -t/sigma
      return;
    }
    const V gm = zmul(g, mdt);                   // -g * d ln v/dt
    zacc(acc[AO + 0], gm, c[CO + 1]);
    zacc(acc[AO + 1], zmul(gm, t), c[CO + 2]);
    // d ln v/d|alpha| = -n/a - a + (n/u)(n/a^2+1) ; d ln v/dn = ln(n/a)+1 - ln u - (n/u)/a ; 0 on the core
    const V gt = zsel(tail, g, zbc<V>(0.0));
    zacc(acc[AO + 2], zmul(gt, c[CO + 12]), zfma(mdt, -c[CO + 11], c[CO + 10]));
    zacc(acc[AO + 3], gt, zfma(mdt, c[CO + 9], zsub(c[CO + 8], lu)));
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
  // Hessian pass: r * (d2 u / da db) / u = r * (d2 ln u + d ln u (x) d ln u) over the upper triangle of (mu, sigma, alpha, n),
  // row-major: [mm, ms, ma, mn, ss, sa, sn, aa, an, nn].  ln u = phi(t, a, n) with a = |alpha|: -t^2/2 on the core,
  // n ln(n/a) - a^2/2 - n ln(B - t) on the tail (B = n/a - a); t = (x - mu) sign(alpha) / sigma.  Scalar per lane: this
  // pass runs once or twice per fit.
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const double sos = c[CO + 1], is = c[CO + 2], a = c[CO + 3], n = c[CO + 4], sg = c[CO + 12];
    const double Ba = -c[CO + 11], Bn = c[CO + 9];
    const double rr[2] = {la(r), lb(r)}, tt[2] = {la(t), lb(t)};
    bool tl[2] = {false, false};
    double md[2] = {0.0, 0.0}, lv[2] = {0.0, 0.0};
    if (any_tail) {
      tl[0] = la(tail);
      tl[1] = lb(tail);
      md[0] = la(mdt);
      md[1] = lb(mdt);
      lv[0] = la(lu);
      lv[1] = lb(lu);
    }
#pragma unroll
    for (int k = 0; k < Lanes<V>::N; ++k) {
      const double tau = tt[k];
      double pt, ptt, pta = 0.0, ptn = 0.0, paa = 0.0, pan = 0.0, pnn = 0.0, da = 0.0, dn = 0.0;
      if (tl[k]) {
        const double iv = -md[k] / n;  // 1 / (B - t)
        pt = n * iv;
        ptt = n * iv * iv;
        pta = -ptt * Ba;
        ptn = iv - ptt * Bn;
        da = c[CO + 10] + pt * c[CO + 11];
        dn = c[CO + 8] - lv[k] - pt * Bn;
        paa = n / (a * a) - 1.0 + ptt * Ba * Ba - pt * (2.0 * n / (a * a * a));
        pan = -1.0 / a - ptn * Ba + pt / (a * a);
        pnn = 1.0 / n - 2.0 * Bn * iv + ptt * Bn * Bn;
      } else {
        pt = -tau;
        ptt = -1.0;
      }
      const double tm = -sos, ts = -tau * is, tms = sos * is, tss = 2.0 * tau * is * is;
      const double dm = pt * tm, ds = pt * ts;
      const double w = rr[k];
      h[0] = fma(w, ptt * tm * tm + dm * dm, h[0]);
      h[1] = fma(w, ptt * tm * ts + pt * tms + dm * ds, h[1]);
      h[2] = fma(w, sg * (pta * tm + dm * da), h[2]);
      h[3] = fma(w, ptn * tm + dm * dn, h[3]);
      h[4] = fma(w, ptt * ts * ts + pt * tss + ds * ds, h[4]);
      h[5] = fma(w, sg * (pta * ts + ds * da), h[5]);
      h[6] = fma(w, ptn * ts + ds * dn, h[6]);
      h[7] = fma(w, paa + da * da, h[7]);
      h[8] = fma(w, sg * (pan + da * dn), h[8]);
      h[9] = fma(w, pnn + dn * dn, h[9]);
    }
  }
};
template <int COL>
struct CBall {
  static constexpr int NS = 4, NC = 13, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 10;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = true;
  template <class V>
  using Node = CBallN<COL, V>;
};

// Second-order leaf primitive of the piecewise "Gaussian core + tail" shapes, one lane: with ln u = phi(t, a[, n]),
// t = (x - mu) s / sigma, returns U_xy = d2 ln u / dx dy + (d ln u / dx)(d ln u / dy) over (mu, sigma, a[, n]) as
// [mm, ms, ma, mn, ss, sa, sn, aa, an, nn] (n entries zero for the exponential tail).  sos = s / sigma, is = 1 / sigma.
struct TailDerivs {
  double pt, ptt, pta, ptn, paa, pan, pnn, da, dn;  // partial derivatives of phi; da, dn = d phi / da, d phi / dn
};
ZNLL_DEV void tail_block(const TailDerivs& d, double tau, double sos, double is, double* U) {
  const double tm = -sos, ts = -tau * is, tms = sos * is, tss = 2.0 * tau * is * is;
  const double dm = d.pt * tm, ds = d.pt * ts;
  U[0] = d.ptt * tm * tm + dm * dm;
  U[1] = d.ptt * tm * ts + d.pt * tms + dm * ds;
  U[2] = d.pta * tm + dm * d.da;
  U[3] = d.ptn * tm + dm * d.dn;
  U[4] = d.ptt * ts * ts + d.pt * tss + ds * ds;
  U[5] = d.pta * ts + ds * d.da;
  U[6] = d.ptn * ts + ds * d.dn;
  U[7] = d.paa + d.da * d.da;
  U[8] = d.pan + d.da * d.dn;
  U[9] = d.pnn + d.dn * d.dn;
}
// power-law tail of the CrystalBall family: phi = n ln(n/a) - a^2/2 - n ln(B - t), B = n/a - a; iv = 1/(B - t), lv = ln(B - t)
ZNLL_DEV TailDerivs cb_tail_derivs(double a, double n, double iv, double lv) {
  TailDerivs d;
  const double Ba = -(n / (a * a) + 1.0), Bn = 1.0 / a;
  d.pt = n * iv;
  d.ptt = n * iv * iv;
  d.pta = -d.ptt * Ba;
  d.ptn = iv - d.ptt * Bn;
  d.da = -n / a - a - d.pt * Ba;
  d.dn = log(n / a) + 1.0 - lv - d.pt * Bn;
  d.paa = n / (a * a) - 1.0 + d.ptt * Ba * Ba - d.pt * (2.0 * n / (a * a * a));
  d.pan = -1.0 / a - d.ptn * Ba + d.pt / (a * a);
  d.pnn = 1.0 / n - 2.0 * Bn * iv + d.ptt * Bn * Bn;
  return d;
}
ZNLL_DEV TailDerivs core_derivs(double tau) {
  TailDerivs d = {-tau, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  return d;
}
// exponential tail of the GaussExpTail family: phi = a^2/2 + a t
ZNLL_DEV TailDerivs exp_tail_derivs(double a, double tau) {
  TailDerivs d = {a, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, a + tau, 0.0};
  return d;
}

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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  // Hessian pass: upper triangle over (mu, sigma, alpha)
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const double rr[2] = {la(r), lb(r)}, tt[2] = {la(t), lb(t)};
    const bool tl[2] = {la(tail), lb(tail)};
    const double sg = c[CO + 6];
#pragma unroll
    for (int k = 0; k < Lanes<V>::N; ++k) {
      double U[10];
      tail_block(tl[k] ? exp_tail_derivs(c[CO + 3], tt[k]) : core_derivs(tt[k]), tt[k], c[CO + 1], c[CO + 2], U);
      h[tri_index(0, 0, 3)] = fma(rr[k], U[0], h[tri_index(0, 0, 3)]);
      h[tri_index(0, 1, 3)] = fma(rr[k], U[1], h[tri_index(0, 1, 3)]);
      h[tri_index(0, 2, 3)] = fma(rr[k], sg * U[2], h[tri_index(0, 2, 3)]);
      h[tri_index(1, 1, 3)] = fma(rr[k], U[4], h[tri_index(1, 1, 3)]);
      h[tri_index(1, 2, 3)] = fma(rr[k], sg * U[5], h[tri_index(1, 2, 3)]);
      h[tri_index(2, 2, 3)] = fma(rr[k], U[7], h[tri_index(2, 2, 3)]);
    }
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct GETail {
  static constexpr int NS = 3, NC = 7, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 6;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = false;
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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  // Hessian pass: upper triangle over (mu, sigmal, alphal, sigmar, alphar); an event touches mu and its own side's slots
  template <int IS, int IA>
  ZNLL_DEV static void scatter3(double w, double sg, const double* U, double* h) {
    h[tri_index(0, 0, 5)] = fma(w, U[0], h[tri_index(0, 0, 5)]);
    h[tri_index(0, IS, 5)] = fma(w, U[1], h[tri_index(0, IS, 5)]);
    h[tri_index(0, IA, 5)] = fma(w, sg * U[2], h[tri_index(0, IA, 5)]);
    h[tri_index(IS, IS, 5)] = fma(w, U[4], h[tri_index(IS, IS, 5)]);
    h[tri_index(IS, IA, 5)] = fma(w, sg * U[5], h[tri_index(IS, IA, 5)]);
    h[tri_index(IA, IA, 5)] = fma(w, U[7], h[tri_index(IA, IA, 5)]);
  }
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const double* cc = c + CO;
    const double rr[2] = {la(r), lb(r)}, tt[2] = {la(t), lb(t)};
    const bool tl[2] = {la(tail), lb(tail)}, lf[2] = {la(left), lb(left)};
#pragma unroll
    for (int k = 0; k < Lanes<V>::N; ++k) {
      const double* q = cc + (lf[k] ? 2 : 7);
      double U[10];
      tail_block(tl[k] ? exp_tail_derivs(q[2], tt[k]) : core_derivs(tt[k]), tt[k], q[0], q[1], U);
      if (lf[k])
        scatter3<1, 2>(rr[k], q[4], U, h);
      else
        scatter3<3, 4>(rr[k], q[4], U, h);
    }
  }
  template <int CO, int AO, class A>
  ZNLL_DEV void bwd_val(V a, const double* __restrict__ c, A* acc) {
    bwd_log<CO, AO>(zmul(a, P), c, acc);
  }
};
template <int COL>
struct GGETail {
  static constexpr int NS = 5, NC = 12, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 15;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = false;
  template <class V>
  using Node = GGETailN<COL, V>;
};

// Orthogonal-polynomial sums of degree DEG (models/polynomials.py:46-175): slots [c_0 .. c_DEG]
// consts [0 2/(hi-lo), 1 -(lo+hi)/(hi-lo), 2 1/I, 3.. c_k/I]
// KIND 3: Bernstein basis B_{k,DEG}(t), t = (x - lo)/(hi - lo) (rescale_zero_one + de_casteljau, :872-905).
// KIND 0: Chebyshev T_n (T_{n+1} = 2zT_n - T_{n-1}, polynomials.py:338-343); 1: Legendre
// ((n+1)P_{n+1} = (2n+1)zP_n - nP_{n-1}, :181-190); 2: Chebyshev of the second kind U_n (U_1 = 2z, :487-560);
// 4: Hermite, physicists' (H_1 = 2z, H_{n+1} = 2(zH_n - nH_{n-1}), :752-764); 5: Laguerre (L_1 = 1 - z,
// (n+1)L_{n+1} = (2n+1-z)L_n - nL_{n-1}, :600-633)
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
      if constexpr (KIND == 5)
        T[1] = zfma(z, -1.0, 1.0);
      else
        T[1] = (KIND == 2 || KIND == 4) ? z2 : z;
      p = zfma(T[1], c[CO + 4], p);
    }
#pragma unroll
    for (int k = 2; k <= DEG; ++k) {
      if constexpr (KIND == 1) {
        const double n = (double)(k - 1);
        T[k] = zsub(zmul(zmul(z, T[k - 1]), (2.0 * n + 1.0) / (n + 1.0)), zmul(T[k - 2], n / (n + 1.0)));
      } else if constexpr (KIND == 4) {
        T[k] = zfma(z2, T[k - 1], zmul(T[k - 2], -2.0 * (double)(k - 1)));
      } else if constexpr (KIND == 5) {
        const double n = (double)(k - 1);
        T[k] = zfma(zfma(z, -1.0 / (n + 1.0), (2.0 * n + 1.0) / (n + 1.0)), T[k - 1], zmul(T[k - 2], -n / (n + 1.0)));
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
  // split evaluation (optimistic sweep only): the value goes into the pair's running product instead of through a log
  template <int CO>
  ZNLL_DEV void fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V& mult) {
    const V p = fwd_val<CO>(cx, c, x);
    flag_mult(cx, p);
    rP = rcp_lanes(p, zlaneprod(p));
    mult = zmul(mult, p);
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
  template <int CO>
  ZNLL_DEV void second(V, const double* __restrict__, double*) const {}  // u is linear in the coefficients: u'' = 0
};
template <int COL, int DEG, int KIND>
struct Poly {
  static constexpr int NS = DEG + 1, NC = 3 + DEG + 1, COLMASK = 1 << COL;
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 0;
  static constexpr bool LOGNAT = false;
  static constexpr bool HAS_ADD = false, HAS_MULT = true;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = true;
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
using Herm = Poly<COL, DEG, 4>;
template <int COL, int DEG>
using Lagr = Poly<COL, DEG, 5>;
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
  // split evaluation (ln P = add + ln mult): a log-native leaf only adds
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V&) {
    Mask neg;
    return fwd_log<CO>(cx, c, x, neg);
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
  // Hessian pass: upper triangle over (mu, sigmal, alphal, nl, sigmar, alphar, nr); an event touches mu and its side's slots
  template <int IS, int IA, int IN>
  ZNLL_DEV static void scatter4(double w, double sg, const double* U, double* h) {
    h[tri_index(0, 0, 7)] = fma(w, U[0], h[tri_index(0, 0, 7)]);
    h[tri_index(0, IS, 7)] = fma(w, U[1], h[tri_index(0, IS, 7)]);
    h[tri_index(0, IA, 7)] = fma(w, sg * U[2], h[tri_index(0, IA, 7)]);
    h[tri_index(0, IN, 7)] = fma(w, U[3], h[tri_index(0, IN, 7)]);
    h[tri_index(IS, IS, 7)] = fma(w, U[4], h[tri_index(IS, IS, 7)]);
    h[tri_index(IS, IA, 7)] = fma(w, sg * U[5], h[tri_index(IS, IA, 7)]);
    h[tri_index(IS, IN, 7)] = fma(w, U[6], h[tri_index(IS, IN, 7)]);
    h[tri_index(IA, IA, 7)] = fma(w, U[7], h[tri_index(IA, IA, 7)]);
    h[tri_index(IA, IN, 7)] = fma(w, sg * U[8], h[tri_index(IA, IN, 7)]);
    h[tri_index(IN, IN, 7)] = fma(w, U[9], h[tri_index(IN, IN, 7)]);
  }
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) const {
    const double* cc = c + CO;
    const double rr[2] = {la(r), lb(r)}, tt[2] = {la(t), lb(t)};
    const bool lf[2] = {la(left), lb(left)};
    bool tl[2] = {false, false};
    double dd[2] = {0.0, 0.0}, lv[2] = {0.0, 0.0};
    if (any_tail) {
      tl[0] = la(tail);
      tl[1] = lb(tail);
      dd[0] = la(dt);
      dd[1] = lb(dt);
      lv[0] = la(lu);
      lv[1] = lb(lu);
    }
#pragma unroll
    for (int k = 0; k < Lanes<V>::N; ++k) {
      const double* q = cc + (lf[k] ? 2 : 13);
      double U[10];
      tail_block(tl[k] ? cb_tail_derivs(q[2], q[3], dd[k] / q[3], lv[k]) : core_derivs(tt[k]), tt[k], q[0], q[1], U);
      if (lf[k])
        scatter4<1, 2, 3>(rr[k], q[10], U, h);
      else
        scatter4<4, 5, 6>(rr[k], q[10], U, h);
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
  static constexpr bool LEAF = true, HESS = true;
  static constexpr int ROOT = 0;
  static constexpr int NH = 28;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = true, HAS_MULT = false;  // split evaluation: ln P = add + ln(mult)
  static constexpr bool TABS = true;
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
  static constexpr int NS = 0, NC = 0, COLMASK = 0, NH = 0;
  static constexpr bool TABS = false;
  static constexpr bool FLATHESS = true;  // every child is a leaf the Hessian pass knows
  static constexpr bool ANY_ADD = false, ANY_MULT = false;
};
template <class H, class... T>
struct Meta<H, T...> {
  static constexpr int NS = H::NS + Meta<T...>::NS;
  static constexpr int NC = H::NC + Meta<T...>::NC;
  static constexpr int COLMASK = H::COLMASK | Meta<T...>::COLMASK;
  static constexpr bool TABS = H::TABS || Meta<T...>::TABS;
  static constexpr int NH = H::NH + Meta<T...>::NH;
  static constexpr bool FLATHESS = H::LEAF && H::HESS && Meta<T...>::FLATHESS;
  static constexpr bool ANY_ADD = H::HAS_ADD || Meta<T...>::ANY_ADD, ANY_MULT = H::HAS_MULT || Meta<T...>::ANY_MULT;
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
  template <int CO>
  ZNLL_DEV void fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V& mult) {
    const V v = fwd_val<CO>(cx, c, x);
    flag_mult(cx, v);
    rP = rcp_lanes(v, zlaneprod(v));
    mult = zmul(mult, v);
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
// SumPDF of exactly two log-native children (the signal + background case), in log-sum-exp form with child 1 as the
// fixed reference:
//   ln P = l1 + ln S,   S = f0*e^(l0-l1) + f1
// One exp per event instead of two, the Sum becomes log-native itself (no exp at all when it sits under a ProductPDF),
// and -- unlike the max-based form -- no per-lane selects: the common path is straight-line.  d ln P/d f_0 = e^d/S,
// d ln P/d f_1 = 1/S; the adjoint of child i is g*f_i*(d ln P/d f_i).  |l0 - l1| >= 600 (where e^d could overflow S, or
// underflows) is the rare path and uses the max-based form, in which one of the two exponentials is exp(0) = 1.
ZNLL_DEV bool lse_in_range(double d) { return (__double2hiint(d) & 0x7fffffff) < 0x4082C000; }  // |d| < 600
template <class V, class C0, class C1>
struct SumLseN {
  typedef typename Lanes<V>::Mask Mask;
  typename C0::template Node<V> k0;
  typename C1::template Node<V> k1;
  V q0, q1;  // d ln P / d f_i
  V P;

  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    Mask n0 = Lanes<V>::mask(false, false), n1 = n0;
    const V l0 = k0.template fwd_log<CO + 2>(cx, c, x, n0);
    const V l1 = k1.template fwd_log<CO + 2 + C0::NC>(cx, c, x, n1);
    const V d = zsub(l0, l1);
    const V one = zbc<V>(1.0);
    if (!cx.careful) {
      flag_abs600(cx, d);
    } else if (!(lse_in_range(la(d)) && lse_in_range(lb(d)))) {  // rare: max-based form, exp through libdevice (0 for a huge |d|, NaN stays NaN)
      const Mask lt = zsignbit(d);
      const V e = Lanes<V>::make(exp(-fabs(la(d))), exp(-fabs(lb(d))));
      V e0 = zsel(lt, e, one), e1 = zsel(lt, one, e);
      e0 = zsel(n0, zneg(e0), e0);
      e1 = zsel(n1, zneg(e1), e1);
      const V S = zfma(e0, c[CO + 0], zmul(e1, c[CO + 1]));
      V lS, rS;
      Mask ng;
      fast_logabs_rcp(S, cx, lS, rS, ng);
      neg = mxor(neg, ng);
      q0 = zmul(e0, rS);
      q1 = zmul(e1, rS);
      return zadd(zsel(lt, l1, l0), lS);
    }
    const V e = exp_core(d, cx);
    V lS, rS;
    Mask ng;
    if (!cx.careful) {  // optimistic sweep: a child with a negative factor has flagged the pair already
      const V S = zfma(e, c[CO + 0], c[CO + 1]);
      fast_logabs_rcp(S, cx, lS, rS, ng);
      q0 = zmul(e, rS);
      q1 = rS;
    } else {
      const V e0 = zsel(n0, zneg(e), e);  // children that are themselves products with a negative factor
      const V e1 = zsel(n1, zbc<V>(-1.0), one);
      const V S = zfma(e0, c[CO + 0], zmul(e1, c[CO + 1]));
      fast_logabs_rcp(S, cx, lS, rS, ng);  // one range test for both; S < 0 (a negative fraction) is the rare path
      neg = mxor(neg, ng);
      q0 = zmul(e0, rS);
      q1 = zmul(e1, rS);
    }
    return zadd(l1, lS);
  }
  template <int CO>
  ZNLL_DEV V fwd_val(const Ctx& cx, const double* __restrict__ c, const V* x) {
    Mask neg = Lanes<V>::mask(false, false);
    const V e = fast_exp(fwd_log<CO>(cx, c, x, neg), cx);
    P = zsel(neg, zneg(e), e);
    return P;
  }
  // split evaluation (optimistic sweep only): ln P = l1 + ln S with S handed to the pair's running product -- no log
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V& mult) {
    Mask n0 = Lanes<V>::mask(false, false), n1 = n0;  // (a negative factor inside a child has flagged the pair already)
    const V l0 = k0.template fwd_log<CO + 2>(cx, c, x, n0);
    const V l1 = k1.template fwd_log<CO + 2 + C0::NC>(cx, c, x, n1);
    const V d = zsub(l0, l1);
    flag_abs600(cx, d);
    const V e = exp_core(d, cx);
    const V S = zfma(e, c[CO + 0], c[CO + 1]);
    flag_mult(cx, S);
    const V rS = rcp_lanes(S, zlaneprod(S));
    q0 = zmul(e, rS);
    q1 = rS;
    mult = zmul(mult, S);
    return l1;
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

// The Hessian pass (hess_kernel, below) knows a Sum root whose branches are leaves or Products of leaves.  Seen from the
// root a branch is one generalised leaf u_k: it contributes r_k * (d2 u_k / u_k) over ALL of its slots -- for a Product
// of leaves over disjoint observables the leaves' own blocks plus, between two leaves, the outer product of their scores.
template <class... Cs>
struct Prod;
template <class... Cs>
struct CrossCount {  // sum over pairs of children i < j of NS_i * NS_j
  static constexpr int value = 0;
};
template <class H, class... T>
struct CrossCount<H, T...> {
  static constexpr int value = H::NS * Meta<T...>::NS + CrossCount<T...>::value;
};
template <class T>
struct Branch {  // a leaf
  static constexpr bool OK = T::LEAF && T::HESS;
  static constexpr int NH = T::NH;
};
template <class... Cs>
struct Branch<Prod<Cs...>> {
  static constexpr bool OK = Meta<Cs...>::FLATHESS;
  static constexpr int NH = Meta<Cs...>::NH + CrossCount<Cs...>::value;
};
template <class... Cs>
struct BranchMeta {
  static constexpr bool OK = true;
  static constexpr int NH = 0;
};
template <class H, class... T>
struct BranchMeta<H, T...> {
  static constexpr bool OK = Branch<H>::OK && BranchMeta<T...>::OK;
  static constexpr int NH = Branch<H>::NH + BranchMeta<T...>::NH;
};
// the children's second-order blocks, one after the other (HK advances by each child's Branch<>::NH)
template <int CK, int HK, int I, class KD, class RF>
ZNLL_DEV void second_rec(KD& kd, const RF& rof, const double* __restrict__ c, double* hacc) {
  if constexpr (KD::N > 0) {
    kd.h.template second<CK>(rof(I), c, hacc + HK);
    second_rec<CK + KD::Head::NC, HK + Branch<typename KD::Head>::NH, I + 1>(kd.t, rof, c, hacc);
  }
}

template <class... Cs>
struct Sum {
  static constexpr int K = sizeof...(Cs);
  static constexpr int NS = K + Meta<Cs...>::NS, NC = K + Meta<Cs...>::NC;
  static constexpr int COLMASK = Meta<Cs...>::COLMASK;
  static constexpr bool LOGNAT = AllLog<Cs...>::value;
  static constexpr bool HAS_ADD = AllLog<Cs...>::value, HAS_MULT = true;  // log-sum-exp: l1 + ln S; value-space sum: ln P
  static constexpr bool TABS = true;
  static constexpr bool LEAF = false, HESS = BranchMeta<Cs...>::OK;  // Sum of leaves and of Products of leaves
  static constexpr int NH = BranchMeta<Cs...>::NH;
  static constexpr int ROOT = 1;  // how hess_kernel treats the root: 1 Sum, 2 Prod, 0 leaf
  template <class V>
  using Node = typename SumPick<V, Cs...>::type;
  template <class V>
  using HessNode = SumN<V, Cs...>;  // the Hessian pass always uses the value-space Sum (uniform access to the children)
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
  // split evaluation: log-native factors add, value-space factors (polynomials, value-space sums) multiply
  template <int CK, bool STARTED, class KD>
  ZNLL_DEV void split_rec(KD& kd, const Ctx& cx, const double* __restrict__ c, const V* x, V& add, V& mult) {
    if constexpr (KD::N > 0) {
      if constexpr (KD::Head::HAS_ADD) {
        const V a = kd.h.template fwd_split<CK>(cx, c, x, mult);
        if constexpr (STARTED)
          add = zadd(add, a);
        else
          add = a;
        split_rec<CK + KD::Head::NC, true>(kd.t, cx, c, x, add, mult);
      } else {
        kd.h.template fwd_split<CK>(cx, c, x, mult);
        split_rec<CK + KD::Head::NC, STARTED>(kd.t, cx, c, x, add, mult);
      }
    }
  }
  template <int CO>
  ZNLL_DEV V fwd_split(const Ctx& cx, const double* __restrict__ c, const V* x, V& mult) {
    V add = zbc<V>(0.0);  // overwritten by the first adding factor (HAS_ADD says whether there is one)
    split_rec<CO, false>(kids, cx, c, x, add, mult);
    return add;
  }
  template <int CO>
  ZNLL_DEV V fwd_log(const Ctx& cx, const double* __restrict__ c, const V* x, Mask& neg) {
    V tot = kids.h.template fwd_log<CO>(cx, c, x, neg);  // at least one factor
    fwd_rec<CO + Kids<V, Cs...>::Head::NC>(kids.t, cx, c, x, tot, neg);
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
  // As a branch of a Sum root (Hessian pass): r * (d2 u / u) over all of the branch's slots.  Layout: the leaves' own
  // triangles in pre-order (their `second`), then for every pair of leaves i < j the full NS_i x NS_j block
  // r * s_a * s_b of their scores s = d ln p / d slot (row-major).
  template <int CO>
  ZNLL_DEV void second(V r, const double* __restrict__ c, double* h) {
    constexpr int K = sizeof...(Cs);
    constexpr int ns[K] = {Cs::NS...};
    constexpr int NSB = Meta<Cs...>::NS;
    second_rec<CO, 0, 0>(kids, [&](int) { return r; }, c, h);
    V sc[NSB > 0 ? NSB : 1];
#pragma unroll
    for (int a = 0; a < NSB; ++a) sc[a] = zbc<V>(0.0);
    bwd_rec<CO, 0>(kids, zbc<V>(1.0), c, sc);
    int q = Meta<Cs...>::NH, oi = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      int oj = oi + ns[i];
#pragma unroll
      for (int j = i + 1; j < K; ++j) {
#pragma unroll
        for (int a = 0; a < ns[i]; ++a) {
          const V ra = zmul(r, sc[oi + a]);
#pragma unroll
          for (int b = 0; b < ns[j]; ++b) {
            zacc(h[q], ra, sc[oj + b]);
            ++q;
          }
        }
        oj += ns[j];
      }
      oi += ns[i];
    }
  }
};
template <class... Cs>
struct Prod {
  static constexpr int NS = Meta<Cs...>::NS, NC = Meta<Cs...>::NC;
  static constexpr int COLMASK = Meta<Cs...>::COLMASK;
  static constexpr bool LOGNAT = true;
  static constexpr bool HAS_ADD = Meta<Cs...>::ANY_ADD, HAS_MULT = Meta<Cs...>::ANY_MULT;
  static constexpr bool TABS = Meta<Cs...>::TABS;
  static constexpr bool LEAF = false, HESS = Meta<Cs...>::FLATHESS;  // flat Product of leaves
  static constexpr int NH = Meta<Cs...>::NH;
  static constexpr int ROOT = 2;
  template <class V>
  using Node = ProdN<V, Cs...>;
  template <class V>
  using HessNode = ProdN<V, Cs...>;
};

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
// Cross-GPU exchange fused into the last block of the (last channel's) kernel: every rank pushes its raw
// accumulators into all peers' exchange buffers with plain stores over NVLink (peer-mapped symmetric memory),
// publishes a sequence flag with release semantics at system scope, waits for the peers' flags, and sums the
// world's contributions in fixed rank order -- bit-identical on every rank, no NCCL call on the step path.
// Buffer layout per rank: [128 B: u64 flag per source rank][2 parities][world][stride] doubles.  The stride is fixed
// (ZNLL_XSTRIDE, the largest accumulator block) so that ONE buffer per process serves every plan: the two parity regions
// of consecutive evaluations never overlap, whichever plans make them.
#define ZNLL_MAXPEERS 8
#define ZNLL_XFLAG_BYTES 128
#define ZNLL_XSTRIDE 256
struct Xchg {
  double* buf[ZNLL_MAXPEERS];  // buf[r]: rank r's exchange buffer as mapped into this process
  double* acc;                 // this rank's accumulators of ALL channels (plan-wide), reduced in place
  unsigned long long seq;      // evaluation counter (> 0), identical on all ranks
  int rank, world, total, stride;
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
                    ((size_t)(par * xc.world + xc.rank) * xc.stride + t);
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
                        (size_t)par * xc.world * xc.stride + t;
    double s = 0.0;
    for (int r = 0; r < xc.world; ++r) s += __ldcg(src + (size_t)r * xc.stride);  // fixed rank order
    xc.acc[t] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

template <int NC>
struct KArgs {
  const double* cols[ZNLL_MAXCOL];
  const double* w;
  long long n;
  double log_offset;
  double log_offset2;      // 2 * log_offset (filled by the launcher)
  double* partials;        // [gridDim.x][NACC]
  unsigned int* ticket;    // zero before launch; reset by the last block
  double* out;             // [NACC]
  Xchg xc;                 // cross-GPU exchange (world == 0: none)
  double c[NC > 0 ? NC : 1];
};

// accumulators: [0] running sum of w*l - offset, [1] its compensation, [2..] slots.
// Returns this event's (this pair's) contribution to [0]:  sum over lanes of w*l, minus `off` -- the log offset times
// the number of lanes, prepared by the caller.
//
// With an optimistic context (cx0.careful == false, the step kernels' pair loop) the forward sweep runs the fast cores
// unconditionally and every range test only feeds a running maximum; ONE test after the sweep (any_flag) decides whether the pair is evaluated
// again carefully (out-of-range argument of an exp / log / reciprocal, a negative factor, ln P < -690 where the
// p + 1e-307 of loss.py:117 matters).  With a careful context only the careful evaluation is instantiated.
template <class M, class V, bool WEIGHTED, bool GRAD, class A>
ZNLL_DEV double event(const Ctx& cx0, const double* __restrict__ c, const V* x, V w, double off, A* acc) {
  typedef typename Lanes<V>::Mask Mask;
  typename M::template Node<V> m;
  // weights multiply first, then the unweighted offset is subtracted per event (loss.py:134-137); the term is formed
  // inside each branch, so that only one double is merged where the optimistic and the careful path meet
  auto term_of = [&](V l) -> double {
    if constexpr (WEIGHTED)
      return zdot_minus(w, l, off);
    else
      return zsum_minus(l, off);
  };
  double term = 0.0;
  bool redo = cx0.careful;
  if (!cx0.careful) {
    reset_flags(cx0);
    V l;
    if constexpr (M::LOGNAT) {
      Mask neg = Lanes<V>::mask(false, false);
      l = m.template fwd_log<0>(cx0, c, x, neg);
      flag_add(cx0, l);  // ln P < -600: the careful evaluation decides whether the +1e-307 matters
      redo = any_flag(cx0);
      if (!redo) {
        term = term_of(l);
        if constexpr (GRAD) {
          // the adjoint of ln P is exactly 1 (or the weight): a compile-time constant the reverse sweep folds away
          if constexpr (WEIGHTED)
            m.template bwd_log<0, 2>(w, c, acc);
          else
            m.template bwd_log<0, 2>(zbc<V>(1.0), c, acc);
        }
      }
    } else {
      const V P = m.template fwd_val<0>(cx0, c, x);
      V g;
      log_eps_rcp<V, GRAD>(P, cx0, l, g);  // l = log P, g = 1/P: for P >= 2^-958 the sum P + 1e-307 IS P, bit for bit
      redo = any_flag(cx0);
      if (!redo) {
        term = term_of(l);
        if constexpr (GRAD) {
          if constexpr (WEIGHTED) g = zmul(g, w);
          m.template bwd_val<0, 2>(g, c, acc);
        }
      }
    }
  }
  if (redo) {  // rare in the pair loop; the only path of a careful context
    const Ctx cx = with_mode(cx0, true);
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
        if constexpr (WEIGHTED)
          m.template bwd_log<0, 2>(w, c, acc);
        else
          m.template bwd_log<0, 2>(zbc<V>(1.0), c, acc);
      }
    } else {
      const V P = m.template fwd_val<0>(cx, c, x);
      V g;
      log_eps_rcp<V, GRAD>(P, cx, l, g);  // l = log(P + 1e-307), g = 1/(P + 1e-307); one range test for both
      if constexpr (GRAD) {
        if constexpr (WEIGHTED) g = zmul(g, w);
        m.template bwd_val<0, 2>(g, c, acc);
      }
    }
    term = term_of(l);
  }
  return term;
}

// Split evaluation of an unweighted pair in the optimistic sweep (round 2): ln P = add + ln(mult), where `mult` -- the
// argument of the root's logarithm: S of a log-sum-exp, P of a value-space sum, the polynomial factors of a product -- is
// NOT put through a log per event but multiplied into the thread's running product `prod`; the thread takes ONE log of
// that product when its exponent drifts out of [2^-512, 2^512) and at the end of its loop (sum of logs = log of the
// product; the rounding error of n factors, ~sqrt(n) 1.1e-16 relative, is that of n rounded logs).  This removes the
// per-event log (11 DP + ~12 integer instructions per lane) from every unweighted tree with a Sum or a polynomial in it;
// the gradient still needs 1/mult per event (rcp_core).  Weighted data needs w_i ln(mult_i) and keeps the per-event log.
// Every factor must be positive and within [2^-200, 2^200), every add above -600 (so that ln P > -690 and the +1e-307 of
// loss.py:117 is invisible); anything else flags the pair, which is then evaluated carefully, event by event, through
// event<>() -- `prod` untouched.  Returns the pair's additive term (sum over lanes of add, minus `off`).
template <class M, class V, bool GRAD, class A>
ZNLL_DEV double event_split(const Ctx& cx0, const double* __restrict__ c, const V* x, double off, A* acc, double& prod) {
  typename M::template Node<V> m;
  reset_flags(cx0);
  V mult = zbc<V>(1.0);
  double term;
  if constexpr (M::HAS_ADD) {
    const V add = m.template fwd_split<0>(cx0, c, x, mult);
    flag_add(cx0, add);
    term = zsum_minus(add, off);
  } else {
    m.template fwd_split<0>(cx0, c, x, mult);
    term = -off;
  }
  if (any_flag(cx0)) return event<M, V, false, GRAD>(with_mode(cx0, true), c, x, zbc<V>(1.0), off, acc);  // rare
  if constexpr (GRAD) m.template bwd_log<0, 2>(zbc<V>(1.0), c, acc);
  prod *= zlaneprod(mult);
  return term;
}
// fold the running product into the value accumulator (rare inside the loop, once at its end)
ZNLL_DEV void fold_product(double* acc, double& prod) {
  acc[0] += log(prod);
  prod = 1.0;
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
  ZNLL_DEV void load(const KArgs<NC>& args, unsigned i) {  // i: index of the event PAIR
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
// ZNLL_RING = 1 (experiment, off by default): trees that read three or four 8-byte streams per event stage them through a
// two-deep cp.async ring in shared memory, so that TWO pairs per thread are in flight instead of one.  Motivation: with
// the instruction stream of the end of round 1, cfg 4 (75.7 instructions per event) saturates no unit and runs at the
// one-pair-in-flight streaming ceiling (5.2 of 5.37 TB/s, stream_skeleton_bench.py; 6.38 TB/s with two pairs).  Every
// thread touches only its own 16-byte slots, so there are no bank conflicts and no block barrier in the loop.
// First measurement (one 100-step run on the B200, cfg 4): kernel 0.456 -> 0.445 ms back to back, 0.485 -> 0.420 ms per
// step; the product-tree parity and out-of-range canary tests pass with it.  It stays off until the whole GPU suite (four
// streams, NVRTC trees) has run with it.
#ifndef ZNLL_RING
#define ZNLL_RING 1
#endif
ZNLL_DEV void cp_async16(void* smem, const void* gmem) {  // LDGSTS, L2 only
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
ZNLL_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZNLL_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int N>
ZNLL_DEV void cp_async_wait_all_but() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// STAGES-deep cp.async ring: STAGES pairs per thread are in flight while one is computed.  The stage index is a run-time
// value (one multiply-add per access), so the loop body exists ONCE whatever the depth.
template <class M, bool WEIGHTED, int STAGES>
struct Ring {
  static constexpr int NSTREAM = zpopc((unsigned)M::COLMASK) + (WEIGHTED ? 1 : 0);
  double2 slot[STAGES][NSTREAM][ZNLL_BLOCK];  // [stage][stream][thread]
  template <int NC>
  ZNLL_DEV void issue(int stage, const KArgs<NC>& args, unsigned i) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) cp_async16(&slot[stage][s++][threadIdx.x], reinterpret_cast<const double2*>(args.cols[k]) + i);
    if constexpr (WEIGHTED) cp_async16(&slot[stage][s][threadIdx.x], reinterpret_cast<const double2*>(args.w) + i);
  }
  ZNLL_DEV void fetch(int stage, EventPair<M, WEIGHTED>& p) const {
    int s = 0;
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) {
        const double2 v = slot[stage][s++][threadIdx.x];
        p.x[k].a = v.x;
        p.x[k].b = v.y;
      }
    p.w.a = p.w.b = 1.0;
    if constexpr (WEIGHTED) {
      const double2 v = slot[stage][s][threadIdx.x];
      p.w.a = v.x;
      p.w.b = v.y;
    }
  }
};
// depth of the ring by the number of 8-byte streams a tree reads (0: register prefetch of one pair instead).  One pair per
// thread in flight is 148 SMs x 512 threads x 16 B x streams = 1.2 MB x streams on the wire; at ~0.8 us of HBM latency
// that caps a one-stream tree at ~1.5 TB/s and a two-stream tree at ~3 TB/s -- BELOW what cfg 2 / 3 (8 B per event in
// 0.46 - 0.49 ms) and cfg 5 (16 B in 0.55 ms) ask for, so once their instruction streams had been cut they became
// memory-LATENCY bound (ncu: long-scoreboard stalls 0.4 - 1.3 per issue).  Deeper prefetch is the cure, and shared
// memory is where it costs no registers.
#ifndef ZNLL_RING_STAGES_SMALL
#define ZNLL_RING_STAGES_SMALL 3  // trees with one or two streams
#endif
#ifndef ZNLL_RING_STAGES_WIDE
#define ZNLL_RING_STAGES_WIDE 2   // three or four streams (48 KB of static shared memory hold the tables + 2 x 4 x 4 KB)
#endif
__host__ __device__ constexpr int ring_stages(int nstream) {
  return (ZNLL_RING == 0 || nstream > 4) ? 0 : (nstream <= 2 ? ZNLL_RING_STAGES_SMALL : ZNLL_RING_STAGES_WIDE);
}

// The pair loop.  Per thread the value is accumulated with plain fp64 adds (a thread sees a few hundred pairs, each term
// is O(1) after the per-event offset: the rounding error of a thread's partial sum is below 1e-13 of it), and the partial
// sums of threads, warps, blocks and GPUs are then merged error-free (TwoSum, block_reduce_and_finish) -- so the result is
// a compensated sum at the granularity that matters for 1e8..1e9 events, and the per-pair Kahan step (4 dependent DADDs
// on the FP64 pipe that bounds these kernels) is gone.  Event pairs are indexed with 32 bits (the host refuses more
// than 8e9 events per channel and GPU): one IADD, one ISETP and one IMAD.WIDE per pair for the whole loop control.
#ifdef ZNLL_MAXNREG  // tuning experiments: an explicit register cap instead of the one ptxas derives from the CTA count
#define ZNLL_STEP_BOUNDS __maxnreg__(ZNLL_MAXNREG)
#else
#define ZNLL_STEP_BOUNDS __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS)
#endif
template <class M, bool WEIGHTED, bool GRAD>
__global__ void ZNLL_STEP_BOUNDS nll_kernel(const __grid_constant__ KArgs<M::NC> args) {
  constexpr int NACC = 2 + M::NS;
  __shared__ MathTabs tabs;
  if constexpr (NeedTabs<M>::value) {
    init_tabs(tabs);
    __syncthreads();
  }
  const Ctx cx = ctx_of(tabs, /*careful=*/false);  // optimistic pair loop: one rare-path test per pair (see event())
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  // unweighted trees whose root takes a log (a Sum, a polynomial factor): the log's arguments go into a running product
  // per thread instead (event_split)
  constexpr bool SPLIT = M::HAS_MULT && !WEIGHTED;
  double prod = 1.0;
  auto do_pair = [&](const EventPair<M, WEIGHTED>& p) {
    if constexpr (SPLIT) {
      acc[0] += event_split<M, D2, GRAD>(cx, c, p.x, args.log_offset2, acc, prod);
      if (!prod_in_range(prod)) fold_product(acc, prod);  // rare: every few hundred pairs at most
    } else {
      acc[0] += event<M, D2, WEIGHTED, GRAD>(cx, c, p.x, p.w, args.log_offset2, acc);
    }
  };
  const unsigned nvec = (unsigned)(args.n >> 1);
  const unsigned stride = gridDim.x * ZNLL_BLOCK;
  unsigned i = blockIdx.x * ZNLL_BLOCK + threadIdx.x;
  constexpr int NSTREAM = zpopc((unsigned)M::COLMASK) + (WEIGHTED ? 1 : 0);
  constexpr int STAGES = ring_stages(NSTREAM);
  // the ring, the tables and the reduction scratch must fit the 48 KB of static shared memory
  constexpr bool RING_FITS = STAGES * NSTREAM * ZNLL_BLOCK * 16 + sizeof(MathTabs) + (ZNLL_BLOCK / 32) * NACC * 8 + 256 <= 48 * 1024;
  if constexpr (STAGES > 0 && RING_FITS) {
    __shared__ Ring<M, WEIGHTED, (STAGES > 0 ? STAGES : 1)> ring;
    EventPair<M, WEIGHTED> p;
    // groups are committed unconditionally (an empty group is legal), so "all but the newest STAGES - 1" is always the
    // stage read next; a stage is refilled by the thread that has just read it (program order)
    unsigned pre = i;
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      if (pre < nvec) ring.issue(s, args, pre);
      cp_async_commit();
      pre += stride;
    }
    int stage = 0;
    while (i < nvec) {
      cp_async_wait_all_but<(STAGES > 0 ? STAGES - 1 : 0)>();
      ring.fetch(stage, p);
      if (pre < nvec) ring.issue(stage, args, pre);
      cp_async_commit();
      pre += stride;
      do_pair(p);
      i += stride;
      stage = (stage + 1 == STAGES) ? 0 : stage + 1;
    }
    cp_async_wait_all();
  } else if constexpr ((ZNLL_PINGPONG != 0) && (NSTREAM >= 3)) {
    // more streams than the ring holds: two register buffers used alternately -- the loop body exists twice and no
    // buffer is ever copied
    EventPair<M, WEIGHTED> b0, b1;
    if (i < nvec) b0.load(args, i);
    while (i < nvec) {
      i += stride;
      if (i < nvec) b1.load(args, i);
      do_pair(b0);
      if (i >= nvec) break;
      i += stride;
      if (i < nvec) b0.load(args, i);
      do_pair(b1);
    }
  } else {
    EventPair<M, WEIGHTED> cur, nxt;
    bool more = i < nvec;
    if (more) nxt.load(args, i);
    while (more) {
      cur = nxt;
      i += stride;
      more = i < nvec;
      if (more) nxt.load(args, i);  // software prefetch: the next pair is in flight while this one computes
      do_pair(cur);
    }
  }
  if constexpr (SPLIT) fold_product(acc, prod);
  if ((args.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail event: scalar, careful instantiation
    double xa[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) xa[k] = args.cols[k][args.n - 1];
    double wa = 1.0;
    if constexpr (WEIGHTED) wa = args.w[args.n - 1];
    kahan_add(acc, event<M, double, WEIGHTED, GRAD>(with_mode(cx, true), c, xa, wa, args.log_offset, acc));
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
  if constexpr (NeedTabs<M>::value) {
    init_tabs(tabs);
    __syncthreads();
  }
  const Ctx cx = ctx_of(tabs);
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  const unsigned nvec = (unsigned)(args.n >> 1);
  const unsigned stride = gridDim.x * ZNLL_BLOCK;
  unsigned i = blockIdx.x * ZNLL_BLOCK + threadIdx.x;
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
    kahan_add(acc, event<M, D2, WEIGHTED, true>(cx, c, cur.x, cur.w, 2.0 * args.log_offset, e));
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

// Exact Hessian pass for FLAT trees -- a leaf, a Sum of leaves, a Product of leaves: all five BASELINE configs --
// (reference: loss.hessian, core/loss.py:765-878, z/math.py:288-352, which differentiates the
// whole graph twice).  With l = ln P(x; slots), normalisation constants held fixed, one pass accumulates
//   T  = sum_i w_i [g_i; 1][g_i; 1]^T   (g_i = d l_i / d slot, UNweighted; upper triangle, (NS+1)(NS+2)/2 entries), and
//   B_k = sum_i r_ik (d2 u_k / da db) / u_k  per leaf k over the leaf's own slots (upper triangle, Node::second), with
//        r_ik = w_i f_k p_k / P (Sum root: the event's responsibility for leaf k) or w_i (Product / leaf root).
// Second derivatives of a flat tree are exactly: -g g^T everywhere (Sum) or inside each leaf block (Product), + B_k in
// leaf blocks, + g_b / f_k between a fraction and its own leaf's slots; the host (znll_eval_hess) assembles these and
// restores the slots' dependence through the normalisation integrals with their first and second derivatives.
// out: [0,1] value pair, then T (row-major upper triangle), then the B_k in pre-order.
template <class M, class V, bool LEAFROOT>
struct HessRootPick {
  typedef typename M::template Node<V> type;
};
template <class M, class V>
struct HessRootPick<M, V, false> {
  typedef typename M::template HessNode<V> type;
};
template <class M, class V, bool WEIGHTED>
ZNLL_DEV double hess_event(const Ctx& cx, const double* __restrict__ c, const V* x, V w, double off, double* acc) {
  typedef typename Lanes<V>::Mask Mask;
  constexpr int NS = M::NS, NE = NS + 1, NQ = NE * (NE + 1) / 2;
  typename HessRootPick<M, V, M::ROOT == 0>::type m;
  V e[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) e[j] = zbc<V>(0.0);
  V l;
  if constexpr (M::ROOT == 1) {
    const V P = m.template fwd_val<0>(cx, c, x);
    V g;
    log_eps_rcp<V, true>(P, cx, l, g);
    m.template bwd_val<0, 0>(g, c, e);
  } else {
    Mask neg = Lanes<V>::mask(false, false);
    l = m.template fwd_log<0>(cx, c, x, neg);
    m.template bwd_log<0, 0>(zbc<V>(1.0), c, e);
  }
  e[NS] = zbc<V>(1.0);
  int q = 2;
#pragma unroll
  for (int a = 0; a < NE; ++a) {
    const V wa = zmul(w, e[a]);
#pragma unroll
    for (int b = a; b < NE; ++b) {
      zacc(acc[q], wa, e[b]);
      ++q;
    }
  }
  double* hacc = acc + 2 + NQ;
  if constexpr (M::ROOT == 1) {
    second_rec<M::K, 0, 0>(m.kids, [&](int i) { return zmul(zmul(w, c[i]), e[i]); }, c, hacc);
  } else if constexpr (M::ROOT == 2) {
    second_rec<0, 0, 0>(m.kids, [&](int) { return w; }, c, hacc);
  } else {
    m.template second<0>(w, c, hacc);
  }
  if constexpr (WEIGHTED)
    return zdot_minus(w, l, off);
  else
    return zsum_minus(l, off);
}

template <class M, bool WEIGHTED>
__global__ void __launch_bounds__(ZNLL_BLOCK, 1) hess_kernel(const __grid_constant__ KArgs<M::NC> args) {
  constexpr int NS = M::NS, NE = NS + 1, NQ = NE * (NE + 1) / 2, NACC = 2 + NQ + M::NH;
  __shared__ MathTabs tabs;
  init_tabs(tabs);
  __syncthreads();
  const Ctx cx = ctx_of(tabs);
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  const double* __restrict__ c = args.c;
  const unsigned nvec = (unsigned)(args.n >> 1);
  const unsigned stride = gridDim.x * ZNLL_BLOCK;
  unsigned i = blockIdx.x * ZNLL_BLOCK + threadIdx.x;
  EventPair<M, WEIGHTED> cur, nxt;
  if (i < nvec) nxt.load(args, i);
  while (i < nvec) {
    cur = nxt;
    i += stride;
    if (i < nvec) nxt.load(args, i);
    kahan_add(acc, hess_event<M, D2, WEIGHTED>(cx, c, cur.x, cur.w, args.log_offset2, acc));
  }
  if ((args.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail event
    double x[ZNLL_MAXCOL];
#pragma unroll
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) x[k] = args.cols[k][args.n - 1];
    double w = 1.0;
    if constexpr (WEIGHTED) w = args.w[args.n - 1];
    kahan_add(acc, hess_event<M, double, WEIGHTED>(cx, c, x, w, args.log_offset, acc));
  }
  block_reduce_and_finish<NACC>(acc, args.partials, args.ticket, args.out, nullptr);
}

// per-event normalised pdf values (BasePDF.pdf, core/basepdf.py:398-429): out[i] = P(x_i); args.partials = out
template <class M>
__global__ void __launch_bounds__(ZNLL_BLOCK, ZNLL_MIN_BLOCKS) pdf_kernel(const __grid_constant__ KArgs<M::NC> args) {
  __shared__ MathTabs tabs;
  if constexpr (NeedTabs<M>::value) {
    init_tabs(tabs);
    __syncthreads();
  }
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
