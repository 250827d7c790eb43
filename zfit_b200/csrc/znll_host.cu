// znll_host.cu -- C ABI (include/znll.h) of the B200-native unbinned-likelihood hot path.
//
// Host side of every evaluation:
//   1. prologue   : walk the model tree, compute each leaf's analytic normalisation integral and
//                   its slot derivatives in fp64 (libm), fill the kernel's constant block;
//   2. launch     : one fused kernel per channel (ahead-of-time catalogue instantiation of the
//                   templates in znll_device.cuh, or the same templates instantiated through NVRTC
//                   for tree shapes outside the catalogue);
//   3. epilogue   : fold the normalisation derivatives into the accumulated slot gradients.
// There is no CPU evaluation path: without a CUDA device every compute entry point fails.
#include <ctype.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/znll.h"
#include "znll_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define CUDA_OK(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e_));   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// catalogue registry
// ------------------------------------------------------------------------------------------------
static std::vector<ZnllCatalogEntry>& catalog() {
  static std::vector<ZnllCatalogEntry> v;
  return v;
}
void znll_catalog_register(const ZnllCatalogEntry* e, int n) {
  for (int i = 0; i < n; ++i) catalog().push_back(e[i]);
}
const ZnllCatalogEntry* znll_catalog_find(const char* name) {
  for (auto& e : catalog())
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}
int znll_catalog_size() { return (int)catalog().size(); }
const ZnllCatalogEntry* znll_catalog_at(int i) { return &catalog()[i]; }

// ------------------------------------------------------------------------------------------------
// leaf integrals and their slot derivatives (host fp64)
// ------------------------------------------------------------------------------------------------
static const double kSqrt2 = 1.4142135623730951;
static const double kSqrtHalf = 0.7071067811865476;
static const double kInvSqrt2Pi = 0.3989422804014327;
static const double kHalfLog2Pi = 0.9189385332046727;
static const double kSqrtPiOver2 = 1.2533141373155001;

// tfp.math.special.ndtr (Normal.cdf), reference call site src/zfit/models/dist_tfp.py:113-120
static double ndtr(double x) {
  const double w = x * kSqrtHalf, z = fabs(w);
  double y;
  if (z < kSqrtHalf)
    y = 1.0 + erf(w);
  else if (w > 0)
    y = 2.0 - erfc(z);
  else
    y = erfc(z);
  return 0.5 * y;
}
static double zphi(double z, int pow_z) {  // z^pow_z * phi(z), 0 at +-inf
  if (isinf(z)) return 0.0;
  const double p = kInvSqrt2Pi * exp(-0.5 * z * z);
  return pow_z ? z * p : p;
}

static void gauss_integral(double mu, double sigma, double lo, double hi, double* I, double* dI) {
  const double zh = (hi - mu) / sigma, zl = (lo - mu) / sigma;
  *I = ndtr(zh) - ndtr(zl);
  dI[0] = (zphi(zl, 0) - zphi(zh, 0)) / sigma;
  dI[1] = (zphi(zl, 1) - zphi(zh, 1)) / sigma;
}

// TruncatedGauss (dist_tfp.py:357-417 over tfp TruncatedNormal): the integral over [lo, hi] of the shape that is zero
// outside [low, high] -- the truncation's own normaliser cancels in pdf/integral.  slots [mu, sigma, low, high]
static void tgauss_integral(const double* s, double lo, double hi, double* I, double* dI) {
  const double mu = s[0], sigma = s[1], low = s[2], high = s[3];
  const bool cut_lo = low > lo, cut_hi = high < hi;
  const double a = cut_lo ? low : lo, b = cut_hi ? high : hi;
  gauss_integral(mu, sigma, a, b, I, dI);
  dI[2] = cut_lo ? -zphi((a - mu) / sigma, 0) / sigma : 0.0;
  dI[3] = cut_hi ? zphi((b - mu) / sigma, 0) / sigma : 0.0;
}

// src/zfit/models/basic.py:211-229 (shifted raw integral)
static void exp_integral(double lam, double lo, double hi, double s, double* I, double* dI) {
  const double eh = exp(lam * (hi - s)), el = exp(lam * (lo - s));
  *I = eh / lam - el / lam;
  dI[0] = ((hi - s) * eh - (lo - s) * el) / lam - (*I) / lam;
}

// src/zfit/models/polynomials.py:443-478
static void cheb_T(double z, int maxdeg, double* T) {
  T[0] = 1.0;
  if (maxdeg >= 1) T[1] = z;
  for (int k = 2; k <= maxdeg; ++k) T[k] = 2.0 * (z * T[k - 1]) - T[k - 2];
}
static void cheb_integral(const double* c, int deg, double lo, double hi, double slo, double shi, double* I,
                          double* dI) {
  const double zl = (2.0 * lo - slo - shi) / (shi - slo), zu = (2.0 * hi - slo - shi) / (shi - slo);
  std::vector<double> Tu(deg + 3), Tl(deg + 3), J(deg + 1);
  cheb_T(zu, deg + 1, Tu.data());
  cheb_T(zl, deg + 1, Tl.data());
  J[0] = zu - zl;
  if (deg >= 1) J[1] = 0.5 * (zu * zu - zl * zl);
  double integral = c[0] * J[0];
  if (deg >= 1) integral += c[1] * J[1];
  if (deg >= 2) {
    double up = 0.0, low = 0.0;
    for (int k = 2; k <= deg; ++k) {
      const double nf = (double)k;
      const double iu = nf * Tu[k + 1] / (nf * nf - 1.0) - zu * Tu[k] / (nf - 1.0);
      const double il = nf * Tl[k + 1] / (nf * nf - 1.0) - zl * Tl[k] / (nf - 1.0);
      J[k] = iu - il;
      up += c[k] * iu;
      low += c[k] * il;
    }
    integral += up - low;
  }
  const double scale = 0.5 * (shi - slo);
  *I = integral * scale;
  for (int k = 0; k <= deg; ++k) dI[k] = J[k] * scale;
}

// src/zfit/models/physics.py:83-142, plus the derivative of exactly those expressions
static void cb_integral(double mu, double sigma, double alpha, double n, double lo, double hi, double* I,
                        double* dI) {
  const bool use_log = fabs(n - 1.0) < 1e-05;
  const double s = fabs(sigma), a = fabs(alpha);
  const double sgn_sigma = sigma < 0 ? -1.0 : 1.0, sgn_alpha = alpha < 0 ? -1.0 : 1.0;
  double tmin = (lo - mu) / s, tmax = (hi - mu) / s;
  if (alpha < 0) {
    const double tm = tmin;
    tmin = -tmax;
    tmax = -tm;
  }
  const double A = pow(n / a, n) * exp(-0.5 * a * a);
  const double B = n / a - a;
  const double dA_dn = A * (log(n / a) + 1.0), dA_da = A * (-n / a - a);
  const double dB_da = -n / (a * a) - 1.0, inva = 1.0 / a;
  // integrand at a boundary t (the derivative of the reference's formula w.r.t. that boundary)
  auto f_at = [&](double t) {
    if (t < -a) {
      const double u = B - t;
      return use_log ? A / u : A * pow(u, -n);
    }
    return exp(-0.5 * t * t);
  };
  double Ival, dIa = 0.0, dIn = 0.0;
  if (tmin >= -a) {
    Ival = s * kSqrtPiOver2 * (erf(tmax / kSqrt2) - erf(tmin / kSqrt2));
  } else {
    const bool all_tail = tmax <= -a;
    double u1 = B - tmin, u2 = all_tail ? B - tmax : n / a;
    if (!(u1 > 0)) u1 = 1.0;  // safe_b_tmin_ones
    if (all_tail && !(u2 > 0)) u2 = 1.0;
    const double du1_da = dB_da, du2_da = all_tail ? dB_da : -n / (a * a);
    double tailI, dT_dn, dT_da;
    if (use_log) {
      const double L = log(u1) - log(u2);
      tailI = A * s * L;
      dT_dn = s * (dA_dn * L + A * (1.0 / u1 - 1.0 / u2) * inva);
      dT_da = s * (dA_da * L + A * (du1_da / u1 - du2_da / u2));
    } else {
      const double m = 1.0 - n;
      const double p1 = 1.0 / pow(u1, n - 1.0), p2 = 1.0 / pow(u2, n - 1.0);
      tailI = A * s / m * (p1 - p2);
      dT_dn = s * ((dA_dn / m + A / (m * m)) * (p1 - p2) +
                   A / m * (p1 * (-log(u1) + m * inva / u1) - p2 * (-log(u2) + m * inva / u2)));
      dT_da = s * (dA_da / m * (p1 - p2) + A * (p1 / u1 * du1_da - p2 / u2 * du2_da));
    }
    double gaussI = 0.0, dG_da = 0.0;
    if (!all_tail) {
      gaussI = s * kSqrtPiOver2 * (erf(tmax / kSqrt2) - erf(-a / kSqrt2));
      dG_da = s * exp(-0.5 * a * a);
    }
    Ival = all_tail ? tailI : tailI + gaussI;
    dIa = dT_da + dG_da;
    dIn = dT_dn;
  }
  const double fmax = f_at(tmax), fmin = f_at(tmin);
  *I = Ival;
  dI[0] = -sgn_alpha * (fmax - fmin);
  dI[1] = sgn_sigma * (Ival / s - (tmax * fmax - tmin * fmin));
  dI[2] = sgn_alpha * dIa;
  dI[3] = dIn;
}

// Legendre: src/zfit/models/polynomials.py:193-228;  Chebyshev2: :565-590 (integral of U_n is T_{n+1}/(n+1))
static void poly_integral(int kind, const double* c, int deg, double lo, double hi, double slo, double shi, double* I,
                          double* dI) {
  if (kind == ZNLL_CHEB) return cheb_integral(c, deg, lo, hi, slo, shi, I, dI);
  if (kind == ZNLL_BERNSTEIN) {
    // polynomials.py:1000-1025: int_0^t B_{k,m} = (1/(m+1)) sum_{j>k} B_{j,m+1}(t); times the volume of the space
    const int m1 = deg + 1;
    auto basis = [&](double t, double* B) {  // B_{j,m1}(t), j = 0..m1
      double binom = 1.0;
      for (int j = 0; j <= m1; ++j) {
        B[j] = binom * pow(t, j) * pow(1.0 - t, m1 - j);
        binom = binom * (double)(m1 - j) / (double)(j + 1);
      }
    };
    std::vector<double> Bu(m1 + 1), Bl(m1 + 1);
    basis((hi - slo) / (shi - slo), Bu.data());
    basis((lo - slo) / (shi - slo), Bl.data());
    double integral = 0.0;
    for (int k = 0; k <= deg; ++k) {
      double j = 0.0;
      for (int q = k + 1; q <= m1; ++q) j += Bu[q] - Bl[q];
      dI[k] = j / (double)m1 * (shi - slo);
      integral += c[k] * dI[k];
    }
    *I = integral;
    return;
  }
  const double zl = (2.0 * lo - slo - shi) / (shi - slo), zu = (2.0 * hi - slo - shi) / (shi - slo);
  std::vector<double> Pu(deg + 3), Pl(deg + 3), J(deg + 1);
  if (kind == ZNLL_LEGENDRE) {
    auto rec = [&](double z, double* P) {
      P[0] = 1.0;
      P[1] = z;
      for (int n = 1; n <= deg; ++n) P[n + 1] = ((2.0 * n + 1.0) * (z * P[n]) - n * P[n - 1]) / (n + 1.0);
    };
    rec(zu, Pu.data());
    rec(zl, Pl.data());
    J[0] = zu - zl;
    for (int k = 1; k <= deg; ++k) J[k] = (Pu[k + 1] - Pu[k - 1]) / (2.0 * k + 1.0) - (Pl[k + 1] - Pl[k - 1]) / (2.0 * k + 1.0);
  } else if (kind == ZNLL_HERMITE) {  // func_integral_hermite, polynomials.py:841-866: int H_k = H_{k+1} / (2 (k+1))
    auto rec = [&](double z, double* P) {
      P[0] = 1.0;
      P[1] = 2.0 * z;
      for (int n = 1; n <= deg; ++n) P[n + 1] = 2.0 * (z * P[n] - n * P[n - 1]);
    };
    rec(zu, Pu.data());
    rec(zl, Pl.data());
    for (int k = 0; k <= deg; ++k) J[k] = (Pu[k + 1] - Pl[k + 1]) / (2.0 * (k + 1.0));
  } else if (kind == ZNLL_LAGUERRE) {  // func_integral_laguerre, polynomials.py:711-746: int L_k = -L^(-1)_{k+1}
    auto rec = [&](double z, double* P) {  // generalised Laguerre, alpha = -1: (n+1) P_{n+1} = (2n - z) P_n - (n-1) P_{n-1}
      P[0] = 1.0;
      P[1] = -z;
      for (int n = 1; n <= deg; ++n) P[n + 1] = ((2.0 * n - z) * P[n] - (n - 1.0) * P[n - 1]) / (n + 1.0);
    };
    rec(zu, Pu.data());
    rec(zl, Pl.data());
    for (int k = 0; k <= deg; ++k) J[k] = -(Pu[k + 1] - Pl[k + 1]);
  } else {  // Chebyshev2
    cheb_T(zu, deg + 1, Pu.data());
    cheb_T(zl, deg + 1, Pl.data());
    for (int k = 0; k <= deg; ++k) J[k] = (Pu[k + 1] - Pl[k + 1]) / (k + 1.0);
  }
  double integral = 0.0;
  for (int k = 0; k <= deg; ++k) integral += c[k] * J[k];
  const double scale = 0.5 * (shi - slo);
  *I = integral * scale;
  for (int k = 0; k <= deg; ++k) dI[k] = J[k] * scale;
}

// GeneralizedCB: double_crystalball_mu_integral_func / generalized_..., src/zfit/models/physics.py:166-232
// slots [mu, sigmal, alphal, nl, sigmar, alphar, nr]
static void gcb_integral(const double* s, double lo, double hi, double* I, double* dI) {
  const double mu = s[0];
  for (int i = 0; i < 7; ++i) dI[i] = 0.0;
  double total = 0.0;
  if (!(mu < lo)) {  // left part: [lo, min(mu, hi)]
    double Il, d[4];
    cb_integral(mu, s[1], s[2], s[3], lo, mu < hi ? mu : hi, &Il, d);
    total += Il;
    dI[0] += d[0];
    dI[1] += d[1];
    dI[2] += d[2];
    dI[3] += d[3];
    if (mu < hi) dI[0] += 1.0;  // the upper limit is mu itself: + f(t=0) = 1
  }
  if (!(mu > hi)) {  // right part: [max(mu, lo), hi] with alpha -> -alphar
    double Ir, d[4];
    cb_integral(mu, s[4], -s[5], s[6], mu > lo ? mu : lo, hi, &Ir, d);
    total += Ir;
    dI[0] += d[0];
    dI[4] += d[1];
    dI[5] -= d[2];
    dI[6] += d[3];
    if (mu > lo) dI[0] -= 1.0;  // the lower limit is mu itself: - f(t=0) = -1
  }
  *I = total;
}

// GaussExpTail: gaussexptail_integral_func, src/zfit/models/physics.py:647-686, plus its derivative
static void get_integral(double mu, double sigma, double alpha, double lo, double hi, double* I, double* dI) {
  const double s = fabs(sigma), a = fabs(alpha);
  const double sgn_sigma = sigma < 0 ? -1.0 : 1.0, sgn_alpha = alpha < 0 ? -1.0 : 1.0;
  double tmin = (lo - mu) / s, tmax = (hi - mu) / s;
  if (alpha < 0) {
    const double tm = tmin;
    tmin = -tmax;
    tmax = -tm;
  }
  auto f_at = [&](double t) { return t < -a ? exp(0.5 * a * a + a * t) : exp(-0.5 * t * t); };
  const double ea = exp(0.5 * a * a);
  double Ival, dIa = 0.0;
  if (tmin >= -a) {
    Ival = s * kSqrtPiOver2 * (erf(tmax / kSqrt2) - erf(tmin / kSqrt2));
  } else {
    const bool all_tail = tmax <= -a;
    const double eh = all_tail ? exp(a * tmax) : exp(-a * a), el = exp(a * tmin);
    const double th = all_tail ? tmax : -a;
    const double tailI = s / a * ea * (eh - el);
    double gaussI = 0.0;
    if (!all_tail) gaussI = s * kSqrtPiOver2 * (erf(tmax / kSqrt2) - erf(-a / kSqrt2));
    Ival = tailI + gaussI;
    // d/da of the tail integral; the moving break point cancels against the Gaussian part (continuity at t = -a)
    dIa = s * ea * (eh * (1.0 + th / a - 1.0 / (a * a)) - el * (1.0 + tmin / a - 1.0 / (a * a)));
  }
  const double fmax = f_at(tmax), fmin = f_at(tmin);
  *I = Ival;
  dI[0] = -sgn_alpha * (fmax - fmin);
  dI[1] = sgn_sigma * (Ival / s - (tmax * fmax - tmin * fmin));
  dI[2] = sgn_alpha * dIa;
}

// GeneralizedGaussExpTail: generalized_gaussexptail_integral_func, physics.py:712-724; slots [mu, sl, al, sr, ar]
static void gget_integral(const double* s, double lo, double hi, double* I, double* dI) {
  const double mu = s[0];
  for (int i = 0; i < 5; ++i) dI[i] = 0.0;
  double total = 0.0;
  if (!(mu < lo)) {
    double Il, d[3];
    get_integral(mu, s[1], s[2], lo, mu < hi ? mu : hi, &Il, d);
    total += Il;
    dI[0] += d[0];
    dI[1] += d[1];
    dI[2] += d[2];
    if (mu < hi) dI[0] += 1.0;
  }
  if (!(mu > hi)) {
    double Ir, d[3];
    get_integral(mu, s[3], -s[4], mu > lo ? mu : lo, hi, &Ir, d);
    total += Ir;
    dI[0] += d[0];
    dI[3] += d[1];
    dI[4] -= d[2];
    if (mu > lo) dI[0] -= 1.0;
  }
  *I = total;
}

static int leaf_n_slots(const znll_node& nd) {
  switch (nd.kind) {
    case ZNLL_GAUSS: return 2;
    case ZNLL_EXP: return 1;
    case ZNLL_CB: return 4;
    case ZNLL_CHEB:
    case ZNLL_LEGENDRE:
    case ZNLL_CHEB2:
    case ZNLL_HERMITE:
    case ZNLL_LAGUERRE:
    case ZNLL_BERNSTEIN: return nd.degree + 1;
    case ZNLL_TGAUSS: return 4;
    case ZNLL_GCB: return 7;
    case ZNLL_GET: return 3;
    case ZNLL_GGET: return 5;
    default: return -1;
  }
}
static int leaf_n_consts(const znll_node& nd) {
  switch (nd.kind) {
    case ZNLL_GAUSS: return 3;
    case ZNLL_EXP: return 3;
    case ZNLL_CB: return 13;
    case ZNLL_CHEB:
    case ZNLL_LEGENDRE:
    case ZNLL_CHEB2:
    case ZNLL_HERMITE:
    case ZNLL_LAGUERRE:
    case ZNLL_BERNSTEIN: return 3 + nd.degree + 1;
    case ZNLL_TGAUSS: return 5;
    case ZNLL_GCB: return 24;
    case ZNLL_GET: return 7;
    case ZNLL_GGET: return 12;
    default: return -1;
  }
}

static int leaf_integral(const znll_node& nd, const double* s, double* I, double* dI) {
  switch (nd.kind) {
    case ZNLL_GAUSS: gauss_integral(s[0], s[1], nd.norm_lo, nd.norm_hi, I, dI); return 0;
    case ZNLL_EXP: exp_integral(s[0], nd.norm_lo, nd.norm_hi, nd.shift, I, dI); return 0;
    case ZNLL_CB: cb_integral(s[0], s[1], s[2], s[3], nd.norm_lo, nd.norm_hi, I, dI); return 0;
    case ZNLL_CHEB:
    case ZNLL_LEGENDRE:
    case ZNLL_CHEB2:
    case ZNLL_HERMITE:
    case ZNLL_LAGUERRE:
    case ZNLL_BERNSTEIN: poly_integral(nd.kind, s, nd.degree, nd.norm_lo, nd.norm_hi, nd.space_lo, nd.space_hi, I, dI); return 0;
    case ZNLL_TGAUSS: tgauss_integral(s, nd.norm_lo, nd.norm_hi, I, dI); return 0;
    case ZNLL_GCB: gcb_integral(s, nd.norm_lo, nd.norm_hi, I, dI); return 0;
    case ZNLL_GET: get_integral(s[0], s[1], s[2], nd.norm_lo, nd.norm_hi, I, dI); return 0;
    case ZNLL_GGET: gget_integral(s, nd.norm_lo, nd.norm_hi, I, dI); return 0;
    default: return fail("znll_leaf_integral: node kind %d is not a leaf", nd.kind);
  }
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct NodeInfo {
  znll_node nd;
  int slot_off, const_off;  // this node's own slots / consts
  int parent_sum_frac_slot; // nearest Sum ancestor's frac slot of the branch containing the node, or -1
};

struct Kernel {
  std::string name;
  int origin = 0;  // 1 catalogue, 2 nvrtc
  const ZnllCatalogEntry* aot = nullptr;
  void* cufunc[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // CUfunction, NVRTC path
  void* cufunc_pdf = nullptr;
  void* cufunc_cov[2] = {nullptr, nullptr};
  void* cufunc_hess[2] = {nullptr, nullptr};  // NVRTC path: the one-pass Hessian kernels when the tree has them, else null
  int n_hess = -1;                            // their second-order leaf accumulators (M::NH)
};

struct Channel {
  std::vector<NodeInfo> nodes;
  int n_slots = 0, n_consts = 0, n_acc = 0, acc_off = 0, slot_off = 0;
  Kernel kernel;
  bool bound = false;
  const double* cols[ZNLL_MAX_COLS] = {nullptr};
  const double* w = nullptr;
  int n_cols = 0;
  long long n = 0;
  double sumw = 0.0;
  std::vector<void*> owned;
  void* d_block = nullptr;  // pooled device block: [ticket, 256 B][grid][NACC] partials
  size_t d_block_bytes = 0;
  double* d_partials = nullptr;
  void* d_cov_block = nullptr;  // covariance pass (pooled, on first use): [256 B unused][grid][NACC] partials, [NACC] results
  size_t d_cov_bytes = 0;
  double* d_cov = nullptr;
  size_t d_cov_len = 0;
  unsigned int* d_ticket = nullptr;
  // per-call scratch
  std::vector<double> consts, dlnI;  // dlnI: per slot, d ln I_leaf / d slot (0 for functor slots)
  std::vector<double> slots;
};

struct znll_plan {
  int device = 0, sm_count = 148;
  std::vector<Channel> ch;
  double* d_acc = nullptr;      // pooled device block
  void* h_block = nullptr;      // pooled mapped pinned block: [flag, 64 B][accumulators]
  size_t acc_cap = 0;           // accumulators the two blocks were sized for
  double* h_acc = nullptr;
  int total_acc = 0, total_slots = 0;
  long long launches = 0;
  int max_grid = 0;
  // fused cross-GPU exchange (znll_plan_set_peers)
  int x_rank = 0, x_world = 0;
  double* x_buf[8] = {nullptr};
  unsigned long long x_seq = 0, pending_seq = 0;
  // zero-copy result path: pinned host memory mapped into the device address space
  double* h_acc_dev = nullptr;
  unsigned long long* h_flag = nullptr;
  unsigned long long* h_flag_dev = nullptr;
  bool direct_pending = false;
  // theta-space map (znll_plan_set_theta_map): empty until set
  int n_theta = 0;
  std::vector<double> th_lo, th_hi;
  std::vector<znll_ref> refs;
  std::vector<znll_rule> slot_rules, yield_rules;
  std::vector<double> th_clip, th_slots, th_values, th_gslots;  // per-call scratch
  // sums a host buffer over the ranks of a multi-GPU job (znll_plan_set_reduce_hook); null: single GPU
  znll_reduce_t reduce_fn = nullptr;
  void* reduce_user = nullptr;
};

// ------------------------------------------------------------------------------------------------
// scratch pool.  A plan needs a few small device blocks (block partials + ticket per channel, the accumulators) and one
// block of mapped pinned host memory (result hand-off).  cudaMalloc / cudaHostAlloc / cudaFree* cost 0.1 - 2 ms each,
// synchronise the device and showed outliers of 20 - 200 ms on the benchmark box, i.e. more than the whole bind of a
// loss; so blocks are recycled through a process-wide pool keyed by (device, kind, size class) and only handed back to
// CUDA at process exit.  Contract of recycled blocks: a device block's first 256 bytes are zero (the kernels' ticket
// lives there and every launch leaves it zero), everything else is stale.
// ------------------------------------------------------------------------------------------------
namespace pool {
enum Kind { DEVICE = 0, PINNED = 1, DEVICE_ACC = 2 };  // DEVICE: ticketed blocks (zero head); DEVICE_ACC: accumulators
struct Key {
  int dev, kind;
  size_t bytes;
  bool operator<(const Key& o) const {
    return dev != o.dev ? dev < o.dev : (kind != o.kind ? kind < o.kind : bytes < o.bytes);
  }
};
static std::mutex mu;
static std::map<Key, std::vector<void*>> idle;
static size_t size_class(size_t bytes) {
  size_t r = 4096;
  while (r < bytes) r <<= 1;
  return r;
}
static cudaError_t get(int dev, Kind kind, size_t bytes, void** out) {
  const size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto& v = idle[Key{dev, (int)kind, cls}];
    if (!v.empty()) {
      *out = v.back();
      v.pop_back();
      return cudaSuccess;
    }
  }
  cudaError_t e;
  if (kind != PINNED) {
    e = cudaMalloc(out, cls);
    if (e == cudaSuccess) e = cudaMemset(*out, 0, cls);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the zeroes are in place before any stream may use the block
  } else {
    e = cudaHostAlloc(out, cls, cudaHostAllocMapped);
    if (e == cudaSuccess) memset(*out, 0, cls);
  }
  return e;
}
static void put(int dev, Kind kind, size_t bytes, void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(mu);
  idle[Key{dev, (int)kind, size_class(bytes)}].push_back(p);
}
}  // namespace pool
#define ZNLL_TICKET_BYTES 256  // head of every per-channel device block: the ticket, then the block partials

#define ZNLL_XSTRIDE_HOST 256  // == ZNLL_XSTRIDE of znll_device.cuh: slots of the exchange buffer per (parity, rank)
static unsigned long long g_xseq[16] = {0};  // per device: number of exchanging evaluations this process has launched

static void relayout(znll_plan* p) {
  int a = 0, s = 0;
  for (auto& c : p->ch) {
    c.acc_off = a;
    c.slot_off = s;
    a += c.n_acc;
    s += c.n_slots;
  }
  p->total_acc = a;
  p->total_slots = s;
}

// recursive pre-order parse; returns index after the subtree
static int parse(Channel& c, const znll_node* nodes, int n_nodes, int i, int parent_frac_slot, std::string& name,
                 int depth) {
  if (i >= n_nodes) return -1;
  if (depth > 16) return -1;
  NodeInfo ni;
  ni.nd = nodes[i];
  ni.slot_off = c.n_slots;
  ni.const_off = c.n_consts;
  ni.parent_sum_frac_slot = parent_frac_slot;
  const znll_node& nd = nodes[i];
  char buf[64];
  if (nd.kind == ZNLL_SUM || nd.kind == ZNLL_PROD) {
    if (nd.n_children < 1 || nd.n_children > 16) return -1;
    const bool is_sum = nd.kind == ZNLL_SUM;
    if (is_sum) {
      c.n_slots += nd.n_children;
      c.n_consts += nd.n_children;
    }
    c.nodes.push_back(ni);
    name += is_sum ? "Sum<" : "Prod<";
    int j = i + 1;
    for (int k = 0; k < nd.n_children; ++k) {
      if (k) name += ",";
      j = parse(c, nodes, n_nodes, j, is_sum ? ni.slot_off + k : parent_frac_slot, name, depth + 1);
      if (j < 0) return -1;
    }
    name += ">";
    return j;
  }
  const int ns = leaf_n_slots(nd), nc = leaf_n_consts(nd);
  if (ns < 0 || nd.n_children != 0) return -1;
  if (nd.column < 0 || nd.column >= ZNLL_MAX_COLS) return -1;
  if ((nd.kind == ZNLL_CHEB || nd.kind == ZNLL_LEGENDRE || nd.kind == ZNLL_CHEB2 || nd.kind == ZNLL_BERNSTEIN ||
       nd.kind == ZNLL_HERMITE || nd.kind == ZNLL_LAGUERRE) &&
      (nd.degree < 0 || nd.degree > 16))
    return -1;
  c.n_slots += ns;
  c.n_consts += nc;
  c.nodes.push_back(ni);
  switch (nd.kind) {
    case ZNLL_GAUSS: snprintf(buf, sizeof buf, "Gauss<%d>", nd.column); break;
    case ZNLL_EXP: snprintf(buf, sizeof buf, "Expo<%d>", nd.column); break;
    case ZNLL_CB: snprintf(buf, sizeof buf, "CBall<%d>", nd.column); break;
    case ZNLL_GCB: snprintf(buf, sizeof buf, "GCBall<%d>", nd.column); break;
    case ZNLL_GET: snprintf(buf, sizeof buf, "GETail<%d>", nd.column); break;
    case ZNLL_GGET: snprintf(buf, sizeof buf, "GGETail<%d>", nd.column); break;
    case ZNLL_LEGENDRE: snprintf(buf, sizeof buf, "Legd<%d,%d>", nd.column, nd.degree); break;
    case ZNLL_CHEB2: snprintf(buf, sizeof buf, "Chb2<%d,%d>", nd.column, nd.degree); break;
    case ZNLL_HERMITE: snprintf(buf, sizeof buf, "Herm<%d,%d>", nd.column, nd.degree); break;
    case ZNLL_LAGUERRE: snprintf(buf, sizeof buf, "Lagr<%d,%d>", nd.column, nd.degree); break;
    case ZNLL_BERNSTEIN: snprintf(buf, sizeof buf, "Bern<%d,%d>", nd.column, nd.degree); break;
    case ZNLL_TGAUSS: snprintf(buf, sizeof buf, "TGauss<%d>", nd.column); break;
    default: snprintf(buf, sizeof buf, "Cheb<%d,%d>", nd.column, nd.degree); break;
  }
  name += buf;
  return i + 1;
}

// prologue: slots -> consts (+ d ln I / d slot for the epilogue)
static void fill_consts(Channel& c, const double* slots) {
  c.consts.assign(c.n_consts, 0.0);
  c.dlnI.assign(c.n_slots, 0.0);
  for (auto& ni : c.nodes) {
    const znll_node& nd = ni.nd;
    const double* s = slots + ni.slot_off;
    double* k = c.consts.data() + ni.const_off;
    if (nd.kind == ZNLL_SUM) {
      for (int i = 0; i < nd.n_children; ++i) k[i] = s[i];
      continue;
    }
    if (nd.kind == ZNLL_PROD) continue;
    double I = 0.0, dI[32];
    leaf_integral(nd, s, &I, dI);
    const int ns = leaf_n_slots(nd);
    for (int i = 0; i < ns; ++i) c.dlnI[ni.slot_off + i] = dI[i] / I;
    const double lnI = log(I);  // I <= 0 -> NaN, which must reach the caller (never an error)
    switch (nd.kind) {
      case ZNLL_GAUSS: {
        const double mu = s[0], sigma = s[1];
        k[0] = 1.0 / sigma;
        k[1] = -(mu / sigma);
        k[2] = -(kHalfLog2Pi + log(sigma)) - lnI;
        break;
      }
      case ZNLL_EXP:
        k[0] = s[0];
        k[1] = nd.shift;
        k[2] = -lnI;
        break;
      case ZNLL_CB: {
        const double mu = s[0], sigma = s[1], alpha = s[2], n = s[3];
        const double a = fabs(alpha), sg = alpha > 0 ? 1.0 : (alpha < 0 ? -1.0 : 0.0);
        k[0] = mu;
        k[1] = sg / sigma;
        k[2] = 1.0 / sigma;
        k[3] = a;
        k[4] = n;
        k[5] = n / a - a;
        k[6] = n * log(n / a) - 0.5 * a * a - lnI;
        k[7] = -lnI;
        k[8] = log(n / a) + 1.0;
        k[9] = 1.0 / a;
        k[10] = -n / a - a;
        k[11] = n / (a * a) + 1.0;
        k[12] = sg;
        break;
      }
      case ZNLL_GET: {
        const double a = fabs(s[2]), sg = s[2] > 0 ? 1.0 : (s[2] < 0 ? -1.0 : 0.0);
        k[0] = s[0];
        k[1] = sg / s[1];
        k[2] = 1.0 / s[1];
        k[3] = a;
        k[4] = 0.5 * a * a - lnI;
        k[5] = -lnI;
        k[6] = sg;
        break;
      }
      case ZNLL_GGET: {
        k[0] = s[0];
        k[1] = -lnI;
        for (int side = 0; side < 2; ++side) {
          const double sigma = s[side ? 3 : 1], alpha = s[side ? 4 : 2];
          const double aeff = side ? -alpha : alpha, a = fabs(alpha);
          double* q = k + (side ? 7 : 2);
          q[0] = (aeff > 0 ? 1.0 : (aeff < 0 ? -1.0 : 0.0)) / sigma;
          q[1] = 1.0 / sigma;
          q[2] = a;
          q[3] = 0.5 * a * a - lnI;
          q[4] = alpha > 0 ? 1.0 : (alpha < 0 ? -1.0 : 0.0);
        }
        break;
      }
      case ZNLL_GCB: {
        k[0] = s[0];
        k[1] = -lnI;
        for (int side = 0; side < 2; ++side) {
          const double sigma = s[side ? 4 : 1], alpha = s[side ? 5 : 2], n = s[side ? 6 : 3];
          const double aeff = side ? -alpha : alpha;  // crystalball_func(x, mu, sigmar, -alphar, nr) on the right
          const double a = fabs(alpha);
          double* q = k + (side ? 13 : 2);
          q[0] = (aeff > 0 ? 1.0 : (aeff < 0 ? -1.0 : 0.0)) / sigma;
          q[1] = 1.0 / sigma;
          q[2] = a;
          q[3] = n;
          q[4] = n / a - a;
          q[5] = n * log(n / a) - 0.5 * a * a - lnI;
          q[6] = log(n / a) + 1.0;
          q[7] = 1.0 / a;
          q[8] = -n / a - a;
          q[9] = n / (a * a) + 1.0;
          q[10] = alpha > 0 ? 1.0 : (alpha < 0 ? -1.0 : 0.0);
        }
        break;
      }
      case ZNLL_TGAUSS: {
        const double mu = s[0], sigma = s[1];
        k[0] = 1.0 / sigma;
        k[1] = -(mu / sigma);
        k[2] = -(kHalfLog2Pi + log(sigma)) - lnI;
        k[3] = s[2];
        k[4] = s[3];
        break;
      }
      case ZNLL_BERNSTEIN: {
        k[0] = 1.0 / (nd.space_hi - nd.space_lo);
        k[1] = -nd.space_lo / (nd.space_hi - nd.space_lo);
        k[2] = 1.0 / I;
        for (int i = 0; i <= nd.degree; ++i) k[3 + i] = s[i] / I;
        break;
      }
      case ZNLL_CHEB:
      case ZNLL_LEGENDRE:
      case ZNLL_CHEB2:
      case ZNLL_HERMITE:
      case ZNLL_LAGUERRE: {
        k[0] = 2.0 / (nd.space_hi - nd.space_lo);
        k[1] = -(nd.space_lo + nd.space_hi) / (nd.space_hi - nd.space_lo);
        k[2] = 1.0 / I;
        for (int i = 0; i <= nd.degree; ++i) k[3 + i] = s[i] / I;
        break;
      }
    }
  }
}

// epilogue: raw accumulators -> channel NLL value and d value / d slot
static void finish(const Channel& c, const double* acc, uint32_t flags, double* value, double* grad) {
  const double S = acc[0] - acc[1];
  *value = -S;
  if (!(flags & ZNLL_WANT_GRAD) || !grad) return;
  for (int j = 0; j < c.n_slots; ++j) grad[j] = acc[2 + j];
  for (auto& ni : c.nodes) {
    if (ni.nd.kind == ZNLL_SUM || ni.nd.kind == ZNLL_PROD) continue;
    // K = sum_i w_i * d l_i / d(-ln I_leaf): the leaf's responsibility, read off the nearest Sum
    // ancestor's fraction accumulator, or sum(w) when the leaf is only under products.
    double K;
    if (ni.parent_sum_frac_slot >= 0)
      K = c.slots[ni.parent_sum_frac_slot] * acc[2 + ni.parent_sum_frac_slot];
    else
      K = c.sumw;
    const int ns = leaf_n_slots(ni.nd);
    for (int i = 0; i < ns; ++i) grad[ni.slot_off + i] -= K * c.dlnI[ni.slot_off + i];
  }
  for (int j = 0; j < c.n_slots; ++j) grad[j] = -grad[j];
}

// ------------------------------------------------------------------------------------------------
// NVRTC path: same templates, instantiated at run time for tree shapes outside the catalogue
// ------------------------------------------------------------------------------------------------
extern "C" const char znll_device_src[];  // generated from znll_device.cuh at build time

namespace rtc {
typedef int (*nvrtcCreateProgram_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*nvrtcCompileProgram_t)(void*, int, const char* const*);
typedef int (*nvrtcAddNameExpression_t)(void*, const char*);
typedef int (*nvrtcGetLoweredName_t)(void*, const char*, const char**);
typedef int (*nvrtcGetCUBINSize_t)(void*, size_t*);
typedef int (*nvrtcGetCUBIN_t)(void*, char*);
typedef int (*nvrtcGetProgramLogSize_t)(void*, size_t*);
typedef int (*nvrtcGetProgramLog_t)(void*, char*);
typedef int (*nvrtcDestroyProgram_t)(void**);
typedef int (*cuModuleLoadData_t)(void**, const void*);
typedef int (*cuModuleGetFunction_t)(void**, void*, const char*);
typedef int (*cuLaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*,
                                void**, void**);
typedef int (*nvrtcVersion_t)(int*, int*);
struct Api {
  bool nvrtc_ok = false, ok = false;  // ok: NVRTC and the driver API (module load + launch)
  std::string why;
  nvrtcCreateProgram_t create;
  nvrtcCompileProgram_t compile;
  nvrtcAddNameExpression_t addname;
  nvrtcGetLoweredName_t lowered;
  nvrtcGetCUBINSize_t cubinsize;
  nvrtcGetCUBIN_t cubin;
  nvrtcGetProgramLogSize_t logsize;
  nvrtcGetProgramLog_t log;
  nvrtcDestroyProgram_t destroy;
  nvrtcVersion_t version;
  cuModuleLoadData_t modload;
  cuModuleGetFunction_t getfn;
  cuLaunchKernel_t launch;
};
static Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* hn = nullptr;
    const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (auto n : nv)
      if ((hn = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!hn) {
      a.why = "libnvrtc.so.12 not found";
      return;
    }
#define L(h, field, sym)                                   \
  a.field = (decltype(a.field))dlsym(h, sym);              \
  if (!a.field) {                                          \
    a.why = std::string("missing symbol ") + sym;          \
    return;                                                \
  }
    L(hn, create, "nvrtcCreateProgram")
    L(hn, compile, "nvrtcCompileProgram")
    L(hn, addname, "nvrtcAddNameExpression")
    L(hn, lowered, "nvrtcGetLoweredName")
    L(hn, cubinsize, "nvrtcGetCUBINSize")
    L(hn, cubin, "nvrtcGetCUBIN")
    L(hn, logsize, "nvrtcGetProgramLogSize")
    L(hn, log, "nvrtcGetProgramLog")
    L(hn, destroy, "nvrtcDestroyProgram")
    L(hn, version, "nvrtcVersion")
    a.nvrtc_ok = true;  // compiling needs no driver (the CPU test suite compiles for sm_100a without a GPU)
    void* hc = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!hc) {
      a.why = "libcuda.so.1 not found";
      return;
    }
    L(hc, modload, "cuModuleLoadData")
    L(hc, getfn, "cuModuleGetFunction")
    L(hc, launch, "cuLaunchKernel")
#undef L
    a.ok = true;
  });
  return a;
}

static std::mutex g_mu;
static std::map<std::string, Kernel> g_cache;

// What one NVRTC compilation yields: the cubin and the lowered (mangled) names of the instantiations, in the fixed order
// nll<false,false>, nll<false,true>, nll<true,false>, nll<true,true>, pdf, [cov<false>, cov<true>].
struct Image {
  std::vector<std::string> lowered;
  std::vector<char> cubin;
};

static const char* const kOpts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};

// compile nll_kernel<Model, W, G> for the four (W, G) variants, pdf_kernel<Model> and -- n_slots permitting -- cov_kernel:
// with 2 + NS + (NS+1)(NS+2)/2 > 256 accumulators the covariance pass would not fit the reduction's shared memory
// (znll_eval_cov refuses such trees), and must not break the build.
static int compile_image(const std::string& model, int n_slots, int hess_nh, Image* img) {
  Api& a = api();
  if (!a.nvrtc_ok) return fail("model '%s' is not in the ahead-of-time catalogue and NVRTC is unavailable (%s)",
                               model.c_str(), a.why.c_str());
  std::string src = std::string(znll_device_src) + "\n";
  void* prog = nullptr;
  if (a.create(&prog, src.c_str(), "znll_rtc.cu", 0, nullptr, nullptr)) return fail("nvrtcCreateProgram failed");
  // qualify nested names: insert "znll::" before every identifier that starts a template name
  auto qualify = [](const std::string& m) {
    std::string o;
    size_t i = 0;
    while (i < m.size()) {
      if (isalpha((unsigned char)m[i])) {
        size_t j = i;
        while (j < m.size() && (isalnum((unsigned char)m[j]) || m[j] == '_')) ++j;
        o += "znll::" + m.substr(i, j - i);
        i = j;
      } else {
        o += m[i++];
      }
    }
    return o;
  };
  const std::string q = qualify(model);
  std::vector<std::string> exprs;
  for (int w = 0; w < 2; ++w)
    for (int g = 0; g < 2; ++g)
      exprs.push_back("znll::nll_kernel<" + q + "," + (w ? "true" : "false") + "," + (g ? "true" : "false") + ">");
  exprs.push_back("znll::pdf_kernel<" + q + ">");
  if (2 + n_slots + (n_slots + 1) * (n_slots + 2) / 2 <= 256) {
    exprs.push_back("znll::cov_kernel<" + q + ",false>");
    exprs.push_back("znll::cov_kernel<" + q + ",true>");
  }
  // the one-pass Hessian of a tree the pass knows (hess_nh >= 0: the host's count of M::NH, hess_branches), accumulators permitting
  if (hess_nh >= 0 && 2 + (n_slots + 1) * (n_slots + 2) / 2 + hess_nh <= 256) {
    exprs.push_back("znll::hess_kernel<" + q + ",false>");
    exprs.push_back("znll::hess_kernel<" + q + ",true>");
  }
  for (auto& e : exprs) a.addname(prog, e.c_str());
  const int rc = a.compile(prog, (int)(sizeof(kOpts) / sizeof(kOpts[0])), kOpts);
  if (rc) {
    size_t n = 0;
    a.logsize(prog, &n);
    std::string log(n + 1, 0);
    a.log(prog, &log[0]);
    a.destroy(&prog);
    return fail("NVRTC compile of '%s' failed: %.800s", model.c_str(), log.c_str());
  }
  img->lowered.clear();
  for (auto& e : exprs) {
    const char* low = nullptr;
    a.lowered(prog, e.c_str(), &low);
    if (!low) {
      a.destroy(&prog);
      return fail("nvrtcGetLoweredName failed for '%s'", e.c_str());
    }
    img->lowered.push_back(low);
  }
  size_t sz = 0;
  a.cubinsize(prog, &sz);
  img->cubin.resize(sz);
  a.cubin(prog, img->cubin.data());
  a.destroy(&prog);
  return 0;
}

// ---- on-disk cache of compiled images --------------------------------------------------------------------------------
// A tree outside the catalogue costs 3-8 s of NVRTC per process; the result depends only on the embedded source, the
// tree's name, the options and the NVRTC version, so it is kept under $ZNLL_CACHE_DIR (default $XDG_CACHE_HOME/zfit_b200
// or ~/.cache/zfit_b200; "" or "off" disables).  Any problem with a cache file -- absent, truncated, wrong key or checksum,
// a cubin the driver refuses -- falls back to compiling; files are written to a temporary name and renamed.
static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}

static std::string cache_dir() {
  const char* e = getenv("ZNLL_CACHE_DIR");
  if (e) return (!*e || !strcmp(e, "off") || !strcmp(e, "0")) ? std::string() : std::string(e);
  if (const char* x = getenv("XDG_CACHE_HOME"); x && *x) return std::string(x) + "/zfit_b200";
  if (const char* h = getenv("HOME"); h && *h) return std::string(h) + "/.cache/zfit_b200";
  return std::string();
}

static uint64_t image_key(const std::string& model, int n_slots, int hess_nh) {
  int major = 0, minor = 0;
  Api& a = api();
  if (a.nvrtc_ok) a.version(&major, &minor);
  uint64_t h = fnv1a(znll_device_src, strlen(znll_device_src));
  h = fnv1a(model.data(), model.size(), h);
  for (auto o : kOpts) h = fnv1a(o, strlen(o) + 1, h);
  const int tail[5] = {n_slots, major, minor, 2 /* file format */, hess_nh};
  return fnv1a(tail, sizeof(tail), h);
}

static std::string cache_file(uint64_t key) {
  const std::string dir = cache_dir();
  if (dir.empty()) return dir;
  char name[64];
  snprintf(name, sizeof(name), "/rtc-%016llx.cubin", (unsigned long long)key);
  return dir + name;
}

static const char kMagic[8] = {'Z', 'N', 'L', 'L', 'R', 'T', 'C', '1'};

static bool cache_load(const std::string& path, uint64_t key, Image* img) {
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<char> buf;
  char chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
  fclose(f);
  uint64_t head[3];  // key, payload bytes, payload checksum
  if (buf.size() < sizeof(kMagic) + sizeof(head) || memcmp(buf.data(), kMagic, sizeof(kMagic))) return false;
  memcpy(head, buf.data() + sizeof(kMagic), sizeof(head));
  const char* pay = buf.data() + sizeof(kMagic) + sizeof(head);
  const size_t len = buf.size() - sizeof(kMagic) - sizeof(head);
  if (head[0] != key || head[1] != len || head[2] != fnv1a(pay, len)) return false;
  size_t off = 0;
  auto take = [&](void* dst, size_t k) {
    if (off + k > len) return false;
    memcpy(dst, pay + off, k);
    off += k;
    return true;
  };
  uint32_t n_names = 0;
  if (!take(&n_names, 4) || n_names < 5 || n_names > 9) return false;
  img->lowered.clear();
  for (uint32_t i = 0; i < n_names; ++i) {
    uint32_t l = 0;
    if (!take(&l, 4) || l == 0 || l > 4096 || off + l > len) return false;
    img->lowered.emplace_back(pay + off, l);
    off += l;
  }
  uint64_t sz = 0;
  if (!take(&sz, 8) || sz == 0 || sz != len - off) return false;
  img->cubin.assign(pay + off, pay + off + sz);
  return true;
}

static void mkdirs(const std::string& dir) {
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);  // EEXIST is fine
}

static void cache_store(const std::string& path, uint64_t key, const Image& img) {
  if (path.empty()) return;
  std::vector<char> pay;
  auto put = [&](const void* src, size_t k) { pay.insert(pay.end(), (const char*)src, (const char*)src + k); };
  const uint32_t n_names = (uint32_t)img.lowered.size();
  put(&n_names, 4);
  for (auto& s : img.lowered) {
    const uint32_t l = (uint32_t)s.size();
    put(&l, 4);
    put(s.data(), l);
  }
  const uint64_t sz = img.cubin.size();
  put(&sz, 8);
  put(img.cubin.data(), sz);
  const uint64_t head[3] = {key, (uint64_t)pay.size(), fnv1a(pay.data(), pay.size())};
  mkdirs(path.substr(0, path.rfind('/')));
  char suffix[48];
  snprintf(suffix, sizeof(suffix), ".tmp.%ld", (long)getpid());
  const std::string tmp = path + suffix;
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;  // read-only or missing cache directory: the cache is an optimisation only
  bool ok = fwrite(kMagic, 1, sizeof(kMagic), f) == sizeof(kMagic) && fwrite(head, 1, sizeof(head), f) == sizeof(head) &&
            fwrite(pay.data(), 1, pay.size(), f) == pay.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path.c_str())) remove(tmp.c_str());
}

// from_cache (nullable): 1 when the image came from disk
static int obtain_image(const std::string& model, int n_slots, int hess_nh, Image* img, int* from_cache, bool use_cache = true) {
  const uint64_t key = image_key(model, n_slots, hess_nh);
  const std::string path = use_cache ? cache_file(key) : std::string();
  if (from_cache) *from_cache = 0;
  if (cache_load(path, key, img)) {
    if (from_cache) *from_cache = 1;
    return 0;
  }
  if (compile_image(model, n_slots, hess_nh, img)) return 1;
  cache_store(path, key, *img);
  return 0;
}

static int load_image(const Image& img, const std::string& model, Kernel* k) {
  Api& a = api();
  void* mod = nullptr;
  cudaFree(0);  // make sure the primary context is current
  if (a.modload(&mod, img.cubin.data())) return fail("cuModuleLoadData failed for '%s'", model.c_str());
  k->name = model;
  k->origin = 2;
  // the lowered names are classified by the kernel they instantiate (Itanium mangling: <length><identifier>); within one
  // kernel they keep compile_image's order (nll: weighted x gradient; cov and hess: unweighted, weighted)
  void** nll[4] = {&k->cufunc[0][0], &k->cufunc[0][1], &k->cufunc[1][0], &k->cufunc[1][1]};
  int n_nll = 0, n_pdf = 0, n_cov = 0, n_hess = 0;
  for (auto& name : img.lowered) {
    void** dst = nullptr;
    if (name.find("10nll_kernel") != std::string::npos && n_nll < 4) dst = nll[n_nll++];
    else if (name.find("10pdf_kernel") != std::string::npos && n_pdf < 1) dst = (++n_pdf, &k->cufunc_pdf);
    else if (name.find("10cov_kernel") != std::string::npos && n_cov < 2) dst = &k->cufunc_cov[n_cov++];
    else if (name.find("11hess_kernel") != std::string::npos && n_hess < 2) dst = &k->cufunc_hess[n_hess++];
    else return fail("unexpected kernel '%s' in the image of '%s'", name.c_str(), model.c_str());
    if (a.getfn(dst, mod, name.c_str())) return fail("cuModuleGetFunction failed for '%s'", name.c_str());
  }
  if (n_nll != 4 || n_pdf != 1 || (n_cov != 0 && n_cov != 2) || (n_hess != 0 && n_hess != 2))
    return fail("incomplete image for '%s' (%d/%d/%d/%d kernels)", model.c_str(), n_nll, n_pdf, n_cov, n_hess);
  return 0;
}

static int build(const std::string& model, int n_slots, int hess_nh, Kernel* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_cache.find(model);
  if (it != g_cache.end()) {
    *out = it->second;
    return 0;
  }
  Api& a = api();
  if (!a.ok) return fail("model '%s' is not in the ahead-of-time catalogue and NVRTC is unavailable (%s)",
                         model.c_str(), a.why.c_str());
  Image img;
  int from_cache = 0;
  if (obtain_image(model, n_slots, hess_nh, &img, &from_cache)) return 1;
  Kernel k;
  if (load_image(img, model, &k)) {
    if (!from_cache) return 1;
    // a cached image the driver refuses (e.g. written by another toolkit): drop it and compile
    const uint64_t key = image_key(model, n_slots, hess_nh);
    remove(cache_file(key).c_str());
    k = Kernel();
    if (obtain_image(model, n_slots, hess_nh, &img, nullptr, /*use_cache=*/false) || load_image(img, model, &k)) return 1;
    cache_store(cache_file(key), key, img);
  }
  k.n_hess = k.cufunc_hess[0] ? hess_nh : -1;
  g_cache[model] = k;
  *out = k;
  return 0;
}

static int launch_fn(void* fn, const char* name, const ZnllLaunchArgs& la, const double* consts, int n_consts,
                     int grid, cudaStream_t stream);
static int launch(const Kernel& k, int w, int g, const ZnllLaunchArgs& la, const double* consts, int n_consts,
                  int grid, cudaStream_t stream) {
  return launch_fn(k.cufunc[w][g], k.name.c_str(), la, consts, n_consts, grid, stream);
}
static int launch_fn(void* fn, const char* name, const ZnllLaunchArgs& la, const double* consts, int n_consts,
                     int grid, cudaStream_t stream) {
  // parameter block == znll::KArgs<NC>: the ZnllLaunchArgs prefix followed by NC doubles
  std::vector<char> buf(sizeof(ZnllLaunchArgs) + sizeof(double) * (n_consts > 0 ? n_consts : 1));
  memcpy(buf.data(), &la, sizeof(la));
  memcpy(buf.data() + sizeof(la), consts, sizeof(double) * n_consts);
  void* params[] = {buf.data()};
  const int rc = api().launch(fn, grid, 1, 1, 256, 1, 1, 0, stream, params, nullptr);
  if (rc) return fail("cuLaunchKernel failed with CUresult %d for '%s'", rc, name);
  return 0;
}
}  // namespace rtc

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int znll_abi_version(void) { return ZNLL_ABI_VERSION; }
const char* znll_last_error(void) { return g_err; }

int znll_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int znll_plan_create(int device, int n_channels, znll_plan** out) {
  if (!out || n_channels < 1) return fail("znll_plan_create: bad arguments");
  if (znll_device_count() < 1)
    return fail("znll_plan_create: no CUDA device visible; this library has no CPU fallback");
  CUDA_OK(cudaSetDevice(device));
  znll_plan* p = new znll_plan();
  p->device = device;
  // one attribute, cached per device: cudaGetDeviceProperties queries every property of the board (clocks, PCIe ...) and
  // was seen to take anything from 2 to 600 ms on the benchmark box -- inside every bind of a loss
  static int sm_cached[64] = {0};
  if (device < 0 || device >= 64) {
    delete p;
    return fail("znll_plan_create: bad device %d", device);
  }
  if (sm_cached[device] == 0) {
    int sms = 0;
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    sm_cached[device] = sms;
  }
  p->sm_count = sm_cached[device];
  p->max_grid = p->sm_count * ZNLL_BLOCKS_PER_SM;  // persistent grid: resident CTAs of 256 threads per SM
  p->ch.resize(n_channels);
  *out = p;
  return 0;
}

static void free_channel_data(Channel& c) {
  for (void* q : c.owned) cudaFree(q);
  c.owned.clear();
  c.bound = false;
}

// accumulators on the device and their mapped pinned twin (+ the hand-off flag), sized for the plan's total
static void release_plan_buffers(znll_plan* p) {
  pool::put(p->device, pool::DEVICE_ACC, sizeof(double) * p->acc_cap, p->d_acc);
  pool::put(p->device, pool::PINNED, 64 + sizeof(double) * p->acc_cap, p->h_block);
  p->d_acc = nullptr;
  p->h_block = nullptr;
  p->h_acc = p->h_acc_dev = nullptr;
  p->h_flag = p->h_flag_dev = nullptr;
  p->acc_cap = 0;
}
static int ensure_plan_buffers(znll_plan* p) {
  if (p->d_acc && p->acc_cap >= (size_t)p->total_acc) return 0;
  release_plan_buffers(p);
  const size_t cap = (size_t)(p->total_acc > 0 ? p->total_acc : 1);
  void *d = nullptr, *h = nullptr;
  CUDA_OK(pool::get(p->device, pool::DEVICE_ACC, sizeof(double) * cap, &d));
  CUDA_OK(pool::get(p->device, pool::PINNED, 64 + sizeof(double) * cap, &h));
  p->d_acc = (double*)d;
  p->h_block = h;
  p->acc_cap = cap;
  p->h_flag = (unsigned long long*)h;
  p->h_acc = (double*)((char*)h + 64);
  *p->h_flag = 0ull;
  p->x_seq = 0;
  void* hd = nullptr;
  CUDA_OK(cudaHostGetDevicePointer(&hd, h, 0));
  p->h_flag_dev = (unsigned long long*)hd;
  p->h_acc_dev = (double*)((char*)hd + 64);
  return 0;
}

void znll_plan_destroy(znll_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (auto& c : p->ch) {
    free_channel_data(c);
    pool::put(p->device, pool::DEVICE, c.d_block_bytes, c.d_block);
    pool::put(p->device, pool::DEVICE, c.d_cov_bytes, c.d_cov_block);
  }
  release_plan_buffers(p);
  delete p;
}

// M::NH of the tree as the device header counts it (upper triangles of the leaves with a curved log-density), or -1 when
// the tree is not flat (a leaf, or one Sum / Product of leaves): only then does hess_kernel exist for it
static int leaf_hess_acc(const znll_node& nd) {
  switch (nd.kind) {
    case ZNLL_GAUSS:
    case ZNLL_EXP:
    case ZNLL_CB:
    case ZNLL_TGAUSS:
    case ZNLL_GET:
    case ZNLL_GGET:
    case ZNLL_GCB: {
      const int ns = leaf_n_slots(nd);
      return ns * (ns + 1) / 2;
    }
    default: return 0;  // the polynomial families are linear in their coefficients
  }
}
// A branch of the root as the Hessian pass sees it: one leaf, or (under a Sum root) a Product of leaves.
struct HessBranch {
  std::vector<int> leaves;  // node indices
  int off = 0, ns = 0;      // first slot, number of slots (contiguous in pre-order)
  int nh = 0;               // device accumulators: the leaves' triangles, then the cross blocks of leaf pairs i < j
};
// Splits the tree into the root's branches; false when the Hessian pass does not know the shape (Znll's device traits
// Sum::HESS / Prod::HESS say the same at compile time): a leaf, a Product of leaves, or a Sum of leaves and of Products
// of leaves.
static bool hess_branches(const Channel& c, std::vector<HessBranch>* out) {
  out->clear();
  if (c.nodes.empty()) return false;
  const znll_node& root = c.nodes[0].nd;
  auto is_leaf = [&](size_t i) { return c.nodes[i].nd.kind != ZNLL_SUM && c.nodes[i].nd.kind != ZNLL_PROD; };
  auto add_leaf = [&](HessBranch& b, size_t i) {
    if (b.leaves.empty()) b.off = c.nodes[i].slot_off;
    b.leaves.push_back((int)i);
    b.ns += leaf_n_slots(c.nodes[i].nd);
    b.nh += leaf_hess_acc(c.nodes[i].nd);
  };
  if (is_leaf(0)) {
    if (c.nodes.size() != 1) return false;
    HessBranch b;
    add_leaf(b, 0);
    out->push_back(b);
    return true;
  }
  size_t i = 1;
  for (int k = 0; k < root.n_children; ++k) {
    if (i >= c.nodes.size()) return false;
    HessBranch b;
    if (is_leaf(i)) {
      add_leaf(b, i++);
    } else if (root.kind == ZNLL_SUM && c.nodes[i].nd.kind == ZNLL_PROD) {
      const int m = c.nodes[i].nd.n_children;
      ++i;
      for (int j = 0; j < m; ++j) {
        if (i >= c.nodes.size() || !is_leaf(i)) return false;
        add_leaf(b, i++);
      }
      for (size_t a = 0; a < b.leaves.size(); ++a)
        for (size_t d = a + 1; d < b.leaves.size(); ++d)
          b.nh += leaf_n_slots(c.nodes[b.leaves[a]].nd) * leaf_n_slots(c.nodes[b.leaves[d]].nd);
    } else {
      return false;
    }
    out->push_back(b);
  }
  return i == c.nodes.size();
}
static int tree_hess_nh(const Channel& c) {
  std::vector<HessBranch> br;
  if (!hess_branches(c, &br)) return -1;
  int nh = 0;
  for (auto& b : br) nh += b.nh;
  return nh;
}

int znll_plan_set_model(znll_plan* p, int ch, const znll_node* nodes, int n_nodes) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || !nodes || n_nodes < 1) return fail("znll_plan_set_model: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  Channel& c = p->ch[ch];
  c.nodes.clear();
  c.n_slots = c.n_consts = 0;
  std::string name;
  const int end = parse(c, nodes, n_nodes, 0, -1, name, 0);
  if (end != n_nodes) return fail("znll_plan_set_model: malformed tree (parsed %d of %d nodes)", end, n_nodes);
  c.n_acc = 2 + c.n_slots;
  if (c.n_acc > 250) return fail("znll_plan_set_model: too many slots (%d)", c.n_slots);
  Kernel k;
  if (const ZnllCatalogEntry* e = znll_catalog_find(name.c_str())) {
    if (e->n_slots != c.n_slots || e->n_consts != c.n_consts)
      return fail("catalogue entry '%s' disagrees with the host layout (%d/%d vs %d/%d)", name.c_str(), e->n_slots,
                  e->n_consts, c.n_slots, c.n_consts);
    k.name = name;
    k.origin = 1;
    k.aot = e;
  } else {
    if (rtc::build(name, c.n_slots, tree_hess_nh(c), &k)) return 1;
  }
  c.kernel = k;
  pool::put(p->device, pool::DEVICE, c.d_block_bytes, c.d_block);
  pool::put(p->device, pool::DEVICE, c.d_cov_bytes, c.d_cov_block);
  c.d_block = c.d_cov_block = nullptr;
  c.d_cov = nullptr;
  c.d_cov_len = c.d_cov_bytes = 0;
  c.d_block_bytes = ZNLL_TICKET_BYTES + sizeof(double) * (size_t)p->max_grid * c.n_acc;
  CUDA_OK(pool::get(p->device, pool::DEVICE, c.d_block_bytes, &c.d_block));
  c.d_ticket = (unsigned int*)c.d_block;
  c.d_partials = (double*)((char*)c.d_block + ZNLL_TICKET_BYTES);
  relayout(p);
  return ensure_plan_buffers(p);
}

static int grid_for(const znll_plan* p, long long n) {
  const long long nvec = (n + 1) / 2;  // two events per thread and iteration
  long long g = (nvec + 255) / 256;
  if (g < 1) g = 1;
  if (g > p->max_grid) g = p->max_grid;
  return (int)g;
}

static int compute_sumw(znll_plan* p, Channel& c) {
  if (!c.w) {
    c.sumw = (double)c.n;
    return 0;
  }
  if (c.n == 0) {
    c.sumw = 0.0;
    return 0;
  }
  void* blk = nullptr;  // [ticket, 256 B][2 results][grid][2] partials
  const size_t blk_bytes = ZNLL_TICKET_BYTES + sizeof(double) * 2 * ((size_t)p->max_grid + 1);
  CUDA_OK(pool::get(p->device, pool::DEVICE, blk_bytes, &blk));
  double* d_out = (double*)((char*)blk + ZNLL_TICKET_BYTES);
  double* d_part = d_out + 2;
  const int grid = grid_for(p, c.n * 2);
  cudaError_t e = znll_launch_sumw(c.w, c.n, d_part, (unsigned int*)blk, d_out, grid, 0);
  p->launches++;
  double h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  pool::put(p->device, pool::DEVICE, blk_bytes, blk);
  if (e != cudaSuccess) return fail("sum of weights failed: %s", cudaGetErrorString(e));
  c.sumw = h[0] - h[1];
  return 0;
}

int znll_plan_bind_device(znll_plan* p, int ch, int n_cols, const double* const* cols_dev, const double* w_dev,
                          int64_t n_events) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || n_cols < 1 || n_cols > ZNLL_MAX_COLS || !cols_dev || n_events < 0)
    return fail("znll_plan_bind_device: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  Channel& c = p->ch[ch];
  if (c.nodes.empty()) return fail("znll_plan_bind_device: set the model of channel %d first", ch);
  if (n_events > 8000000000ll)  // the pair loop indexes event pairs with 32 bits
    return fail("znll_plan_bind_device: %lld events in one channel of one GPU exceed the limit of 8e9; shard the data", (long long)n_events);
  for (auto& ni : c.nodes)
    if (ni.nd.kind < ZNLL_SUM && ni.nd.column >= n_cols)
      return fail("znll_plan_bind_device: model reads column %d but only %d bound", ni.nd.column, n_cols);
  free_channel_data(c);
  for (int i = 0; i < ZNLL_MAX_COLS; ++i) c.cols[i] = nullptr;
  for (int i = 0; i < n_cols; ++i) {
    if (n_events > 0 && (!cols_dev[i] || ((uintptr_t)cols_dev[i] & 15)))
      return fail("znll_plan_bind_device: column %d is null or not 16-byte aligned", i);
    c.cols[i] = cols_dev[i];
  }
  if (w_dev && ((uintptr_t)w_dev & 15)) return fail("znll_plan_bind_device: weights not 16-byte aligned");
  c.w = w_dev;
  c.n_cols = n_cols;
  c.n = n_events;
  c.bound = true;
  return compute_sumw(p, c);
}

int znll_plan_bind_host(znll_plan* p, int ch, int n_cols, const double* const* cols_host, const double* w_host,
                        int64_t n_events) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || n_cols < 1 || n_cols > ZNLL_MAX_COLS || !cols_host || n_events < 0)
    return fail("znll_plan_bind_host: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  Channel& c = p->ch[ch];
  free_channel_data(c);
  std::vector<const double*> dev(n_cols);
  std::vector<void*> owned;
  const size_t bytes = sizeof(double) * (size_t)(n_events > 0 ? n_events : 1);
  for (int i = 0; i <= n_cols; ++i) {
    const double* src = i < n_cols ? cols_host[i] : w_host;
    if (i == n_cols && !w_host) break;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e == cudaSuccess && n_events > 0) e = cudaMemcpy(d, src, sizeof(double) * (size_t)n_events, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      for (void* q : owned) cudaFree(q);
      if (d) cudaFree(d);
      return fail("znll_plan_bind_host: %s", cudaGetErrorString(e));
    }
    owned.push_back(d);
    if (i < n_cols) dev[i] = (const double*)d;
  }
  const double* wd = w_host ? (const double*)owned.back() : nullptr;
  const int rc = znll_plan_bind_device(p, ch, n_cols, dev.data(), wd, n_events);
  if (rc) {
    for (void* q : owned) cudaFree(q);
    return rc;
  }
  p->ch[ch].owned = owned;
  return 0;
}

int znll_plan_get_sumw(znll_plan* p, int ch, double* sumw) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || !sumw) return fail("znll_plan_get_sumw: bad arguments");
  *sumw = p->ch[ch].sumw;
  return 0;
}
int znll_plan_set_reduce_hook(znll_plan* p, znll_reduce_t fn, void* user) {
  if (!p) return fail("znll_plan_set_reduce_hook: bad arguments");
  p->reduce_fn = fn;
  p->reduce_user = user;
  return 0;
}
int znll_plan_set_sumw(znll_plan* p, int ch, double sumw) {
  if (!p || ch < 0 || ch >= (int)p->ch.size()) return fail("znll_plan_set_sumw: bad arguments");
  p->ch[ch].sumw = sumw;
  return 0;
}

int znll_plan_set_peers(znll_plan* p, int rank, int world, void* const* peer_bufs, size_t buf_bytes) {
  if (!p) return fail("znll_plan_set_peers: bad arguments");
  if (world <= 1) {
    p->x_world = 0;
    p->x_seq = 0;
    if (p->h_flag) *p->h_flag = 0ull;
    return 0;
  }
  if (world > 8 || rank < 0 || rank >= world || !peer_bufs) return fail("znll_plan_set_peers: bad rank/world (max 8 peers)");
  if (p->total_acc > 256) return fail("znll_plan_set_peers: too many accumulators (%d) for the fused exchange", p->total_acc);
  const size_t need = znll_plan_exchange_bytes(p, world);
  if (buf_bytes < need) return fail("znll_plan_set_peers: exchange buffer too small (%zu < %zu bytes)", buf_bytes, need);
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r] || ((uintptr_t)peer_bufs[r] & 15)) return fail("znll_plan_set_peers: peer buffer %d null or misaligned", r);
    p->x_buf[r] = (double*)peer_bufs[r];
  }
  p->x_rank = rank;
  p->x_world = world;
  // evaluations that exchange are numbered by a process-wide counter (g_xseq): one exchange buffer serves every plan of
  // the process, and all ranks make the same sequence of evaluations (SPMD).  The plan's own counter numbers the
  // evaluations it makes without peers; the host flag of evaluations made before this call must not match again
  p->x_seq = 0;
  if (p->h_flag) *p->h_flag = 0ull;
  return 0;
}

size_t znll_plan_exchange_bytes(const znll_plan* p, int world) {
  if (!p || world < 1) return 0;
  return 128 + sizeof(double) * 2 * (size_t)world * (size_t)ZNLL_XSTRIDE_HOST;
}

int znll_plan_n_slots(const znll_plan* p, int ch) { return (p && ch >= 0 && ch < (int)p->ch.size()) ? p->ch[ch].n_slots : -1; }
int znll_plan_n_acc(const znll_plan* p, int ch) { return (p && ch >= 0 && ch < (int)p->ch.size()) ? p->ch[ch].n_acc : -1; }
int znll_plan_total_slots(const znll_plan* p) { return p ? p->total_slots : -1; }
int znll_plan_total_acc(const znll_plan* p) { return p ? p->total_acc : -1; }
const char* znll_plan_kernel_name(const znll_plan* p, int ch) {
  return (p && ch >= 0 && ch < (int)p->ch.size()) ? p->ch[ch].kernel.name.c_str() : "";
}
int znll_plan_kernel_origin(const znll_plan* p, int ch) {
  return (p && ch >= 0 && ch < (int)p->ch.size()) ? p->ch[ch].kernel.origin : 0;
}
int64_t znll_plan_launch_count(const znll_plan* p) { return p ? p->launches : 0; }
double* znll_plan_acc_device(znll_plan* p) { return p ? p->d_acc : nullptr; }

static int launch_impl(znll_plan* p, const double* slots, uint32_t flags, double log_offset, void* stream, bool direct);

int znll_launch(znll_plan* p, const double* slots, uint32_t flags, double log_offset, void* stream) {
  return launch_impl(p, slots, flags, log_offset, stream, false);
}

static int launch_impl(znll_plan* p, const double* slots, uint32_t flags, double log_offset, void* stream, bool direct) {
  if (!p || !slots) return fail("znll_launch: bad arguments");
  if (direct && (p->total_acc > 256 || p->ch.back().n == 0)) direct = false;
  p->direct_pending = direct;
  CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the (legacy) default stream
  const int g = (flags & ZNLL_WANT_GRAD) ? 1 : 0;
  for (auto& c : p->ch) {
    if (c.nodes.empty() || !c.bound) return fail("znll_launch: channel without model or data");
    c.slots.assign(slots + c.slot_off, slots + c.slot_off + c.n_slots);
    fill_consts(c, c.slots.data());
    if (c.n == 0 && p->x_world <= 1) {  // empty shard: contributes exact zeros
      CUDA_OK(cudaMemsetAsync(p->d_acc + c.acc_off, 0, sizeof(double) * c.n_acc, st));
      continue;
    }
    ZnllLaunchArgs la;
    for (int i = 0; i < ZNLL_MAX_COLS; ++i) la.cols[i] = c.cols[i];
    la.w = c.w;
    la.n = c.n;
    la.log_offset = log_offset;
    la.log_offset2 = 2.0 * log_offset;
    la.partials = c.d_partials;
    la.ticket = c.d_ticket;
    la.out = p->d_acc + c.acc_off;
    memset(&la.xc, 0, sizeof(la.xc));
    if (&c == &p->ch.back() && (p->x_world > 1 || direct)) {
      // the last channel's kernel also performs the cross-GPU exchange and / or the zero-copy hand-off to the host
      la.xc.acc = p->d_acc;
      la.xc.seq = p->pending_seq = (p->x_world > 1) ? ++g_xseq[p->device & 15] : ++p->x_seq;
      la.xc.total = p->total_acc;
      la.xc.stride = ZNLL_XSTRIDE_HOST;
      if (p->x_world > 1) {
        for (int r = 0; r < p->x_world; ++r) la.xc.buf[r] = p->x_buf[r];
        la.xc.rank = p->x_rank;
        la.xc.world = p->x_world;
      }
      if (direct) {
        la.xc.hout = p->h_acc_dev;
        la.xc.hflag = p->h_flag_dev;
      }
    }
    const int w = c.w ? 1 : 0;
    const int grid = grid_for(p, c.n);
    if (c.kernel.origin == 1) {
      cudaError_t e = c.kernel.aot->fn[w][g](la, c.consts.data(), grid, st);
      if (e != cudaSuccess) return fail("kernel launch failed (%s): %s", c.kernel.name.c_str(), cudaGetErrorString(e));
    } else {
      if (rtc::launch(c.kernel, w, g, la, c.consts.data(), c.n_consts, grid, st)) return 1;
    }
    p->launches++;
  }
  return 0;
}

int znll_collect(znll_plan* p, uint32_t flags, double* values, double* grad_slots, void* stream) {
  if (!p || !values) return fail("znll_collect: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the (legacy) default stream
  if (p->direct_pending) {
    // the kernel's last block wrote the accumulators into mapped pinned memory and then the sequence flag
    p->direct_pending = false;
    volatile unsigned long long* flag = p->h_flag;
    unsigned long long spins = 0;
    while (*flag != p->pending_seq) {
      if ((++spins & 0xfffffull) == 0) {  // every ~1M polls: has the stream failed or finished without publishing?
        cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess && *flag != p->pending_seq) return fail("znll_collect: kernel finished without publishing its result");
        if (q != cudaSuccess && q != cudaErrorNotReady) return fail("znll_collect: %s", cudaGetErrorString(q));
      }
    }
    __sync_synchronize();
  } else {
    CUDA_OK(cudaMemcpyAsync(p->h_acc, p->d_acc, sizeof(double) * p->total_acc, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  for (size_t k = 0; k < p->ch.size(); ++k) {
    const Channel& c = p->ch[k];
    finish(c, p->h_acc + c.acc_off, flags, &values[k], grad_slots ? grad_slots + c.slot_off : nullptr);
  }
  return 0;
}

int znll_eval(znll_plan* p, const double* slots, uint32_t flags, double log_offset, double* values,
              double* grad_slots, void* stream) {
  if (launch_impl(p, slots, flags, log_offset, stream, true)) return 1;
  return znll_collect(p, flags, values, grad_slots, stream);
}

int znll_eval_cov(znll_plan* p, const double* slots, double log_offset, double* values, double* grad_slots,
                  double* cov, void* stream) {
  if (!p || !slots || !values || !cov) return fail("znll_eval_cov: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  size_t cov_off = 0;
  for (size_t k = 0; k < p->ch.size(); ++k) {
    Channel& c = p->ch[k];
    if (c.nodes.empty() || !c.bound) return fail("znll_eval_cov: channel without model or data");
    const int NS = c.n_slots, NE = NS + 1, NQ = NE * (NE + 1) / 2, NACC = 2 + NS + NQ;
    if (NACC > 256) return fail("znll_eval_cov: too many slots (%d) for the covariance pass", NS);
    c.slots.assign(slots + c.slot_off, slots + c.slot_off + NS);
    fill_consts(c, c.slots.data());
    std::vector<double> acc(NACC, 0.0);
    if (c.n > 0) {
      const int grid = (int)std::min<long long>(p->sm_count, std::max<long long>(1, ((c.n + 1) / 2 + 255) / 256));
      const size_t need = ((size_t)p->sm_count + 1) * NACC;
      if (c.d_cov_len < need) {
        pool::put(p->device, pool::DEVICE, c.d_cov_bytes, c.d_cov_block);
        c.d_cov_block = nullptr;
        c.d_cov = nullptr;
        c.d_cov_len = 0;
        c.d_cov_bytes = ZNLL_TICKET_BYTES + sizeof(double) * need;
        CUDA_OK(pool::get(p->device, pool::DEVICE, c.d_cov_bytes, &c.d_cov_block));
        c.d_cov = (double*)((char*)c.d_cov_block + ZNLL_TICKET_BYTES);
        c.d_cov_len = need;
      }
      double *d_part = c.d_cov, *d_out = c.d_cov + (size_t)p->sm_count * NACC;
      ZnllLaunchArgs la;
      for (int i = 0; i < ZNLL_MAX_COLS; ++i) la.cols[i] = c.cols[i];
      la.w = c.w;
      la.n = c.n;
      la.log_offset = log_offset;
      la.log_offset2 = 2.0 * log_offset;
    la.log_offset2 = 2.0 * log_offset;
      la.partials = d_part;
      la.ticket = c.d_ticket;
      la.out = d_out;
      memset(&la.xc, 0, sizeof(la.xc));
      const int w = c.w ? 1 : 0;
      int rc = 0;
      if (c.kernel.origin == 1) {
        cudaError_t e = c.kernel.aot->cov[w](la, c.consts.data(), grid, st);
        if (e != cudaSuccess) rc = fail("cov kernel launch failed: %s", cudaGetErrorString(e));
      } else {
        rc = rtc::launch_fn(c.kernel.cufunc_cov[w], c.kernel.name.c_str(), la, c.consts.data(), c.n_consts, grid, st);
      }
      p->launches++;
      cudaError_t e = cudaSuccess;
      if (!rc) e = cudaMemcpyAsync(acc.data(), d_out, sizeof(double) * NACC, cudaMemcpyDeviceToHost, st);
      if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (rc) return rc;
      if (e != cudaSuccess) return fail("znll_eval_cov: %s", cudaGetErrorString(e));
    }
    finish(c, acc.data(), ZNLL_WANT_GRAD, &values[k], grad_slots ? grad_slots + c.slot_off : nullptr);
    // Q = sum e e^T (symmetric), e = [c_0..c_{NS-1}, w];  per-event d log p/d slot = M e  with the same normalisation
    // map as the gradient epilogue:  G_j = c_j - k_leaf * dlnI_j,  k_leaf = f_b * c_{frac_b}  or  w
    std::vector<double> Q((size_t)NE * NE), Mx((size_t)NE * NE, 0.0), T((size_t)NE * NE);
    int q = 2 + NS;
    for (int a = 0; a < NE; ++a)
      for (int b = a; b < NE; ++b) Q[(size_t)a * NE + b] = Q[(size_t)b * NE + a] = acc[q++];
    for (int j = 0; j < NE; ++j) Mx[(size_t)j * NE + j] = 1.0;
    for (auto& ni : c.nodes) {
      if (ni.nd.kind == ZNLL_SUM || ni.nd.kind == ZNLL_PROD) continue;
      const int ns = leaf_n_slots(ni.nd);
      for (int i = 0; i < ns; ++i) {
        const int j = ni.slot_off + i;
        if (ni.parent_sum_frac_slot >= 0)
          Mx[(size_t)j * NE + ni.parent_sum_frac_slot] -= c.dlnI[j] * c.slots[ni.parent_sum_frac_slot];
        else
          Mx[(size_t)j * NE + NS] -= c.dlnI[j];
      }
    }
    for (int a = 0; a < NE; ++a)
      for (int b = 0; b < NE; ++b) {
        double t = 0.0;
        for (int m = 0; m < NE; ++m) t += Mx[(size_t)a * NE + m] * Q[(size_t)m * NE + b];
        T[(size_t)a * NE + b] = t;
      }
    double* out = cov + cov_off;
    for (int a = 0; a < NE; ++a)
      for (int b = 0; b < NE; ++b) {
        double t = 0.0;
        for (int m = 0; m < NE; ++m) t += T[(size_t)a * NE + m] * Mx[(size_t)b * NE + m];
        out[(size_t)a * NE + b] = t;
      }
    cov_off += (size_t)NE * NE;
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------
// exact Hessian in one fused pass (a leaf, a Product of leaves, a Sum of leaves and of Products of leaves)
// ------------------------------------------------------------------------------------------------
// d2 ln I / d slot_a d slot_b of one leaf: central differences of the analytic d ln I (leaf_integral) with one Richardson
// step, O(h^4); the step of every slot is 1e-3 of that slot's natural scale (the width for a location or a width, the
// inverse range for an exponential slope, ...), so neither truncation nor rounding exceeds ~1e-11 relative.
static void leaf_step_scales(const znll_node& nd, const double* s, int ns, double* scale) {
  for (int i = 0; i < ns; ++i) scale[i] = 1.0;
  switch (nd.kind) {
    case ZNLL_GAUSS: scale[0] = scale[1] = fabs(s[1]); break;
    case ZNLL_EXP: scale[0] = 1.0 / std::max(fabs(nd.norm_hi - nd.norm_lo), 1e-300); break;
    case ZNLL_CB:
      scale[0] = scale[1] = fabs(s[1]);
      scale[2] = std::max(fabs(s[2]), 0.1);
      scale[3] = std::max(fabs(s[3]) * 0.5, 0.1);
      break;
    case ZNLL_TGAUSS:
      scale[0] = scale[1] = scale[2] = scale[3] = fabs(s[1]);
      break;
    case ZNLL_GET:
      scale[0] = scale[1] = fabs(s[1]);
      scale[2] = std::max(fabs(s[2]), 0.1);
      break;
    case ZNLL_GGET:
      scale[0] = std::min(fabs(s[1]), fabs(s[3]));
      scale[1] = fabs(s[1]);
      scale[3] = fabs(s[3]);
      scale[2] = std::max(fabs(s[2]), 0.1);
      scale[4] = std::max(fabs(s[4]), 0.1);
      break;
    case ZNLL_GCB:
      scale[0] = std::min(fabs(s[1]), fabs(s[4]));
      scale[1] = fabs(s[1]);
      scale[4] = fabs(s[4]);
      scale[2] = std::max(fabs(s[2]), 0.1);
      scale[5] = std::max(fabs(s[5]), 0.1);
      scale[3] = std::max(fabs(s[3]) * 0.5, 0.1);
      scale[6] = std::max(fabs(s[6]) * 0.5, 0.1);
      break;
    default: break;  // polynomial coefficients: O(1)
  }
}
static void leaf_d2lnI(const znll_node& nd, const double* s, int ns, double* out /* [ns][ns] */) {
  double scale[32], sp[32], dI[32], I = 0.0;
  leaf_step_scales(nd, s, ns, scale);
  std::vector<double> R((size_t)ns * ns, 0.0);
  for (int a = 0; a < ns; ++a) {
    double g[2][32];
    for (int level = 0; level < 2; ++level) {
      const double h = 1e-3 * scale[a] * (level ? 0.5 : 1.0);
      double up[32], dn[32];
      for (int sign = 0; sign < 2; ++sign) {
        for (int i = 0; i < ns; ++i) sp[i] = s[i];
        sp[a] += sign ? -h : h;
        leaf_integral(nd, sp, &I, dI);
        for (int b = 0; b < ns; ++b) (sign ? dn : up)[b] = dI[b] / I;
      }
      for (int b = 0; b < ns; ++b) g[level][b] = (up[b] - dn[b]) / (2.0 * h);
    }
    for (int b = 0; b < ns; ++b) R[(size_t)a * ns + b] = (4.0 * g[1][b] - g[0][b]) / 3.0;
  }
  for (int a = 0; a < ns; ++a)
    for (int b = 0; b < ns; ++b) out[(size_t)a * ns + b] = 0.5 * (R[(size_t)a * ns + b] + R[(size_t)b * ns + a]);
}

// host half of the Hessian pass: raw accumulators [value pair | T | leaf blocks] -> value, d value / d slot (grad_slots,
// this channel's slots, nullable) and d2 value / d slot^2 (out, NS x NS row-major).  c.slots / c.dlnI are current.
static int hess_assemble(Channel& c, const std::vector<double>& acc, int NH, double* value, double* grad_slots, double* out) {
  const int NS = c.n_slots, NE = NS + 1, NQ = NE * (NE + 1) / 2, NACC = 2 + NQ + NH;
  if ((int)acc.size() != NACC) return fail("znll_eval_hess: %d accumulators given, %d expected", (int)acc.size(), NACC);
*value = -(acc[0] - acc[1]);
  // unpack: T = sum w [g;1][g;1]^T, then the leaves' second-order blocks in pre-order
  std::vector<double> T((size_t)NE * NE);
  int q = 2;
  for (int a = 0; a < NE; ++a)
    for (int b = a; b < NE; ++b) T[(size_t)a * NE + b] = T[(size_t)b * NE + a] = acc[q++];
  auto Q = [&](int a, int b) { return T[(size_t)a * NE + b]; };
  auto A = [&](int a) { return T[(size_t)a * NE + NS]; };
  const double W = T[(size_t)NS * NE + NS];
  const bool sum_root = c.nodes[0].nd.kind == ZNLL_SUM;
  struct Leaf {  // a branch of the root: one leaf, or a Product of leaves under a Sum root (one generalised leaf)
    int off, ns, frac;  // first slot, number of slots, fraction slot of the branch under a Sum root (-1 otherwise)
    std::vector<double> B, L2;  // sum r (d2 u / u) (ns x ns, symmetric) and d2 lambda = -d2 ln I (block diagonal per leaf)
  };
  std::vector<HessBranch> branches;
  if (!hess_branches(c, &branches)) return fail("znll_eval_hess: the tree has no one-pass Hessian");
  std::vector<Leaf> leaves;
  int hq = 2 + NQ;
  for (auto& br : branches) {
    Leaf lf;
    lf.off = br.off;
    lf.ns = br.ns;
    lf.frac = sum_root ? c.nodes[br.leaves[0]].parent_sum_frac_slot : -1;
    lf.B.assign((size_t)lf.ns * lf.ns, 0.0);
    lf.L2.assign((size_t)lf.ns * lf.ns, 0.0);
    auto Bx = [&](int a, int b) -> double& { return lf.B[(size_t)a * lf.ns + b]; };
    // the leaves' own triangles (every leaf but the polynomial families, which are linear in their coefficients) ...
    for (int li : br.leaves) {
      const NodeInfo& ni = c.nodes[li];
      const int o = ni.slot_off - br.off, n1 = leaf_n_slots(ni.nd);
      if (leaf_hess_acc(ni.nd) > 0)
        for (int a = 0; a < n1; ++a)
          for (int b = a; b < n1; ++b) Bx(o + a, o + b) = Bx(o + b, o + a) = acc[hq++];
      std::vector<double> l2((size_t)n1 * n1, 0.0);
      leaf_d2lnI(ni.nd, c.slots.data() + ni.slot_off, n1, l2.data());
      for (int a = 0; a < n1; ++a)
        for (int b = 0; b < n1; ++b) lf.L2[(size_t)(o + a) * lf.ns + o + b] = -l2[(size_t)a * n1 + b];
    }
    // ... then the cross blocks of a Product branch: sum r s_a s_b for leaves i < j, row-major
    for (size_t i = 0; i < br.leaves.size(); ++i)
      for (size_t j = i + 1; j < br.leaves.size(); ++j) {
        const NodeInfo &ni = c.nodes[br.leaves[i]], &nj = c.nodes[br.leaves[j]];
        const int oi = ni.slot_off - br.off, oj = nj.slot_off - br.off;
        for (int a = 0; a < leaf_n_slots(ni.nd); ++a)
          for (int b = 0; b < leaf_n_slots(nj.nd); ++b) Bx(oi + a, oj + b) = Bx(oj + b, oi + a) = acc[hq++];
      }
    leaves.push_back(lf);
  }
  if (hq != NACC) return fail("znll_eval_hess: accumulator layout mismatch (%d vs %d)", hq, NACC);
  // d lambda_k / d slot_a = -d ln I_k / d slot_a for a in leaf k
  std::vector<double> H((size_t)NS * NS, 0.0), G(NS, 0.0);
  auto Hx = [&](int a, int b) -> double& { return H[(size_t)a * NS + b]; };
  if (sum_root) {
    for (int a = 0; a < NS; ++a)
      for (int b = 0; b < NS; ++b) Hx(a, b) = -Q(a, b);
    for (auto& lf : leaves) {
      const double fk = c.slots[lf.frac];
      for (int a = 0; a < lf.ns; ++a) {
        for (int b = 0; b < lf.ns; ++b) Hx(lf.off + a, lf.off + b) += lf.B[(size_t)a * lf.ns + b];
        const double cross = A(lf.off + a) / fk;  // (1/P) d2 P / d f_k d slot = g_slot / f_k   (f_k == 0: inf/nan)
        Hx(lf.frac, lf.off + a) += cross;
        Hx(lf.off + a, lf.frac) += cross;
      }
    }
    // the slots' route through the normalisation integrals: lambda_k = -ln I_k enters like ln f_k
    const int NL = (int)leaves.size();
    std::vector<double> R1(NL);
    for (int k2 = 0; k2 < NL; ++k2) R1[k2] = c.slots[leaves[k2].frac] * A(leaves[k2].frac);
    auto group_of = [&](int a) {
      for (int k2 = 0; k2 < NL; ++k2)
        if (a == leaves[k2].frac || (a >= leaves[k2].off && a < leaves[k2].off + leaves[k2].ns)) return k2;
      return -1;
    };
    auto C1 = [&](int a, int k2) {  // sum w d rho_k / d a
      const double fk = c.slots[leaves[k2].frac];
      return (group_of(a) == k2 ? A(a) : 0.0) - fk * Q(leaves[k2].frac, a);
    };
    auto C2 = [&](int k2, int m) {
      const double fk = c.slots[leaves[k2].frac], fm = c.slots[leaves[m].frac];
      return (k2 == m ? R1[k2] : 0.0) - fk * fm * Q(leaves[k2].frac, leaves[m].frac);
    };
    for (int a = 0; a < NS; ++a) G[a] = A(a);
    for (int k2 = 0; k2 < NL; ++k2) {
      const Leaf& lk = leaves[k2];
      for (int b = 0; b < lk.ns; ++b) {
        const double l1b = -c.dlnI[lk.off + b];
        G[lk.off + b] += R1[k2] * l1b;
        for (int a = 0; a < NS; ++a) {
          const double t = C1(a, k2) * l1b;
          Hx(a, lk.off + b) += t;
          Hx(lk.off + b, a) += t;
        }
        for (int b2 = 0; b2 < lk.ns; ++b2) Hx(lk.off + b, lk.off + b2) += R1[k2] * lk.L2[(size_t)b * lk.ns + b2];
      }
      for (int m = 0; m < NL; ++m) {
        const Leaf& lm = leaves[m];
        const double c2 = C2(k2, m);
        for (int a = 0; a < lk.ns; ++a)
          for (int b = 0; b < lm.ns; ++b) Hx(lk.off + a, lm.off + b) += c2 * c.dlnI[lk.off + a] * c.dlnI[lm.off + b];
      }
    }
  } else {  // Product of leaves, or a single leaf: block diagonal
    for (auto& lf : leaves) {
      for (int a = 0; a < lf.ns; ++a) {
        G[lf.off + a] = A(lf.off + a) - W * c.dlnI[lf.off + a];
        for (int b = 0; b < lf.ns; ++b)
          Hx(lf.off + a, lf.off + b) = lf.B[(size_t)a * lf.ns + b] - Q(lf.off + a, lf.off + b) + W * lf.L2[(size_t)a * lf.ns + b];
      }
    }
  }
  // value = -L
  if (grad_slots)
    for (int a = 0; a < NS; ++a) grad_slots[a] = -G[a];
  for (int a = 0; a < NS; ++a)
    for (int b = 0; b < NS; ++b) out[(size_t)a * NS + b] = -0.5 * (Hx(a, b) + Hx(b, a));
  return 0;
}

static bool channel_has_hess(const Channel& c) {
  if (c.kernel.origin == 1) return c.kernel.aot && c.kernel.aot->hess[0] != nullptr;
  return c.kernel.origin == 2 && c.kernel.cufunc_hess[0] != nullptr && c.kernel.n_hess >= 0;
}

int znll_plan_has_hess(const znll_plan* p, int ch) {
  return (p && ch >= 0 && ch < (int)p->ch.size() && channel_has_hess(p->ch[ch])) ? 1 : 0;
}

// values[k] = -sum_i (w_i l_i - offset), grad_slots = d value / d slot, hess = d2 value / d slot_a d slot_b per channel
// ((n_slots)^2, row-major, channels concatenated) -- all in ONE pass over the events per channel.  Returns 3 (and leaves
// the outputs untouched) when a channel's tree is not flat or needs more than the reduction's 256 accumulators: the
// caller then differences the analytic gradient instead.  Catalogue and NVRTC-built trees alike.
int znll_eval_hess(znll_plan* p, const double* slots, double log_offset, double* values, double* grad_slots,
                              double* hess, void* stream) {
  if (!p || !slots || !values || !hess) return fail("znll_eval_hess: bad arguments");
  for (auto& c : p->ch) {
    if (c.nodes.empty() || !c.bound) return fail("znll_eval_hess: channel without model or data");
    if (!channel_has_hess(c)) {
      fail("znll_eval_hess: no one-pass Hessian for the tree '%s' (deeper than a Sum of Products of leaves, or too many accumulators)", c.kernel.name.c_str());
      return 3;
    }
  }
  CUDA_OK(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  size_t h_off = 0;
  for (size_t k = 0; k < p->ch.size(); ++k) {
    Channel& c = p->ch[k];
    const int NS = c.n_slots, NE = NS + 1, NQ = NE * (NE + 1) / 2, NH = c.kernel.origin == 1 ? c.kernel.aot->n_hess : c.kernel.n_hess,
              NACC = 2 + NQ + NH;
    c.slots.assign(slots + c.slot_off, slots + c.slot_off + NS);
    fill_consts(c, c.slots.data());
    std::vector<double> acc(NACC, 0.0);
    if (c.n > 0) {
      const int grid = (int)std::min<long long>(p->sm_count, std::max<long long>(1, ((c.n + 1) / 2 + 255) / 256));
      const size_t need = ((size_t)p->sm_count + 1) * NACC;
      if (c.d_cov_len < need) {
        pool::put(p->device, pool::DEVICE, c.d_cov_bytes, c.d_cov_block);
        c.d_cov_block = nullptr;
        c.d_cov = nullptr;
        c.d_cov_len = 0;
        c.d_cov_bytes = ZNLL_TICKET_BYTES + sizeof(double) * need;
        CUDA_OK(pool::get(p->device, pool::DEVICE, c.d_cov_bytes, &c.d_cov_block));
        c.d_cov = (double*)((char*)c.d_cov_block + ZNLL_TICKET_BYTES);
        c.d_cov_len = need;
      }
      double *d_part = c.d_cov, *d_out = c.d_cov + (size_t)p->sm_count * NACC;
      ZnllLaunchArgs la;
      for (int i = 0; i < ZNLL_MAX_COLS; ++i) la.cols[i] = c.cols[i];
      la.w = c.w;
      la.n = c.n;
      la.log_offset = log_offset;
      la.log_offset2 = 2.0 * log_offset;
      la.partials = d_part;
      la.ticket = c.d_ticket;
      la.out = d_out;
      memset(&la.xc, 0, sizeof(la.xc));
      cudaError_t e = cudaSuccess;
      if (c.kernel.origin == 1) e = c.kernel.aot->hess[c.w ? 1 : 0](la, c.consts.data(), grid, st);
      else if (rtc::launch_fn(c.kernel.cufunc_hess[c.w ? 1 : 0], c.kernel.name.c_str(), la, c.consts.data(), c.n_consts, grid, st))
        return 1;
      p->launches++;
      if (e == cudaSuccess) e = cudaMemcpyAsync(acc.data(), d_out, sizeof(double) * NACC, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return fail("znll_eval_hess: %s", cudaGetErrorString(e));
    }
    if (hess_assemble(c, acc, NH, &values[k], grad_slots ? grad_slots + c.slot_off : nullptr, hess + h_off)) return 1;
    h_off += (size_t)NS * NS;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// theta-space evaluation (include/znll.h): slots from theta, the fused pass, the chain rule back, the extended term
// ------------------------------------------------------------------------------------------------
static inline double ref_value(const znll_plan* p, const znll_ref& r, const double* th) {
  return r.kind == 1 ? th[r.index] : r.value;
}
static double rule_value(const znll_plan* p, const znll_rule& ru, const double* th) {
  const znll_ref* r = p->refs.data() + ru.first;
  switch (ru.kind) {
    case ZNLL_RULE_VALUE: return ref_value(p, r[0], th);
    case ZNLL_RULE_FRACTION: {
      double tot = 0.0;
      for (int k = 1; k < ru.n; ++k) tot += ref_value(p, r[k], th);
      return ref_value(p, r[0], th) / tot;
    }
    case ZNLL_RULE_REMAINDER: {
      double v = 1.0;
      for (int k = 0; k < ru.n; ++k) v -= ref_value(p, r[k], th);
      return v;
    }
    default: {  // ZNLL_RULE_SUM
      double v = 0.0;
      for (int k = 0; k < ru.n; ++k) v += ref_value(p, r[k], th);
      return v;
    }
  }
}
// grad[i] += g * d rule / d theta_i
static void rule_backward(const znll_plan* p, const znll_rule& ru, const double* th, double g, double* grad) {
  const znll_ref* r = p->refs.data() + ru.first;
  auto add = [&](const znll_ref& rf, double d) {
    if (rf.kind == 1) grad[rf.index] += g * d;
  };
  switch (ru.kind) {
    case ZNLL_RULE_VALUE: add(r[0], 1.0); break;
    case ZNLL_RULE_FRACTION: {
      double tot = 0.0;
      for (int k = 1; k < ru.n; ++k) tot += ref_value(p, r[k], th);
      const double a = ref_value(p, r[0], th);
      add(r[0], 1.0 / tot);
      for (int k = 1; k < ru.n; ++k) add(r[k], -a / (tot * tot));
      break;
    }
    case ZNLL_RULE_REMAINDER:
      for (int k = 0; k < ru.n; ++k) add(r[k], -1.0);
      break;
    default:
      for (int k = 0; k < ru.n; ++k) add(r[k], 1.0);
      break;
  }
}

int znll_plan_set_theta_map(znll_plan* p, int n_theta, const double* lower, const double* upper, const znll_ref* refs,
                            int n_refs, const znll_rule* slot_rules, const znll_rule* yield_rules) {
  if (!p || n_theta < 0 || n_refs < 0 || (n_refs > 0 && !refs) || !slot_rules) return fail("znll_plan_set_theta_map: bad arguments");
  auto check = [&](const znll_rule& ru) {
    if (ru.kind < 0 || ru.kind > ZNLL_RULE_SUM || ru.first < 0 || ru.n < 1 || ru.first + ru.n > n_refs) return false;
    if (ru.kind == ZNLL_RULE_FRACTION && ru.n < 2) return false;
    for (int k = 0; k < ru.n; ++k) {
      const znll_ref& r = refs[ru.first + k];
      if (r.kind != 0 && r.kind != 1) return false;
      if (r.kind == 1 && (r.index < 0 || r.index >= n_theta)) return false;
    }
    return true;
  };
  for (int j = 0; j < p->total_slots; ++j)
    if (!check(slot_rules[j])) return fail("znll_plan_set_theta_map: malformed rule of slot %d", j);
  if (yield_rules)
    for (size_t k = 0; k < p->ch.size(); ++k)
      if (!check(yield_rules[k])) return fail("znll_plan_set_theta_map: malformed yield rule of channel %d", (int)k);
  p->n_theta = n_theta;
  p->th_lo.assign(n_theta, -INFINITY);
  p->th_hi.assign(n_theta, INFINITY);
  for (int i = 0; i < n_theta; ++i) {
    if (lower && lower[i] == lower[i]) p->th_lo[i] = lower[i];
    if (upper && upper[i] == upper[i]) p->th_hi[i] = upper[i];
  }
  p->refs.assign(refs, refs + n_refs);
  p->slot_rules.assign(slot_rules, slot_rules + p->total_slots);
  p->yield_rules.clear();
  if (yield_rules) p->yield_rules.assign(yield_rules, yield_rules + p->ch.size());
  p->th_clip.assign(n_theta, 0.0);
  p->th_slots.assign(p->total_slots, 0.0);
  p->th_values.assign(p->ch.size(), 0.0);
  p->th_gslots.assign(p->total_slots, 0.0);
  return 0;
}

int znll_eval_theta(znll_plan* p, const double* theta, uint32_t flags, double log_offset, double* value,
                    double* grad_theta, void* stream) {
  if (!p || !theta || !value) return fail("znll_eval_theta: bad arguments");
  if ((int)p->slot_rules.size() != p->total_slots || p->total_slots == 0) return fail("znll_eval_theta: no theta map set");
  const bool want_grad = (flags & ZNLL_WANT_GRAD) && grad_theta;
  const bool full = (flags & ZNLL_FULL) != 0;
  double* th = p->th_clip.data();
  for (int i = 0; i < p->n_theta; ++i) th[i] = std::min(std::max(theta[i], p->th_lo[i]), p->th_hi[i]);  // NaN passes through
  for (int j = 0; j < p->total_slots; ++j) p->th_slots[j] = rule_value(p, p->slot_rules[j], th);
  const uint32_t f = want_grad ? ZNLL_WANT_GRAD : 0u;
  if (launch_impl(p, p->th_slots.data(), f, full ? 0.0 : log_offset, stream, true)) return 1;
  if (znll_collect(p, f, p->th_values.data(), want_grad ? p->th_gslots.data() : nullptr, stream)) return 1;
  double total = 0.0;
  for (double v : p->th_values) total += v;
  if (want_grad) {
    for (int i = 0; i < p->n_theta; ++i) grad_theta[i] = 0.0;
    for (int j = 0; j < p->total_slots; ++j) rule_backward(p, p->slot_rules[j], th, p->th_gslots[j], grad_theta);
  }
  // tf.nn.log_poisson_loss(samplesize, log(yield), compute_full_loss = full) per channel, + log_offset when an offset
  // is in use (core/loss.py:1301-1317)
  for (size_t k = 0; k < p->yield_rules.size(); ++k) {
    const double t = p->ch[k].sumw;
    const double y = rule_value(p, p->yield_rules[k], th);
    const double log_y = y > 0 ? log(y) : (y == 0 ? -INFINITY : NAN);
    double term = exp(log_y) - log_y * t;
    if (full) {
      if (!(t >= 0.0 && t <= 1.0)) term += t * log(t) - t + 0.5 * log(2.0 * M_PI * t);
    } else {
      term += log_offset;
    }
    total += term;
    if (want_grad) rule_backward(p, p->yield_rules[k], th, y != 0 ? 1.0 - t / y : NAN, grad_theta);
  }
  *value = total;
  return 0;
}

// information matrix in theta space (the MIGRAD metric seed): sum over channels of B^T E B with E from znll_eval_cov
// (rescaled by sum w / sum w^2 for weighted channels) and B = d [slots; log yield] / d theta -- loss.information_matrix
static int theta_information(znll_plan* p, const double* theta, double* info /* n_theta^2 */, void* stream) {
  const int n = p->n_theta;
  double* th = p->th_clip.data();
  for (int i = 0; i < n; ++i) th[i] = std::min(std::max(theta[i], p->th_lo[i]), p->th_hi[i]);
  for (int j = 0; j < p->total_slots; ++j) p->th_slots[j] = rule_value(p, p->slot_rules[j], th);
  size_t cov_len = 0;
  for (auto& c : p->ch) cov_len += (size_t)(c.n_slots + 1) * (c.n_slots + 1);
  std::vector<double> cov(cov_len), vals(p->ch.size()), gs(p->total_slots);
  if (znll_eval_cov(p, p->th_slots.data(), 0.0, vals.data(), gs.data(), cov.data(), stream)) return 1;
  // one process per GPU: the pass saw this rank's shard; the matrix is linear in the events
  if (p->reduce_fn && p->reduce_fn(p->reduce_user, cov.data(), (int)cov.size())) return fail("theta_information: the reduce hook failed");
  for (int i = 0; i < n * n; ++i) info[i] = 0.0;
  size_t off = 0;
  for (size_t k = 0; k < p->ch.size(); ++k) {
    const Channel& c = p->ch[k];
    const int ns = c.n_slots, ne = ns + 1;
    std::vector<double> B((size_t)ne * n, 0.0);
    for (int j = 0; j < ns; ++j) rule_backward(p, p->slot_rules[c.slot_off + j], th, 1.0, B.data() + (size_t)j * n);
    if (!p->yield_rules.empty()) {
      const double y = rule_value(p, p->yield_rules[k], th);
      rule_backward(p, p->yield_rules[k], th, 1.0 / y, B.data() + (size_t)ns * n);
    }
    const double* E = cov.data() + off;
    double scale = 1.0;
    if (c.w && E[(size_t)ns * ne + ns] > 0) scale = c.sumw / E[(size_t)ns * ne + ns];
    std::vector<double> EB((size_t)ne * n, 0.0);
    for (int a = 0; a < ne; ++a)
      for (int b = 0; b < ne; ++b) {
        const double e = E[(size_t)a * ne + b] * scale;
        if (e == 0.0) continue;
        for (int i = 0; i < n; ++i) EB[(size_t)a * n + i] += e * B[(size_t)b * n + i];
      }
    for (int a = 0; a < ne; ++a)
      for (int i = 0; i < n; ++i) {
        const double bi = B[(size_t)a * n + i];
        if (bi == 0.0) continue;
        for (int j = 0; j < n; ++j) info[(size_t)i * n + j] += bi * EB[(size_t)a * n + j];
      }
    off += (size_t)ne * ne;
  }
  return 0;
}

struct ThetaFit {
  znll_plan* plan;
  std::vector<double> theta;      // full vector; the fitted entries are overwritten per call
  std::vector<double> grad_full, info_full;
  std::vector<int32_t> index;     // fit parameter -> theta entry
  uint32_t flags;
  double log_offset;
  void* stream;
  bool nonfinite = false;
};
static int theta_fcn(void* user, const double* x, int what, double* value, double* grad) {
  ThetaFit* f = (ThetaFit*)user;
  for (size_t i = 0; i < f->index.size(); ++i) f->theta[f->index[i]] = x[i];
  double v = 0.0;
  const bool g = (what & ZNLL_FCN_GRADIENT) != 0;
  if (znll_eval_theta(f->plan, f->theta.data(), f->flags | (g ? ZNLL_WANT_GRAD : 0u), f->log_offset, &v,
                      g ? f->grad_full.data() : nullptr, f->stream))
    return 1;
  if (!(v == v) || v == INFINITY || v == -INFINITY) {
    f->nonfinite = true;
    return 1;
  }
  if (what & ZNLL_FCN_VALUE) *value = v;
  if (g)
    for (size_t i = 0; i < f->index.size(); ++i) grad[i] = f->grad_full[f->index[i]];
  return 0;
}
static int theta_info(void* user, const double* x, double* info) {
  ThetaFit* f = (ThetaFit*)user;
  for (size_t i = 0; i < f->index.size(); ++i) f->theta[f->index[i]] = x[i];
  if (theta_information(f->plan, f->theta.data(), f->info_full.data(), f->stream)) return 1;
  const int n = (int)f->index.size(), nt = f->plan->n_theta;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double v = f->info_full[(size_t)f->index[i] * nt + f->index[j]];
      if (!(v == v)) return 1;
      info[(size_t)i * n + j] = v;
    }
  return 0;
}

int znll_migrad_theta(znll_plan* p, int n_fit, const int32_t* fit_index, const double* theta0, const double* steps,
                      const double* lower, const double* upper, const znll_migrad_options* options, uint32_t flags,
                      double log_offset, int use_information, double* theta_out, double* cov_out, double* grad_out,
                      znll_migrad_result* result, void* stream) {
  if (!p || n_fit < 1 || !fit_index || !theta0 || !steps || !options || !result) return fail("znll_migrad_theta: bad arguments");
  if ((int)p->slot_rules.size() != p->total_slots || p->total_slots == 0) return fail("znll_migrad_theta: no theta map set");
  ThetaFit f;
  f.plan = p;
  f.theta.assign(theta0, theta0 + p->n_theta);
  f.grad_full.assign(p->n_theta, 0.0);
  f.info_full.assign((size_t)p->n_theta * p->n_theta, 0.0);
  f.index.assign(fit_index, fit_index + n_fit);
  for (int i = 0; i < n_fit; ++i)
    if (fit_index[i] < 0 || fit_index[i] >= p->n_theta) return fail("znll_migrad_theta: fit index %d out of range", fit_index[i]);
  f.flags = flags & ZNLL_FULL;
  f.log_offset = log_offset;
  f.stream = stream;
  std::vector<double> x0(n_fit), x(n_fit);
  for (int i = 0; i < n_fit; ++i) x0[i] = theta0[fit_index[i]];
  const int rc = znll_migrad(theta_fcn, use_information ? theta_info : nullptr, &f, n_fit, x0.data(), steps, lower, upper,
                             options, x.data(), cov_out, grad_out, result);
  if (theta_out) {
    for (int i = 0; i < p->n_theta; ++i) theta_out[i] = theta0[i];
    for (int i = 0; i < n_fit; ++i) theta_out[fit_index[i]] = x[i];
  }
  if (rc && f.nonfinite) return 2;
  return rc;
}

// launches pdf_kernel of channel ch into a device buffer of c.n doubles (no synchronisation)
static int pdf_launch(znll_plan* p, int ch, const double* slots, double* d_out, cudaStream_t st) {
  Channel& c = p->ch[ch];
  c.slots.assign(slots, slots + c.n_slots);
  fill_consts(c, c.slots.data());
  ZnllLaunchArgs la;
  for (int i = 0; i < ZNLL_MAX_COLS; ++i) la.cols[i] = c.cols[i];
  la.w = c.w;
  la.n = c.n;
  la.log_offset = 0.0;
  la.log_offset2 = 0.0;
  la.partials = d_out;
  la.ticket = c.d_ticket;
  la.out = nullptr;
  memset(&la.xc, 0, sizeof(la.xc));
  long long g = (c.n + 255) / 256;
  if (g > p->max_grid * 4) g = p->max_grid * 4;
  int rc = 0;
  if (c.kernel.origin == 1) {
    cudaError_t e = c.kernel.aot->pdf(la, c.consts.data(), (int)g, st);
    if (e != cudaSuccess) rc = fail("pdf kernel launch failed: %s", cudaGetErrorString(e));
  } else {
    rc = rtc::launch_fn(c.kernel.cufunc_pdf, c.kernel.name.c_str(), la, c.consts.data(), c.n_consts, (int)g, st);
  }
  p->launches++;
  return rc;
}

int znll_pdf(znll_plan* p, int ch, const double* slots, double* out_host, void* stream) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || !slots || !out_host) return fail("znll_pdf: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  Channel& c = p->ch[ch];
  if (c.nodes.empty() || !c.bound) return fail("znll_pdf: channel without model or data");
  if (c.n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the (legacy) default stream
  double* d_out = nullptr;
  CUDA_OK(cudaMalloc(&d_out, sizeof(double) * (size_t)c.n));
  const int rc = pdf_launch(p, ch, slots, d_out, st);
  cudaError_t e = cudaSuccess;
  if (!rc) e = cudaMemcpyAsync(out_host, d_out, sizeof(double) * (size_t)c.n, cudaMemcpyDeviceToHost, st);
  if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_out);
  if (rc) return rc;
  if (e != cudaSuccess) return fail("znll_pdf: %s", cudaGetErrorString(e));
  return 0;
}

int znll_pdf_device(znll_plan* p, int ch, const double* slots, double* out_dev, void* stream) {
  if (!p || ch < 0 || ch >= (int)p->ch.size() || !slots || !out_dev) return fail("znll_pdf_device: bad arguments");
  CUDA_OK(cudaSetDevice(p->device));
  Channel& c = p->ch[ch];
  if (c.nodes.empty() || !c.bound) return fail("znll_pdf_device: channel without model or data");
  if (c.n == 0) return 0;
  return pdf_launch(p, ch, slots, out_dev, (cudaStream_t)stream);
}

int znll_leaf_integral(const znll_node* leaf, const double* slots, double* integral, double* dintegral) {
  if (!leaf || !slots || !integral || !dintegral) return fail("znll_leaf_integral: bad arguments");
  if (leaf_n_slots(*leaf) < 0) return fail("znll_leaf_integral: node kind %d is not a leaf", leaf->kind);
  return leaf_integral(*leaf, slots, integral, dintegral);
}

// host-only test hooks: the prologue and the epilogue that znll_eval runs around the kernel, reachable without a device
static int test_parse(Channel& c, const znll_node* nodes, int n_nodes, const double* slots, std::string* name) {
  std::string nm;
  const int end = parse(c, nodes, n_nodes, 0, -1, nm, 0);
  if (end != n_nodes) return fail("malformed tree (parsed %d of %d nodes)", end, n_nodes);
  c.n_acc = 2 + c.n_slots;
  c.slots.assign(slots, slots + c.n_slots);
  fill_consts(c, c.slots.data());
  if (name) *name = nm;
  return 0;
}

int znll_test_prologue(const znll_node* nodes, int n_nodes, const double* slots, double* consts, int max_consts,
                       int* n_consts, int* n_slots, char* kernel_name, int name_len) {
  if (!nodes || n_nodes < 1 || !slots || !consts || !n_consts || !n_slots) return fail("znll_test_prologue: bad arguments");
  Channel c;
  std::string name;
  if (test_parse(c, nodes, n_nodes, slots, &name)) return 1;
  if (c.n_consts > max_consts) return fail("znll_test_prologue: %d consts do not fit %d", c.n_consts, max_consts);
  for (int i = 0; i < c.n_consts; ++i) consts[i] = c.consts[i];
  *n_consts = c.n_consts;
  *n_slots = c.n_slots;
  if (kernel_name && name_len > 0) snprintf(kernel_name, (size_t)name_len, "%s", name.c_str());
  return 0;
}

int znll_test_rtc_image(const char* model, int n_slots, int hess_nh, int use_cache, size_t* cubin_bytes, int* n_names,
                        int* from_cache) {
  if (!model || !*model || n_slots < 0 || !cubin_bytes || !n_names || !from_cache) return fail("znll_test_rtc_image: bad arguments");
  rtc::Image img;
  if (rtc::obtain_image(model, n_slots, hess_nh, &img, from_cache, use_cache != 0)) return 1;
  *cubin_bytes = img.cubin.size();
  *n_names = (int)img.lowered.size();
  return 0;
}

int znll_test_hess_epilogue(const znll_node* nodes, int n_nodes, const double* slots, const double* acc, int n_acc, int n_hess,
                            double* value, double* grad_slots, double* hess) {
  if (!nodes || !slots || !acc || !value || !hess) return fail("znll_test_hess_epilogue: bad arguments");
  Channel c;
  if (test_parse(c, nodes, n_nodes, slots, nullptr)) return 1;
  std::vector<double> a(acc, acc + n_acc);
  return hess_assemble(c, a, n_hess, value, grad_slots, hess);
}

int znll_test_epilogue(const znll_node* nodes, int n_nodes, const double* slots, const double* acc, double sumw,
                       double* value, double* grad_slots) {
  if (!nodes || n_nodes < 1 || !slots || !acc || !value) return fail("znll_test_epilogue: bad arguments");
  Channel c;
  if (test_parse(c, nodes, n_nodes, slots, nullptr)) return 1;
  c.sumw = sumw;
  finish(c, acc, grad_slots ? ZNLL_WANT_GRAD : 0u, value, grad_slots);
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmark (roofline denominator)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
         x6 = x0 + 6e-3, x7 = x0 + 7e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int znll_test_math(int kind, int64_t n, const double* in_host, double* out_host) {
  if (!in_host || !out_host || n < 0) return fail("znll_test_math: bad arguments");
  if (znll_device_count() < 1) return fail("znll_test_math: no CUDA device");
  if (n == 0) return 0;
  double *d_in = nullptr, *d_out = nullptr;
  CUDA_OK(cudaMalloc(&d_in, sizeof(double) * (size_t)n));
  CUDA_OK(cudaMalloc(&d_out, sizeof(double) * (size_t)n));
  cudaError_t e = cudaMemcpy(d_in, in_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = znll_launch_test_math(kind, d_in, d_out, (long long)n, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail("znll_test_math: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int znll_fp64_peak(int device, double ms, double* tflops) {
  if (!tflops) return fail("znll_fp64_peak: bad arguments");
  if (znll_device_count() < 1) return fail("znll_fp64_peak: no CUDA device");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 2 * 4, block = 256;
  double* d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(double) * (size_t)grid * block));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int iters = 2000;
  dfma_kernel<<<grid, block>>>(d, 200, 0.999999, 1e-7);  // warm-up
  float t = 0.f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    CUDA_OK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&t, e0, e1);
    if (t >= ms || rep == 5) break;
    iters = (int)(iters * (ms / (t > 1e-3f ? t : 1e-3f)) * 1.05) + 1;
  }
  const double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * block;
  *tflops = flops / (t * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}
