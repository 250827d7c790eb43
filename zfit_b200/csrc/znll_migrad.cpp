// Native MIGRAD-protocol driver.
//
// The reference drives the loss with iminuit's C++ Minuit2 (src/zfit/minimizers/minimizer_minuit.py:175-319); that
// library is not available here, so this file implements the same protocol natively: Minuit's parameter
// transformations, a variable-metric (DFP, DavidonErrorUpdator) iteration with a line search, the EDM stop rule with
// Minuit2's Dcovar bookkeeping, HESSE from differences of the analytic gradient, the driver's own numerical gradient
// when the caller has none -- plus one thing Minuit does not have: the metric can be seeded and refreshed from the
// loss's information matrix (one fused pass on the GPU).  zfit_b200/_migrad.py is the same algorithm in Python (kept as
// the executable specification and cross-check); this file removes ~85 us of interpreter time per pass.
//
// The driver sees the objective only through two C callbacks, so it is independent of CUDA and of Python.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/znll.h"

namespace {

typedef std::vector<double> Vec;
typedef std::vector<double> Mat;  // row-major n x n

const double kFar = 0.5;            // second line-search probe only when the secant minimum is this far from lambda = 1
const double kRefreshEdm = 1.0;     // refresh the metric from the information matrix below this inflated EDM
const double kRefreshDcovar = 0.05; // relative uncertainty credited to the refreshed metric (strategy-1 HESSE threshold)

struct Abort {};  // a callback reported failure: unwind to the entry point

inline bool finite(double v) { return std::isfinite(v); }
bool all_finite(const Vec& v) {
  for (double x : v)
    if (!finite(x)) return false;
  return true;
}
double dot(const Vec& a, const Vec& b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
Vec matvec(const Mat& A, const Vec& x) {
  const size_t n = x.size();
  Vec y(n, 0.0);
  for (size_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (size_t j = 0; j < n; ++j) s += A[i * n + j] * x[j];
    y[i] = s;
  }
  return y;
}

// eigenvalues of a symmetric matrix (cyclic Jacobi; n is the number of fit parameters, i.e. small)
Vec sym_eigenvalues(Mat A, int n) {
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
      }
  }
  Vec w(n);
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
  std::sort(w.begin(), w.end());
  return w;
}

// inverse of a symmetric positive definite matrix via Cholesky; false if not positive definite
bool spd_inverse(const Mat& A, int n, Mat* out) {
  Mat L(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0.0) || !finite(s)) return false;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  Mat Li(n * n, 0.0);  // inverse of L (lower triangular)
  for (int i = 0; i < n; ++i) {
    Li[i * n + i] = 1.0 / L[i * n + i];
    for (int j = 0; j < i; ++j) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= L[i * n + k] * Li[k * n + j];
      Li[i * n + j] = s / L[i * n + i];
    }
  }
  out->assign(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
      for (int k = i; k < n; ++k) s += Li[k * n + i] * Li[k * n + j];
      (*out)[i * n + j] = (*out)[j * n + i] = s;
    }
  return true;
}

// Minuit2 SinParameterTransformation / SqrtLow/UpParameterTransformation
struct Transform {
  int n;
  Vec lo, up;
  std::vector<int> kind;  // 0 free, 1 both limits, 2 lower only, 3 upper only
  Transform(int n_, const double* lower, const double* upper) : n(n_), lo(n_), up(n_), kind(n_) {
    for (int i = 0; i < n; ++i) {
      lo[i] = (lower && finite(lower[i])) ? lower[i] : -std::numeric_limits<double>::infinity();
      up[i] = (upper && finite(upper[i])) ? upper[i] : std::numeric_limits<double>::infinity();
      const bool l = finite(lo[i]), u = finite(up[i]);
      kind[i] = l && u ? 1 : (l ? 2 : (u ? 3 : 0));
    }
  }
  Vec to_internal(const Vec& x) const {
    Vec y(x);
    for (int i = 0; i < n; ++i) {
      if (kind[i] == 1) {
        double yy = 2.0 * (x[i] - lo[i]) / (up[i] - lo[i]) - 1.0;
        yy = std::min(1.0, std::max(-1.0, yy));
        y[i] = std::asin(yy);
      } else if (kind[i] == 2) {
        const double t = x[i] - lo[i] + 1.0;
        y[i] = std::sqrt(std::max(t * t - 1.0, 0.0));
      } else if (kind[i] == 3) {
        const double t = up[i] - x[i] + 1.0;
        y[i] = std::sqrt(std::max(t * t - 1.0, 0.0));
      }
    }
    return y;
  }
  Vec to_external(const Vec& y) const {
    Vec x(y);
    for (int i = 0; i < n; ++i) {
      if (kind[i] == 1)
        x[i] = lo[i] + 0.5 * (up[i] - lo[i]) * (std::sin(y[i]) + 1.0);
      else if (kind[i] == 2)
        x[i] = lo[i] - 1.0 + std::sqrt(y[i] * y[i] + 1.0);
      else if (kind[i] == 3)
        x[i] = up[i] + 1.0 - std::sqrt(y[i] * y[i] + 1.0);
      x[i] = std::min(std::max(x[i], lo[i]), up[i]);  // rounding must not step an ulp outside the limits
    }
    return x;
  }
  Vec dext_dint(const Vec& y) const {
    Vec d(n, 1.0);
    for (int i = 0; i < n; ++i) {
      if (kind[i] == 1)
        d[i] = 0.5 * (up[i] - lo[i]) * std::cos(y[i]);
      else if (kind[i] == 2)
        d[i] = y[i] / std::sqrt(y[i] * y[i] + 1.0);
      else if (kind[i] == 3)
        d[i] = -y[i] / std::sqrt(y[i] * y[i] + 1.0);
    }
    return d;
  }
};

struct Driver {
  znll_fcn_t fcn;
  znll_info_t info;
  void* user;
  int n;
  Transform tr;
  bool has_grad, use_fg, hesse;
  double edm_goal;
  long long maxfcn;
  long long nf = 0, ng = 0;
  Vec ysteps, num_steps;

  Driver(znll_fcn_t f, znll_info_t i, void* u, int n_, const double* lo, const double* up, const znll_migrad_options& o)
      : fcn(f), info(i), user(u), n(n_), tr(n_, lo, up), has_grad(o.has_gradient != 0),
        use_fg(o.has_gradient != 0 && o.has_value_gradient != 0), hesse(o.hesse != 0), edm_goal(o.edm_goal),
        maxfcn(o.maxfcn > 0 ? o.maxfcn : 10000) {}

  double f_int(const Vec& y) {
    ++nf;
    const Vec x = tr.to_external(y);
    double v = std::numeric_limits<double>::quiet_NaN();
    if (fcn(user, x.data(), ZNLL_FCN_VALUE, &v, nullptr)) throw Abort();
    return v;
  }
  Vec num_grad(const Vec& y, double fy) {  // MnNumericalGradientCalculator, simplified: central differences, adaptive steps
    Vec g(n, 0.0);
    for (int i = 0; i < n; ++i) {
      const double h = num_steps[i];
      Vec yp(y), ym(y);
      yp[i] += h;
      ym[i] -= h;
      const double fp = f_int(yp), fm = f_int(ym);
      g[i] = (fp - fm) / (2 * h);
      const double g2 = (fp + fm - 2 * fy) / (h * h);
      if (finite(g2) && std::fabs(g2) > 0) {
        const double hopt = std::sqrt(8.0 * 1e-11 * (std::fabs(fy) + 1.0) / std::fabs(g2));
        num_steps[i] = std::min(std::max(hopt, 1e-9 * std::max(1.0, std::fabs(y[i]))), 10 * h);
      }
    }
    return g;
  }
  Vec g_int(const Vec& y, const double* fy = nullptr) {
    if (has_grad) {
      ++ng;
      const Vec x = tr.to_external(y);
      Vec g(n, std::numeric_limits<double>::quiet_NaN());
      double v = 0.0;
      if (fcn(user, x.data(), ZNLL_FCN_GRADIENT, &v, g.data())) throw Abort();
      const Vec J = tr.dext_dint(y);
      for (int i = 0; i < n; ++i) g[i] *= J[i];
      return g;
    }
    const double f0 = fy ? *fy : f_int(y);
    return num_grad(y, f0);
  }
  void fg_int(const Vec& y, double* f, Vec* g) {
    ++nf;
    ++ng;
    const Vec x = tr.to_external(y);
    g->assign(n, std::numeric_limits<double>::quiet_NaN());
    *f = std::numeric_limits<double>::quiet_NaN();
    if (fcn(user, x.data(), ZNLL_FCN_VALUE | ZNLL_FCN_GRADIENT, f, g->data())) throw Abort();
    const Vec J = tr.dext_dint(y);
    for (int i = 0; i < n; ++i) (*g)[i] *= J[i];
  }

  // inverse of the information matrix in internal coordinates; false when it cannot serve as a metric
  bool information_metric(const Vec& y, Mat* V) {
    if (!info) return false;
    const Vec x = tr.to_external(y);
    Mat H(n * n, std::numeric_limits<double>::quiet_NaN());
    if (info(user, x.data(), H.data())) return false;  // unavailable: fall back to Minuit's own seed
    for (double v : H)
      if (!finite(v)) return false;
    const Vec J = tr.dext_dint(y);
    Mat S(n * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) S[i * n + j] = 0.5 * (H[i * n + j] + H[j * n + i]) * J[i] * J[j];
    Vec d(n);
    for (int i = 0; i < n; ++i) d[i] = std::sqrt(std::max(S[i * n + i], 1e-300));
    Mat C(n * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) C[i * n + j] = S[i * n + j] / d[i] / d[j];
    const Vec w = sym_eigenvalues(C, n);
    if (!(w[0] >= 1e-10 * w[n - 1]) || !(w[0] > 0)) return false;
    Mat Ci;
    if (!spd_inverse(C, n, &Ci)) return false;
    V->assign(n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (*V)[i * n + j] = Ci[i * n + j] / d[i] / d[j];
    return true;
  }

  // MnSeedGenerator: diagonal metric from second derivatives along each axis (or the information matrix, first time)
  Mat seed_metric(const Vec& y, double f, const Vec& g, bool first) {
    Mat V;
    if (first && information_metric(y, &V)) return V;
    V.assign(n * n, 0.0);
    for (int i = 0; i < n; ++i) {
      const double h = std::max(ysteps[i] * 0.05, 1e-7);
      Vec yp(y);
      yp[i] += h;
      double g2;
      if (has_grad) {
        g2 = (g_int(yp)[i] - g[i]) / h;
      } else {
        Vec ym(y);
        ym[i] -= h;
        g2 = (f_int(yp) + f_int(ym) - 2 * f) / (h * h);
      }
      if (!finite(g2) || g2 <= 0) g2 = 1.0 / std::max(ysteps[i] * ysteps[i], 1e-16);
      V[i * n + i] = 1.0 / g2;
    }
    return V;
  }

  struct Probe {
    double lam, f;
    Vec g;
  };

  // parabolic line search with value probes (MnLineSearch, simplified): lambda = 0 means "no improvement"
  void line_search(const Vec& y, const Vec& step, double f0, double gdel, double* lam_out, double* f_out) {
    std::vector<Probe> pts;
    auto probe = [&](double lam) {
      Vec yy(y);
      for (int i = 0; i < n; ++i) yy[i] += lam * step[i];
      const double fl = f_int(yy);
      if (finite(fl)) pts.push_back({lam, fl, Vec()});
      return fl;
    };
    auto best = [&]() {
      size_t b = 0;
      for (size_t k = 1; k < pts.size(); ++k)
        if (pts[k].f < pts[b].f) b = k;
      return b;
    };
    const double f1 = probe(1.0);
    double lam2;
    if (finite(f1)) {
      const double denom = 2.0 * (f1 - f0 - gdel);
      lam2 = denom > 0 ? -gdel / denom : 2.0;
      lam2 = std::min(std::max(lam2, 0.02), 8.0);
    } else {
      lam2 = 0.1;
    }
    if (std::fabs(lam2 - 1.0) > 0.02) {
      const double f2 = probe(lam2);
      if (finite(f1) && finite(f2)) {
        const double l1 = 1.0, l2 = lam2;
        const double a = ((f1 - f0) / l1 - (f2 - f0) / l2) / (l1 - l2);
        const double b = (f1 - f0) / l1 - a * l1;
        if (a > 0) {
          const double lam3 = std::min(std::max(-b / (2 * a), 0.01), 10.0);
          if (std::min(std::fabs(lam3 - l1), std::fabs(lam3 - l2)) > 0.05 * lam3) probe(lam3);
        }
      }
    }
    double lam = 1.0;
    if (!pts.empty()) {
      lam = pts[0].lam;
      for (auto& p : pts) lam = std::min(lam, p.lam);
    }
    int tries = 0;
    while ((pts.empty() || pts[best()].f >= f0) && tries < 10) {
      ++tries;
      lam *= 0.2;
      probe(lam);
    }
    if (pts.empty() || pts[best()].f >= f0) {
      *lam_out = 0.0;
      *f_out = f0;
      return;
    }
    *lam_out = pts[best()].lam;
    *f_out = pts[best()].f;
  }

  // line search with value+gradient probes: the secant of the directional derivative locates the line minimum
  bool line_search_fg(const Vec& y, const Vec& step, double f0, double gdel, Probe* out) {
    std::vector<Probe> c;
    auto probe = [&](double lam, double* fl, double* dl) {
      Vec yy(y);
      for (int i = 0; i < n; ++i) yy[i] += lam * step[i];
      Probe p;
      p.lam = lam;
      fg_int(yy, &p.f, &p.g);
      if (finite(p.f) && all_finite(p.g)) {
        *fl = p.f;
        *dl = dot(p.g, step);
        c.push_back(std::move(p));
        return true;
      }
      *fl = std::numeric_limits<double>::infinity();
      return false;
    };
    auto fmin = [&]() {
      double m = std::numeric_limits<double>::infinity();
      for (auto& p : c) m = std::min(m, p.f);
      return m;
    };
    double f1, d1;
    if (probe(1.0, &f1, &d1)) {
      double lam2 = d1 > gdel ? gdel / (gdel - d1) : 2.5;  // still descending at lambda = 1 -> extend
      lam2 = std::min(std::max(lam2, 0.05), 6.0);
      const bool far = std::fabs(lam2 - 1.0) > kFar;
      if (f1 >= f0 || far) {
        double f2, d2;
        probe(lam2, &f2, &d2);
      }
    }
    double lam = 1.0;
    if (!c.empty()) {
      lam = c[0].lam;
      for (auto& p : c) lam = std::min(lam, p.lam);
    }
    int tries = 0;
    while ((c.empty() || fmin() >= f0) && tries < 10) {
      ++tries;
      lam *= 0.2;
      double fl, dl;
      probe(lam, &fl, &dl);
    }
    if (c.empty()) return false;
    size_t b = 0;
    for (size_t k = 1; k < c.size(); ++k)
      if (c[k].f < c[b].f) b = k;
    if (c[b].f >= f0) return false;
    *out = std::move(c[b]);
    return true;
  }

  // HESSE in internal coordinates
  Mat hesse_internal(const Vec& y, double f, const Vec& g, const Mat& V) {
    Mat H(n * n, 0.0);
    Vec sig(n);
    for (int i = 0; i < n; ++i) sig[i] = std::sqrt(std::max(V[i * n + i], 1e-300));
    if (has_grad) {
      // rows from differences of the analytic gradient: forward differences reuse g(y) (P passes); when the forward
      // matrix is asymmetric beyond 1e-3 the backward points are added and the differences become central (2P passes)
      Vec hs(n);
      for (int i = 0; i < n; ++i) hs[i] = std::max(0.02 * sig[i], 1e-8);
      std::vector<Vec> gp(n);
      for (int i = 0; i < n; ++i) {
        Vec yp(y);
        yp[i] += hs[i];
        gp[i] = g_int(yp);
        for (int j = 0; j < n; ++j) H[i * n + j] = (gp[i][j] - g[j]) / hs[i];
      }
      Vec d(n);
      for (int i = 0; i < n; ++i) d[i] = std::sqrt(std::max(std::fabs(H[i * n + i]), 1e-300));
      double asym = 0.0, norm = 0.0;
      bool ok = true;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          const double cij = H[i * n + j] / d[i] / d[j], cji = H[j * n + i] / d[i] / d[j];
          if (!finite(cij)) ok = false;
          asym += (cij - cji) * (cij - cji);
          norm += cij * cij;
        }
      if (!(ok && std::sqrt(asym) <= 1e-3 * std::sqrt(norm))) {
        for (int i = 0; i < n; ++i) {
          Vec ym(y);
          ym[i] -= hs[i];
          const Vec gm = g_int(ym);
          for (int j = 0; j < n; ++j) H[i * n + j] = (gp[i][j] - gm[j]) / (2 * hs[i]);
        }
      }
      Mat S(n * n);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S[i * n + j] = 0.5 * (H[i * n + j] + H[j * n + i]);
      return S;
    }
    Vec hs(n), fp(n), fm(n);
    for (int i = 0; i < n; ++i) hs[i] = std::max(0.1 * sig[i], 1e-7);
    for (int i = 0; i < n; ++i) {
      Vec yp(y), ym(y);
      yp[i] += hs[i];
      ym[i] -= hs[i];
      fp[i] = f_int(yp);
      fm[i] = f_int(ym);
      H[i * n + i] = (fp[i] + fm[i] - 2 * f) / (hs[i] * hs[i]);
    }
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        Vec ypp(y);
        ypp[i] += hs[i];
        ypp[j] += hs[j];
        H[i * n + j] = H[j * n + i] = (f_int(ypp) + f - fp[i] - fp[j]) / (hs[i] * hs[j]);
      }
    return H;
  }

  double quad(const Mat& V, const Vec& g) { return dot(g, matvec(V, g)); }

  void run(const double* x0, const double* steps, double* x_out, double* cov_out, double* grad_out, znll_migrad_result* r) {
    const char* message = "";
    bool converged = false;
    Vec y = tr.to_internal(Vec(x0, x0 + n));
    Vec d0 = tr.dext_dint(y);
    ysteps.assign(n, 0.0);
    num_steps.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
      const double d = std::max(std::fabs(d0[i]), 1e-8);
      ysteps[i] = std::min(std::fabs(steps[i]) / d, 0.5);
      num_steps[i] = std::max(ysteps[i] * 0.1, 1e-8);
    }
    double f;
    Vec g;
    if (use_fg) {
      fg_int(y, &f, &g);
    } else {
      f = f_int(y);
      g = g_int(y, &f);
    }
    Mat V = seed_metric(y, f, g, true);
    double edm = 0.5 * quad(V, g);
    double dcovar = 1.0;  // Minuit2's estimate of the relative error of V: 1 for a seed, 0 after HESSE, updated by DFP
    int restarts = 0, niter = 0;
    bool refreshed = false;
    for (int it = 1; it < 100000; ++it) {
      niter = it;
      if (nf + ng > maxfcn) {
        message = "call limit reached";
        break;
      }
      if (!(finite(f) && all_finite(g))) {
        message = "non-finite value or gradient";
        break;
      }
      if (edm * (1.0 + 3.0 * dcovar) < edm_goal) {  // VariableMetricBuilder: the EDM is inflated by the metric's uncertainty
        if (!hesse || dcovar <= 0.05) {              // strategy 1: HESSE only while the metric is still uncertain
          converged = true;
          break;
        }
        const Mat H = hesse_internal(y, f, g, V);
        Mat Vh;
        const Vec w = sym_eigenvalues(H, n);
        const bool pd = w[0] > 0 && spd_inverse(H, n, &Vh);
        if (pd) {
          V = Vh;
          edm = 0.5 * quad(V, g);
          dcovar = 0.0;
          if (edm < edm_goal) {
            converged = true;
            break;
          }
        } else if (restarts < 3) {
          ++restarts;
          V = seed_metric(y, f, g, false);
          edm = 0.5 * quad(V, g);
          dcovar = 1.0;
          if (edm < edm_goal) {  // the seed agrees that we are there; accept
            converged = true;
            message = "converged (hesse not positive definite, seed metric used)";
            break;
          }
        } else {
          message = "hesse failed: matrix not positive definite";
          break;
        }
      }
      if (info && !refreshed && it > 1 && edm * (1.0 + 3.0 * dcovar) < kRefreshEdm) {
        // about one sigma from the minimum the score outer product IS the Hessian up to O(1/sqrt(N))
        refreshed = true;
        Mat Vi;
        if (information_metric(y, &Vi) && quad(Vi, g) > 0) {
          V = Vi;
          dcovar = kRefreshDcovar;
          edm = 0.5 * quad(V, g);
        }
      }
      Vec step = matvec(V, g);
      for (double& s : step) s = -s;
      double gdel = dot(step, g);
      if (gdel >= 0) {  // metric not positive definite: reseed
        V = seed_metric(y, f, g, false);
        dcovar = 1.0;
        step = matvec(V, g);
        for (double& s : step) s = -s;
        gdel = dot(step, g);
        if (gdel >= 0) {
          message = "no descent direction";
          break;
        }
      }
      double lam = 0.0, f_new = f;
      Vec g_acc;
      bool have_g = false;
      auto search = [&]() {
        if (use_fg) {
          Probe p;
          if (line_search_fg(y, step, f, gdel, &p)) {
            lam = p.lam;
            f_new = p.f;
            g_acc = std::move(p.g);
            have_g = true;
          } else {
            lam = 0.0;
            f_new = f;
            have_g = false;
          }
        } else {
          line_search(y, step, f, gdel, &lam, &f_new);
        }
      };
      search();
      if (lam == 0.0) {
        if (restarts < 3) {
          ++restarts;
          V = seed_metric(y, f, g, false);
          edm = 0.5 * quad(V, g);
          dcovar = 1.0;
          if (edm < edm_goal) continue;
          step = matvec(V, g);
          for (double& s : step) s = -s;
          gdel = dot(step, g);
          if (gdel < 0) search();
        }
        if (lam == 0.0) {
          message = "line search failed to improve";
          converged = edm < edm_goal;
          break;
        }
      }
      Vec dy(n), y_new(n);
      for (int i = 0; i < n; ++i) {
        dy[i] = lam * step[i];
        y_new[i] = y[i] + dy[i];
      }
      const Vec g_new = have_g ? g_acc : g_int(y_new, &f_new);
      Vec dg(n);
      for (int i = 0; i < n; ++i) dg[i] = g_new[i] - g[i];
      // ---- DavidonErrorUpdator
      const double delgam = dot(dy, dg);
      const Vec Vg = matvec(V, dg);
      const double gvg = dot(dg, Vg);
      if (delgam > 0 && gvg > 0) {
        double sum_upd = 0.0, sum_v = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            double u = dy[i] * dy[j] / delgam - Vg[i] * Vg[j] / gvg;
            if (delgam > gvg) u += gvg * (dy[i] / delgam - Vg[i] / gvg) * (dy[j] / delgam - Vg[j] / gvg);
            V[i * n + j] += u;
            sum_upd += std::fabs(u);
            sum_v += std::fabs(V[i * n + j]);
          }
        dcovar = 0.5 * (dcovar + sum_upd / std::max(sum_v, 1e-300));
      }
      y = y_new;
      f = f_new;
      g = g_new;
      edm = 0.5 * quad(V, g);
    }
    const Vec x = tr.to_external(y), J = tr.dext_dint(y);
    for (int i = 0; i < n; ++i) {
      x_out[i] = x[i];
      if (grad_out) grad_out[i] = g[i] / (std::fabs(J[i]) > 1e-300 ? J[i] : 1.0);
      if (cov_out)
        for (int j = 0; j < n; ++j) cov_out[i * n + j] = V[i * n + j] * J[i] * J[j];
    }
    r->fval = f;
    r->edm = edm;
    r->niter = niter;
    r->nfcn = (int)nf;
    r->ngrad = (int)ng;
    r->converged = converged ? 1 : 0;
    r->valid = (converged && finite(f)) ? 1 : 0;
    if (converged && !message[0]) message = "converged";
    snprintf(r->message, sizeof r->message, "%s", message);
  }
};

}  // namespace

extern "C" int znll_migrad(znll_fcn_t fcn, znll_info_t info, void* user, int n, const double* x0, const double* steps,
                           const double* lower, const double* upper, const znll_migrad_options* options, double* x_out,
                           double* cov_out, double* grad_out, znll_migrad_result* result) {
  if (!fcn || n < 1 || !x0 || !steps || !options || !x_out || !result) return 1;
  memset(result, 0, sizeof *result);
  Driver d(fcn, info, user, n, lower, upper, *options);
  try {
    d.run(x0, steps, x_out, cov_out, grad_out, result);
  } catch (const Abort&) {
    result->nfcn = (int)d.nf;
    result->ngrad = (int)d.ng;
    snprintf(result->message, sizeof result->message, "%s", "aborted by the objective callback");
    return 2;
  }
  return 0;
}
