"""TEST INFRASTRUCTURE ONLY: the device header compiled for the host.

``zfit_b200/csrc/znll_device.cuh`` is pure arithmetic over CUDA built-ins, so with a small shim
(``tests/hostsim/cuda_shim.h``) g++ compiles it as host code.  This module builds such a shared object for a list of
model trees and runs the kernels' *per-event* code -- ``znll::event`` (every node's forward and reverse sweep, the
hand-rolled fp64 exp / log / reciprocal, the Kahan accumulation) in the kernels' own two-events-per-iteration loop --
single-threaded on the CPU.  Together with the host-only hooks ``znll_test_prologue`` / ``znll_test_epilogue`` of the
product library this gives the `-m "not gpu"` suite an end-to-end check of the device arithmetic against the oracle,
and lets kernel math be changed and checked without a GPU.  What it does NOT cover is the launch machinery (grid,
shuffles, tickets, cross-block and cross-GPU reductions, host publishing): that stays with the `-m gpu` tests.

Nothing in ``zfit_b200`` imports this; the product has no CPU path.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

from zfit_b200 import _znll as Z

HERE = Path(__file__).resolve().parent
SHIM_DIR = HERE / "hostsim"
BUILD_DIR = SHIM_DIR / "_build"
DEVICE_HEADER = HERE.parent / "zfit_b200" / "csrc" / "znll_device.cuh"

_RCP_ASM = 'asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));'

_DRIVER = r"""
#include "cuda_shim.h"
#include "znll_device_hostsim.cuh"
#include <cstring>
using namespace znll;
namespace {
#ifdef HOSTSIM_OLD_TABLES  // headers before the shared MathTabs block (kept for old-versus-new comparisons)
double2 g_logtab[128];
bool g_ready = false;
Ctx make_ctx() {
  if (!g_ready) {
    for (unsigned t = 0; t < 128; ++t) {
      threadIdx.x = t;
      init_logtab(g_logtab);
    }
    threadIdx.x = 0;
    g_ready = true;
  }
  Ctx cx;
  cx.logtab = g_logtab;
  return cx;
}
#define HOSTSIM_EXP(x, cx) fast_exp(x)
#define HOSTSIM_PAIR_OFFSET(off) (off)
#else
#define HOSTSIM_PAIR_OFFSET(off) (2.0 * (off))  // event() takes the offset of the whole pair
MathTabs g_tabs;
bool g_ready = false;
Ctx make_ctx() {
  if (!g_ready) {
    for (unsigned t = 0; t < ZNLL_BLOCK; ++t) {  // what the threads of a block do
      threadIdx.x = t;
      init_tabs(g_tabs);
    }
    threadIdx.x = 0;
    g_ready = true;
  }
  return ctx_of(g_tabs);
}
#ifdef ZNLL_HAS_OPTIMISTIC
#define HOSTSIM_PAIR_CTX(cx) with_mode(cx, false)  // nll_kernel's pair loop: optimistic sweep, one rare-path test per pair
#define HOSTSIM_TAIL_CTX(cx) with_mode(cx, true)
#define HOSTSIM_PAIR_ADD(acc, term) (acc)[0] += (term)  // plain per-thread accumulation, compensated merges across threads
#endif
#define HOSTSIM_EXP(x, cx) fast_exp(x, cx)
#endif
#ifndef HOSTSIM_PAIR_CTX
#define HOSTSIM_PAIR_CTX(cx) (cx)
#define HOSTSIM_TAIL_CTX(cx) (cx)
#define HOSTSIM_PAIR_ADD(acc, term) kahan_add(acc, term)
#endif
// event() returns the pair's summed term; headers before that change returned the two lanes
inline double pair_term(double t) { return t; }
inline double pair_term(D2 t) { return t.a + t.b; }
template <class M>
void load_pair(const double* const* cols, long long i, D2* x) {
  for (int k = 0; k < ZNLL_MAXCOL; ++k)
    if ((M::COLMASK >> k) & 1) {
      x[k].a = cols[k][2 * i];
      x[k].b = cols[k][2 * i + 1];
    }
}
// the event loop of nll_kernel, then the odd tail event.  Headers with the optimistic pair loop accumulate the value
// plainly per thread and merge the threads' partial sums error-free, so the simulation keeps the kernel's distribution of
// pairs over threads (pair i belongs to thread i mod T, T = grid x 256 as grid_for() of znll_host.cu chooses it) and merges
// with the kernel's own kahan_merge; older headers are run as one thread that owns all pairs (they compensate per pair).
template <class M, bool W, bool G>
void run_nll(const double* const* cols, const double* w, long long n, const double* c, double off, double* out) {
  constexpr int NACC = 2 + M::NS;
  const Ctx cx = make_ctx();
  const long long nvec = n >> 1;
#ifdef ZNLL_HAS_OPTIMISTIC
  long long grid = (((n + 1) / 2) + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > 296) grid = 296;
  const long long T = grid * 256;
#else
  const long long T = 1;
#endif
  const long long nthreads = nvec < T ? (nvec > 0 ? nvec : 1) : T;
  double* accs = new double[nthreads * NACC]();
  double* prods = new double[nthreads];
  for (long long t = 0; t < nthreads; ++t) prods[t] = 1.0;
  for (long long i = 0; i < nvec; ++i) {
    double* acc = accs + (i % nthreads) * NACC;
    D2 x[ZNLL_MAXCOL], wv;
    load_pair<M>(cols, i, x);
    wv.a = wv.b = 1.0;
    if (W) {
      wv.a = w[2 * i];
      wv.b = w[2 * i + 1];
    }
#ifdef ZNLL_HAS_SPLIT
    if constexpr (M::HAS_MULT && !W) {  // the running-product form of nll_kernel's pair loop (event_split)
      double& prod = prods[i % nthreads];
      acc[0] += event_split<M, D2, G>(HOSTSIM_PAIR_CTX(cx), c, x, HOSTSIM_PAIR_OFFSET(off), acc, prod);
      if (!prod_in_range(prod)) fold_product(acc, prod);
      continue;
    }
#endif
    HOSTSIM_PAIR_ADD(acc, pair_term(event<M, D2, W, G>(HOSTSIM_PAIR_CTX(cx), c, x, wv, HOSTSIM_PAIR_OFFSET(off), acc)));
  }
#ifdef ZNLL_HAS_SPLIT
  if constexpr (M::HAS_MULT && !W)
    for (long long t = 0; t < nthreads; ++t) fold_product(accs + t * NACC, prods[t]);
#endif
  if (n & 1) {  // thread 0 of block 0
    double xa[ZNLL_MAXCOL];
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) xa[k] = cols[k][n - 1];
    const double wa = W ? w[n - 1] : 1.0;
    kahan_add(accs, event<M, double, W, G>(HOSTSIM_TAIL_CTX(cx), c, xa, wa, off, accs));
  }
  double s = accs[0], cc = accs[1];
  for (long long t = 1; t < nthreads; ++t) kahan_merge(s, cc, accs[t * NACC], accs[t * NACC + 1]);
  out[0] = s;
  out[1] = cc;
  for (int j = 2; j < NACC; ++j) {
    double v = 0.0;
    for (long long t = 0; t < nthreads; ++t) v += accs[t * NACC + j];
    out[j] = v;
  }
  delete[] accs;
  delete[] prods;
}
// the body of hess_kernel (one thread owns all events): raw accumulators of the exact-Hessian pass
#ifdef ZNLL_HAS_OPTIMISTIC
template <class M, bool W>
int run_hess(const double* const* cols, const double* w, long long n, const double* c, double off, double* out, int* n_hess) {
  if constexpr (M::HESS) {
    constexpr int NS = M::NS, NE = NS + 1, NQ = NE * (NE + 1) / 2, NACC = 2 + NQ + M::NH;
    const Ctx cx = make_ctx();
    double acc[NACC];
    for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
    for (long long i = 0; i < (n >> 1); ++i) {
      D2 x[ZNLL_MAXCOL], wv;
      load_pair<M>(cols, i, x);
      wv.a = wv.b = 1.0;
      if (W) {
        wv.a = w[2 * i];
        wv.b = w[2 * i + 1];
      }
      kahan_add(acc, hess_event<M, D2, W>(cx, c, x, wv, 2.0 * off, acc));
    }
    if (n & 1) {
      double xa[ZNLL_MAXCOL];
      for (int k = 0; k < ZNLL_MAXCOL; ++k)
        if ((M::COLMASK >> k) & 1) xa[k] = cols[k][n - 1];
      const double wa = W ? w[n - 1] : 1.0;
      kahan_add(acc, hess_event<M, double, W>(cx, c, xa, wa, off, acc));
    }
    for (int j = 0; j < NACC; ++j) out[j] = acc[j];
    *n_hess = M::NH;
    return NACC;
  } else {
    return -1;
  }
}
#else
template <class M, bool W>
int run_hess(const double* const*, const double*, long long, const double*, double, double*, int*) { return -1; }
#endif
// the body of pdf_kernel
template <class M>
void run_pdf(const double* const* cols, long long n, const double* c, double* out) {
  const Ctx cx = make_ctx();
  for (long long i = 0; i < n; ++i) {
    double x[ZNLL_MAXCOL];
    for (int k = 0; k < ZNLL_MAXCOL; ++k)
      if ((M::COLMASK >> k) & 1) x[k] = cols[k][i];
    typename M::template Node<double> m;
    if constexpr (M::LOGNAT) {
      bool neg = false;
      const double lp = m.template fwd_log<0>(cx, c, x, neg);
      out[i] = neg ? -exp(lp) : exp(lp);
    } else {
      out[i] = m.template fwd_val<0>(cx, c, x);
    }
  }
}
typedef void (*nll_fn)(const double* const*, const double*, long long, const double*, double, double*);
typedef void (*pdf_fn)(const double* const*, long long, const double*, double*);
typedef int (*hess_fn)(const double* const*, const double*, long long, const double*, double, double*, int*);
struct Entry {
  const char* name;
  int n_slots, n_consts;
  nll_fn nll[2][2];
  pdf_fn pdf;
  hess_fn hess[2];
};
#define HOSTSIM_ENTRY(...)                                                                                        \
  {                                                                                                               \
    #__VA_ARGS__, __VA_ARGS__::NS, __VA_ARGS__::NC,                                                               \
        {{run_nll<__VA_ARGS__, false, false>, run_nll<__VA_ARGS__, false, true>},                                 \
         {run_nll<__VA_ARGS__, true, false>, run_nll<__VA_ARGS__, true, true>}},                                  \
        run_pdf<__VA_ARGS__>, {run_hess<__VA_ARGS__, false>, run_hess<__VA_ARGS__, true>}                         \
  }
const Entry kEntries[] = {
@ENTRIES@
};
const Entry* find(const char* name) {
  for (const Entry& e : kEntries)
    if (!std::strcmp(e.name, name)) return &e;
  return nullptr;
}
}  // namespace

extern "C" {
int hostsim_layout(const char* tree, int* n_slots, int* n_consts) {
  const Entry* e = find(tree);
  if (!e) return 1;
  *n_slots = e->n_slots;
  *n_consts = e->n_consts;
  return 0;
}
int hostsim_nll(const char* tree, int weighted, int grad, const double* const* cols, const double* w, long long n,
                const double* consts, double log_offset, double* acc_out) {
  const Entry* e = find(tree);
  if (!e) return 1;
  e->nll[weighted ? 1 : 0][grad ? 1 : 0](cols, w, n, consts, log_offset, acc_out);
  return 0;
}
// -> number of accumulators written (acc_out must hold 512), or -1 when the tree has no one-pass Hessian
int hostsim_hess(const char* tree, int weighted, const double* const* cols, const double* w, long long n,
                 const double* consts, double log_offset, double* acc_out, int* n_hess) {
  const Entry* e = find(tree);
  if (!e) return -2;
  return e->hess[weighted ? 1 : 0](cols, w, n, consts, log_offset, acc_out, n_hess);
}
int hostsim_pdf(const char* tree, const double* const* cols, long long n, const double* consts, double* out) {
  const Entry* e = find(tree);
  if (!e) return 1;
  e->pdf(cols, n, consts, out);
  return 0;
}
// the body of test_math_kernel; with `pairs` the two-lane instantiation the event loop uses
void hostsim_math(int kind, int pairs, long long n, const double* in, double* out) {
  const Ctx cx = make_ctx();
  long long i = 0;
  if (pairs)
    for (; i + 1 < n; i += 2) {
      D2 x;
      x.a = in[i];
      x.b = in[i + 1];
      const D2 r = kind == 0 ? HOSTSIM_EXP(x, cx) : (kind == 1 ? fast_log(x, cx) : fast_rcp(x));
      out[i] = r.a;
      out[i + 1] = r.b;
    }
  for (; i < n; ++i) out[i] = kind == 0 ? HOSTSIM_EXP(in[i], cx) : (kind == 1 ? fast_log(in[i], cx) : fast_rcp(in[i]));
}
}
"""


def _patched_header(text: str) -> str:
    if _RCP_ASM not in text:
        msg = "znll_device.cuh: the reciprocal-seed asm statement the host shim replaces was not found"
        raise RuntimeError(msg)
    return text.replace(_RCP_ASM, "y = hostsim_rcp_seed(x);")


class HostSim:
    """The device header compiled for the host for a given list of trees (kernel names as the plan reports them)."""

    def __init__(self, trees, defines=(), header_text=None):
        self.trees = list(dict.fromkeys(trees))
        text = DEVICE_HEADER.read_text() if header_text is None else header_text
        header = _patched_header(text)
        entries = ",\n".join(f"    HOSTSIM_ENTRY({t})" for t in self.trees)
        driver = _DRIVER.replace("@ENTRIES@", entries)
        shim = (SHIM_DIR / "cuda_shim.h").read_text()
        defines = list(defines)
        if "struct MathTabs" not in text:
            defines.append("HOSTSIM_OLD_TABLES")
        flags = ["-std=c++17", "-O1", "-fPIC", "-shared", "-march=native", *[f"-D{d}" for d in defines]]
        key = hashlib.sha256("\0".join([header, driver, shim, *flags]).encode()).hexdigest()[:16]
        work = BUILD_DIR / key
        so = work / "hostsim.so"
        if not so.exists():
            work.mkdir(parents=True, exist_ok=True)
            (work / "znll_device_hostsim.cuh").write_text(header)
            (work / "cuda_shim.h").write_text(shim)
            (work / "hostsim.cpp").write_text(driver)
            tmp = work / f"hostsim.{os.getpid()}.so"
            r = subprocess.run(["g++", *flags, "-I", str(work), str(work / "hostsim.cpp"), "-o", str(tmp)],
                               capture_output=True, text=True)
            if r.returncode:
                msg = "host compile of znll_device.cuh failed:\n" + r.stderr[-4000:]
                raise RuntimeError(msg)
            os.replace(tmp, so)
        self.lib = C.CDLL(str(so))
        dp = C.POINTER(C.c_double)
        self.lib.hostsim_nll.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_longlong, dp,
                                         C.c_double, dp]
        self.lib.hostsim_pdf.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_longlong, dp, dp]
        self.lib.hostsim_layout.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.lib.hostsim_math.argtypes = [C.c_int, C.c_int, C.c_longlong, dp, dp]
        self.lib.hostsim_math.restype = None
        self.lib.hostsim_hess.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_longlong, dp, C.c_double, dp,
                                          C.POINTER(C.c_int)]

    def layout(self, tree):
        ns, nc = C.c_int(0), C.c_int(0)
        if self.lib.hostsim_layout(tree.encode(), C.byref(ns), C.byref(nc)):
            msg = f"tree {tree} was not compiled into this HostSim"
            raise KeyError(msg)
        return ns.value, nc.value

    @staticmethod
    def _cols(cols):
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        arr = (C.c_void_p * Z.MAX_COLS)(*([c.ctypes.data for c in cols] + [None] * (Z.MAX_COLS - len(cols))))
        return cols, arr

    def accumulators(self, tree, consts, cols, weights=None, grad=True, log_offset=0.0):
        """Raw accumulators ``[sum, Kahan compensation, slot sums ...]`` of one pass over the events."""
        ns, nc = self.layout(tree)
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        if consts.size != nc:
            msg = f"{tree}: the prologue produced {consts.size} constants, the device code expects {nc}"
            raise ValueError(msg)
        cols, arr = self._cols(cols)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        out = np.zeros(2 + ns)
        dp = C.POINTER(C.c_double)
        rc = self.lib.hostsim_nll(tree.encode(), int(w is not None), int(grad), arr, None if w is None else w.ctypes.data,
                                  len(cols[0]), consts.ctypes.data_as(dp), float(log_offset), out.ctypes.data_as(dp))
        assert rc == 0
        return out

    def eval(self, nodes, slots, cols, weights=None, grad=True, log_offset=0.0):
        """What ``znll_eval`` returns for one channel -- prologue and epilogue from the product library (host code), the
        kernel's event loop from the host-compiled device header.  -> (value, d value / d slot | None, kernel name)"""
        name, consts = Z.test_prologue(nodes, slots)
        acc = self.accumulators(name, consts, cols, weights, grad, log_offset)
        sumw = float(len(cols[0])) if weights is None else float(np.sum(np.asarray(weights, dtype=np.float64)))
        value, g = Z.test_epilogue(nodes, slots, acc, sumw, grad=grad)
        return value, g, name

    def hess(self, nodes, slots, cols, weights=None, log_offset=0.0):
        """What ``znll_eval_hess`` returns for one channel: (value, d value / d slot, d2 value / d slot^2), or None when the
        tree has no one-pass Hessian.  The accumulators come from the host-compiled ``hess_event``, the assembly from the
        product library's own host code (``znll_test_hess_epilogue``)."""
        name, consts = Z.test_prologue(nodes, slots)
        cols, arr = self._cols(cols)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        acc = np.zeros(512)
        nh = C.c_int(0)
        dp = C.POINTER(C.c_double)
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        n_acc = self.lib.hostsim_hess(name.encode(), int(w is not None), arr, None if w is None else w.ctypes.data, len(cols[0]),
                                      consts.ctypes.data_as(dp), float(log_offset), acc.ctypes.data_as(dp), C.byref(nh))
        if n_acc < 0:
            return None
        return Z.test_hess_epilogue(nodes, slots, acc[:n_acc], nh.value)

    def pdf(self, nodes, slots, cols):
        name, consts = Z.test_prologue(nodes, slots)
        cols, arr = self._cols(cols)
        out = np.empty(len(cols[0]))
        dp = C.POINTER(C.c_double)
        rc = self.lib.hostsim_pdf(name.encode(), arr, len(cols[0]), consts.ctypes.data_as(dp), out.ctypes.data_as(dp))
        assert rc == 0
        return out

    def math(self, kind, x, pairs=False):
        """The device code's own exp (0) / log (1) / reciprocal (2); ``pairs``: through the two-lane instantiation."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        out = np.empty_like(x)
        dp = C.POINTER(C.c_double)
        self.lib.hostsim_math(kind, int(pairs), x.size, x.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out


class HostSimEngine:
    """Test double with the ``CudaEngine`` interface (``loss._set_engine_factory``): model lowering, prologue and epilogue
    are the product's own, the kernel's event loop is the host-compiled device header.  Lets API-level cases (random
    trees, composed parameters, shared parameters) exercise the device arithmetic on a CPU box."""

    def __init__(self, models, datas, fit_ranges):
        self.slot_params, self.channel_slots, self.channels = [], [], []
        self._args, self._oracle = (list(models), list(datas), list(fit_ranges)), None
        datas = [d if r is None else d.with_obs(r, guarantee_limits=False) for d, r in zip(datas, fit_ranges)]
        for model, data, rng in zip(models, datas, fit_ranges):
            nodes, slots = model._lower(data.obs, rng)
            start = len(self.slot_params)
            self.slot_params.extend(slots)
            self.channel_slots.append((start, len(self.slot_params)))
            cols = [np.ascontiguousarray(data.value(o), dtype=np.float64) for o in data.obs]
            name, _ = Z.test_prologue(nodes, [p.value() for p in slots])
            self.channels.append((nodes, name, cols, data.weights))
        self.sim = HostSim([name for _, name, _, _ in self.channels])
        self.n_local = [d.num_entries for d in datas]
        self.n_entries = list(self.n_local)
        self.samplesize = [d.samplesize for d in datas]
        self.kernel_names = [name for _, name, _, _ in self.channels]
        self.distributed = False

    def eval(self, slots, grad=True, log_offset=0.0, kahan=False):
        values, gs = [], np.zeros(len(self.slot_params))
        for (nodes, _, cols, w), (lo, hi) in zip(self.channels, self.channel_slots):
            v, g, _ = self.sim.eval(nodes, np.asarray(slots[lo:hi], dtype=np.float64), cols, w, grad=grad, log_offset=log_offset)
            values.append(v)
            if grad:
                gs[lo:hi] = g
        return np.asarray(values), (gs if grad else None)

    def eval_hess(self, slots, log_offset=0.0):
        values, gs, mats = [], np.zeros(len(self.slot_params)), []
        for (nodes, _, cols, w), (lo, hi) in zip(self.channels, self.channel_slots):
            got = self.sim.hess(nodes, np.asarray(slots[lo:hi], dtype=np.float64), cols, w, log_offset=log_offset)
            if got is None:
                return None
            v, g, H = got
            values.append(v)
            gs[lo:hi] = g
            mats.append(H)
        return np.asarray(values), gs, mats

    def eval_cov(self, slots):
        """The covariance pass only seeds the MIGRAD metric here: taken from the oracle engine (the kernel's own
        covariance pass is covered by the GPU tests)."""
        if self._oracle is None:
            from tests._oracle_engine import OracleEngine

            self._oracle = OracleEngine(*self._args)
        return self._oracle.eval_cov(slots)

    def launch_count(self):
        return 0

    def close(self):
        pass
