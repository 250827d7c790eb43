"""ctypes binding of ``libznll.so`` (C ABI declared in ``include/znll.h``).

This is the only place the Python host layer touches native code.  There is no CPU fallback:
if the library is missing or no CUDA device is visible, evaluating a loss raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
import os as _os

LIB_PATH = Path(_os.environ.get("ZNLL_LIB", _HERE / "libznll.so"))  # ZNLL_LIB: tuning experiments only

GAUSS, EXP, CB, CHEB, LEGENDRE, CHEB2, GCB, GET, GGET, BERNSTEIN, TGAUSS, HERMITE, LAGUERRE, SUM, PROD = (
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 17)
WANT_GRAD, KAHAN = 1, 2
MAX_COLS = 8


class Node(C.Structure):
    """Mirror of ``znll_node`` (include/znll.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("n_children", C.c_int32),
        ("column", C.c_int32),
        ("degree", C.c_int32),
        ("norm_lo", C.c_double),
        ("norm_hi", C.c_double),
        ("space_lo", C.c_double),
        ("space_hi", C.c_double),
        ("shift", C.c_double),
    ]


class Ref(C.Structure):
    """Mirror of ``znll_ref``: kind 0 = the constant ``value``, 1 = ``theta[index]``."""

    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("value", C.c_double)]


class Rule(C.Structure):
    """Mirror of ``znll_rule``: how one slot (or a channel's yield) follows from refs ``first .. first + n - 1``."""

    _fields_ = [("kind", C.c_int32), ("first", C.c_int32), ("n", C.c_int32), ("pad", C.c_int32)]


RULE_VALUE, RULE_FRACTION, RULE_REMAINDER, RULE_SUM = 0, 1, 2, 3
FULL = 4


class ZnllError(RuntimeError):
    pass


_lib = None

_DP = C.POINTER(C.c_double)
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int)

# every symbol include/znll.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "znll_abi_version": (C.c_int, []),
    "znll_last_error": (C.c_char_p, []),
    "znll_device_count": (C.c_int, []),
    "znll_plan_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "znll_plan_destroy": (None, [C.c_void_p]),
    "znll_plan_set_model": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Node), C.c_int]),
    "znll_plan_bind_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64]),
    "znll_plan_bind_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64]),
    "znll_partition_events": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "znll_cut_events": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, _DP, _DP,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "znll_prep_last_error": (C.c_char_p, []),
    "znll_plan_get_sumw": (C.c_int, [C.c_void_p, C.c_int, _DP]),
    "znll_plan_set_sumw": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "znll_plan_set_reduce_hook": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "znll_plan_n_slots": (C.c_int, [C.c_void_p, C.c_int]),
    "znll_plan_n_acc": (C.c_int, [C.c_void_p, C.c_int]),
    "znll_plan_total_slots": (C.c_int, [C.c_void_p]),
    "znll_plan_total_acc": (C.c_int, [C.c_void_p]),
    "znll_plan_kernel_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "znll_plan_kernel_origin": (C.c_int, [C.c_void_p, C.c_int]),
    "znll_plan_launch_count": (C.c_int64, [C.c_void_p]),
    "znll_eval": (C.c_int, [C.c_void_p, _DP, C.c_uint32, C.c_double, _DP, _DP, C.c_void_p]),
    "znll_launch": (C.c_int, [C.c_void_p, _DP, C.c_uint32, C.c_double, C.c_void_p]),
    "znll_plan_acc_device": (C.c_void_p, [C.c_void_p]),
    "znll_collect": (C.c_int, [C.c_void_p, C.c_uint32, _DP, _DP, C.c_void_p]),
    "znll_plan_set_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]),
    "znll_plan_exchange_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "znll_eval_cov": (C.c_int, [C.c_void_p, _DP, C.c_double, _DP, _DP, _DP, C.c_void_p]),
    "znll_eval_hess": (C.c_int, [C.c_void_p, _DP, C.c_double, _DP, _DP, _DP, C.c_void_p]),
    "znll_plan_has_hess": (C.c_int, [C.c_void_p, C.c_int]),
    "znll_plan_set_theta_map": (C.c_int, [C.c_void_p, C.c_int, _DP, _DP, C.POINTER(Ref), C.c_int, C.POINTER(Rule), C.POINTER(Rule)]),
    "znll_eval_theta": (C.c_int, [C.c_void_p, _DP, C.c_uint32, C.c_double, _DP, _DP, C.c_void_p]),
    "znll_migrad_theta": (C.c_int, None),  # structs of _migrad_native.py: argtypes are set there
    "znll_pdf": (C.c_int, [C.c_void_p, C.c_int, _DP, _DP, C.c_void_p]),
    "znll_pdf_device": (C.c_int, [C.c_void_p, C.c_int, _DP, C.c_void_p, C.c_void_p]),
    "znll_leaf_integral": (C.c_int, [C.POINTER(Node), _DP, _DP, _DP]),
    "znll_test_prologue": (C.c_int, [C.POINTER(Node), C.c_int, _DP, _DP, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.c_char_p, C.c_int]),
    "znll_test_rtc_image": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "znll_test_epilogue": (C.c_int, [C.POINTER(Node), C.c_int, _DP, _DP, C.c_double, _DP, _DP]),
    "znll_test_hess_epilogue": (C.c_int, [C.POINTER(Node), C.c_int, _DP, _DP, C.c_int, C.c_int, _DP, _DP, _DP]),
    "znll_migrad": (C.c_int, None),  # callback pointers and structs: see _migrad_native.py
    "znll_test_math": (C.c_int, [C.c_int, C.c_int64, _DP, _DP]),
    "znll_fp64_peak": (C.c_int, [C.c_int, C.c_double, _DP]),
}


def lib():
    """Load ``libznll.so`` (built in-tree by ``zfit_b200.build``).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            msg = (
                f"{LIB_PATH} is missing: build the CUDA extension with `python -m zfit_b200.build` "
                "(there is no CPU fallback)."
            )
            raise ZnllError(msg)
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.znll_abi_version() != 1:
            msg = "libznll.so ABI version mismatch; rebuild"
            raise ZnllError(msg)
        _lib = handle
    return _lib


def _check(rc):
    if rc:
        raise ZnllError(lib().znll_last_error().decode(errors="replace"))


def _dptr(a):
    return a.ctypes.data_as(_DP)


def leaf_integral(node: Node, slots):
    """Host-only: normalisation integral of a leaf and d I / d slot (no GPU needed)."""
    slots = np.ascontiguousarray(slots, dtype=np.float64)
    integral = np.zeros(1)
    dint = np.zeros(32)
    _check(lib().znll_leaf_integral(C.byref(node), _dptr(slots), _dptr(integral), _dptr(dint)))
    return float(integral[0]), dint[: len(slots)].copy()


class Plan:
    """Owns a ``znll_plan``: per-channel model trees, bound device columns, launch + collect."""

    def __init__(self, n_channels: int, device: int = 0):
        self._h = C.c_void_p()
        self._keep = []  # objects owning borrowed device memory
        L = lib()
        if L.znll_device_count() < 1:
            msg = "no CUDA device visible: zfit_b200 evaluates losses only on the GPU (no CPU fallback)"
            raise ZnllError(msg)
        _check(L.znll_plan_create(device, n_channels, C.byref(self._h)))
        self.n_channels = n_channels
        self.device = device
        self._buffers()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().znll_plan_destroy(self._h)
            self._h = C.c_void_p()
            self._reduce_cb = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, ch: int, nodes):
        arr = (Node * len(nodes))(*nodes)
        _check(lib().znll_plan_set_model(self._h, ch, arr, len(nodes)))
        self._buffers()

    def _buffers(self):
        L = lib()
        self.total_slots = L.znll_plan_total_slots(self._h)
        # persistent argument buffers of the step path with their ctypes pointers made once: building a pointer object
        # per call costs more than the C call itself (10.8 -> 2.3 us of Python per evaluation)
        self._values = np.zeros(self.n_channels)
        self._grad = np.zeros(max(self.total_slots, 1))
        self._slots = np.zeros(max(self.total_slots, 1))
        self._p_values, self._p_grad, self._p_slots = _dptr(self._values), _dptr(self._grad), _dptr(self._slots)
        self._f_eval, self._f_launch, self._f_collect = L.znll_eval, L.znll_launch, L.znll_collect

    def _stage_slots(self, slots):
        if np.size(slots) != self.total_slots:
            msg = f"expected {self.total_slots} slots, got {np.size(slots)}"
            raise ValueError(msg)
        if self.total_slots:
            self._slots[: self.total_slots] = slots

    def bind_device_ptrs(self, ch: int, col_ptrs, w_ptr, n: int, keepalive=None):
        cols = (C.c_void_p * len(col_ptrs))(*col_ptrs)
        _check(lib().znll_plan_bind_device(self._h, ch, len(col_ptrs), cols, C.c_void_p(w_ptr or 0), n))
        self._keep.append(keepalive)

    def bind_host(self, ch: int, cols, weights=None):
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        n = len(cols[0])
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        _check(lib().znll_plan_bind_host(self._h, ch, len(cols), ptrs, C.c_void_p(w.ctypes.data if w is not None else 0), n))

    def sumw(self, ch: int) -> float:
        out = np.zeros(1)
        _check(lib().znll_plan_get_sumw(self._h, ch, _dptr(out)))
        return float(out[0])

    def set_sumw(self, ch: int, v: float):
        _check(lib().znll_plan_set_sumw(self._h, ch, float(v)))

    def set_reduce_hook(self, fn):
        """``fn(array)``: sum a 1-D float64 numpy array over the ranks in place (identical on every rank); None removes it."""
        if fn is None:
            _check(lib().znll_plan_set_reduce_hook(self._h, None, None))
            self._reduce_cb = None
            return

        def trampoline(_user, buf, n):
            try:
                fn(np.ctypeslib.as_array(buf, shape=(n,)))
            except Exception:  # an exception must not cross the C frames
                import traceback

                traceback.print_exc()
                return 1
            return 0

        self._reduce_cb = REDUCE_FN(trampoline)  # kept alive with the plan
        _check(lib().znll_plan_set_reduce_hook(self._h, C.cast(self._reduce_cb, C.c_void_p), None))

    def exchange_bytes(self, world: int) -> int:
        return int(lib().znll_plan_exchange_bytes(self._h, world))

    def set_peers(self, rank: int, world: int, peer_ptrs, nbytes: int, keepalive=None):
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
        _check(lib().znll_plan_set_peers(self._h, rank, world, arr, nbytes))
        self._keep.append(keepalive)

    def n_slots(self, ch: int) -> int:
        return lib().znll_plan_n_slots(self._h, ch)

    def n_acc(self, ch: int) -> int:
        return lib().znll_plan_n_acc(self._h, ch)

    def total_acc(self) -> int:
        return lib().znll_plan_total_acc(self._h)

    def kernel_name(self, ch: int) -> str:
        return lib().znll_plan_kernel_name(self._h, ch).decode()

    def kernel_origin(self, ch: int) -> int:
        return lib().znll_plan_kernel_origin(self._h, ch)

    def launch_count(self) -> int:
        return int(lib().znll_plan_launch_count(self._h))

    def eval(self, slots, *, grad=True, log_offset=0.0, kahan=False, stream=0):
        """-> (values[n_channels], grad_slots[total_slots] | None); synchronous."""
        self._stage_slots(slots)
        flags = (WANT_GRAD if grad else 0) | (KAHAN if kahan else 0)
        if self._f_eval(self._h, self._p_slots, flags, log_offset, self._p_values, self._p_grad, stream or None):
            _check(1)
        return self._values.copy(), (self._grad[: self.total_slots].copy() if grad else None)

    def eval_cov(self, slots, *, log_offset=0.0, stream=0):
        """-> (values, grad_slots, [per channel (n_slots+1, n_slots+1) matrix]); see znll_eval_cov."""
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        sizes = [self.n_slots(k) + 1 for k in range(self.n_channels)]
        cov = np.zeros(sum(n * n for n in sizes))
        _check(lib().znll_eval_cov(self._h, _dptr(slots), float(log_offset), _dptr(self._values), _dptr(self._grad),
                                   _dptr(cov), C.c_void_p(stream)))
        mats, off = [], 0
        for n in sizes:
            mats.append(cov[off : off + n * n].reshape(n, n).copy())
            off += n * n
        return self._values.copy(), self._grad[: self.total_slots].copy(), mats

    def set_theta_map(self, n_theta, lower, upper, refs, slot_rules, yield_rules=None):
        """``refs``: [(kind, index, value)], ``slot_rules`` / ``yield_rules``: [(kind, first, n)] (include/znll.h)."""
        lo = np.ascontiguousarray(lower, dtype=np.float64)
        up = np.ascontiguousarray(upper, dtype=np.float64)
        r = (Ref * max(len(refs), 1))(*[Ref(int(k), int(i), float(v)) for k, i, v in refs])
        sr = (Rule * max(len(slot_rules), 1))(*[Rule(int(k), int(f), int(n), 0) for k, f, n in slot_rules])
        yr = None if yield_rules is None else (Rule * len(yield_rules))(*[Rule(int(k), int(f), int(n), 0) for k, f, n in yield_rules])
        _check(lib().znll_plan_set_theta_map(self._h, int(n_theta), _dptr(lo), _dptr(up), r, len(refs), sr, yr))
        self.n_theta = int(n_theta)
        self._theta = np.zeros(max(self.n_theta, 1))
        self._gtheta = np.zeros(max(self.n_theta, 1))
        self._tvalue = np.zeros(1)
        self._p_theta, self._p_gtheta, self._p_tvalue = _dptr(self._theta), _dptr(self._gtheta), _dptr(self._tvalue)
        self._f_eval_theta = lib().znll_eval_theta

    def eval_theta(self, theta, *, grad=True, full=False, log_offset=0.0, stream=0):
        """-> (value, d value / d theta | None): slots from theta, the fused pass, the chain rule and the extended term
        behind the C ABI (``znll_eval_theta``)."""
        self._theta[: self.n_theta] = theta
        flags = (WANT_GRAD if grad else 0) | (FULL if full else 0)
        if self._f_eval_theta(self._h, self._p_theta, flags, log_offset, self._p_tvalue, self._p_gtheta if grad else None,
                              stream or None):
            _check(1)
        return float(self._tvalue[0]), (self._gtheta[: self.n_theta].copy() if grad else None)

    def has_hess(self) -> bool:
        """True when every channel's tree has the one-pass Hessian kernel (flat tree of supported nodes, catalogue)."""
        return all(lib().znll_plan_has_hess(self._h, k) for k in range(self.n_channels))

    def eval_hess(self, slots, *, log_offset=0.0, stream=0):
        """-> (values, grad_slots, [per channel (n_slots, n_slots) d2 value / d slot^2]) in one pass, or None when a
        tree is unsupported (znll_eval_hess returned 3)."""
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        sizes = [self.n_slots(k) for k in range(self.n_channels)]
        hess = np.zeros(max(1, sum(n * n for n in sizes)))
        rc = lib().znll_eval_hess(self._h, _dptr(slots), float(log_offset), _dptr(self._values), _dptr(self._grad),
                                  _dptr(hess), C.c_void_p(stream))
        if rc == 3:
            return None
        _check(rc)
        mats, off = [], 0
        for n in sizes:
            mats.append(hess[off : off + n * n].reshape(n, n).copy())
            off += n * n
        return self._values.copy(), self._grad[: self.total_slots].copy(), mats

    def pdf_values(self, ch: int, slots, n: int, stream=0):
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        out = np.empty(n, dtype=np.float64)
        _check(lib().znll_pdf(self._h, ch, _dptr(slots), _dptr(out), C.c_void_p(stream)))
        return out

    def pdf_values_device(self, ch: int, slots, out_ptr: int, stream=0):
        """Per-event pdf values into a caller-owned device buffer (asynchronous on ``stream``)."""
        slots = np.ascontiguousarray(slots, dtype=np.float64)
        _check(lib().znll_pdf_device(self._h, ch, _dptr(slots), C.c_void_p(out_ptr), C.c_void_p(stream)))

    def launch(self, slots, *, grad=True, log_offset=0.0, stream=0):
        self._stage_slots(slots)
        if self._f_launch(self._h, self._p_slots, WANT_GRAD if grad else 0, log_offset, stream or None):
            _check(1)

    def acc_device_ptr(self) -> int:
        return int(lib().znll_plan_acc_device(self._h) or 0)

    def collect(self, *, grad=True, stream=0):
        if self._f_collect(self._h, WANT_GRAD if grad else 0, self._p_values, self._p_grad, stream or None):
            _check(1)
        return self._values.copy(), (self._grad[: self.total_slots].copy() if grad else None)


def exchange_bytes_max(world: int) -> int:
    """Size of the process-wide exchange buffer of the fused cross-GPU reduction: 128 B of flags + two parities x world
    x 256 accumulator slots (``ZNLL_XSTRIDE``); equals ``znll_plan_exchange_bytes`` of any plan."""
    return 128 + 8 * 2 * int(world) * 256


def partition_events(device: int, src_ptrs, dst_ptrs, src_w: int, dst_w: int, key_col: int, lo: float, hi: float, n: int,
                     stream=0):
    """Stable 32-bucket partition of resident event columns by ``key_col`` (device pointers; asynchronous on ``stream``)."""
    src = (C.c_void_p * len(src_ptrs))(*src_ptrs)
    dst = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
    L = lib()
    if L.znll_partition_events(device, len(src_ptrs), src, dst, C.c_void_p(src_w or 0), C.c_void_p(dst_w or 0), key_col,
                               float(lo), float(hi), int(n), C.c_void_p(stream or 0)):
        raise ZnllError(L.znll_prep_last_error().decode(errors="replace"))


def cut_events(device: int, src_ptrs, dst_ptrs, src_w: int, dst_w: int, lower, upper, n: int, *, less_a: int = 0, less_b: int = 0,
               stream=0) -> int:
    """Inclusive cut ``lower <= x <= upper`` on every column (and ``less_a < less_b`` when given) with a stable compaction of
    resident columns (device pointers); returns the number of kept events (synchronises ``stream``)."""
    src = (C.c_void_p * len(src_ptrs))(*src_ptrs)
    dst = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
    lo = np.ascontiguousarray(lower, dtype=np.float64)
    hi = np.ascontiguousarray(upper, dtype=np.float64)
    if lo.size != len(src_ptrs) or hi.size != len(src_ptrs):
        msg = "one lower and one upper limit per column"
        raise ValueError(msg)
    kept = C.c_int64(0)
    L = lib()
    if L.znll_cut_events(device, len(src_ptrs), src, dst, C.c_void_p(src_w or 0), C.c_void_p(dst_w or 0), _dptr(lo), _dptr(hi),
                         C.c_void_p(less_a or 0), C.c_void_p(less_b or 0), int(n), C.byref(kept), C.c_void_p(stream or 0)):
        raise ZnllError(L.znll_prep_last_error().decode(errors="replace"))
    return int(kept.value)


def fp64_peak(device: int = 0, ms: float = 200.0) -> float:
    out = np.zeros(1)
    _check(lib().znll_fp64_peak(device, float(ms), _dptr(out)))
    return float(out[0])


def test_math(kind: int, x):
    """Device evaluation of the kernel's fast exp (0) / log (1) / reciprocal (2) -- test hook."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _check(lib().znll_test_math(kind, x.size, _dptr(x), _dptr(out)))
    return out


def test_prologue(nodes, slots):
    """Host only: (kernel name, constant block) the prologue builds for a tree at these slots -- test hook."""
    arr = (Node * len(nodes))(*nodes)
    slots = np.ascontiguousarray(slots, dtype=np.float64)
    consts = np.zeros(1024)
    nc, ns = C.c_int(0), C.c_int(0)
    name = C.create_string_buffer(512)
    _check(lib().znll_test_prologue(arr, len(nodes), _dptr(slots), _dptr(consts), consts.size, C.byref(nc), C.byref(ns), name, 512))
    if ns.value != slots.size:
        msg = f"expected {ns.value} slots, got {slots.size}"
        raise ValueError(msg)
    return name.value.decode(), consts[: nc.value].copy()


def test_rtc_image(model: str, n_slots: int, use_cache: bool = True, hess_nh: int = -1):
    """Host only: (cubin bytes, number of kernels, came from the on-disk cache) for a tree outside the catalogue -- test hook.
    ``hess_nh`` >= 0: also build the one-pass Hessian kernels of a flat tree with that many second-order accumulators."""
    size, names, cached = C.c_size_t(0), C.c_int(0), C.c_int(0)
    _check(lib().znll_test_rtc_image(model.encode(), n_slots, int(hess_nh), int(use_cache), C.byref(size), C.byref(names),
                                     C.byref(cached)))
    return size.value, names.value, bool(cached.value)


def test_hess_epilogue(nodes, slots, acc, n_hess):
    """Host only: raw accumulators of the Hessian pass -> (value, d value / d slot, d2 value / d slot^2) -- test hook."""
    arr = (Node * len(nodes))(*nodes)
    slots = np.ascontiguousarray(slots, dtype=np.float64)
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    ns = slots.size
    value, g, h = np.zeros(1), np.zeros(max(ns, 1)), np.zeros(max(ns * ns, 1))
    _check(lib().znll_test_hess_epilogue(arr, len(nodes), _dptr(slots), _dptr(acc), acc.size, int(n_hess), _dptr(value), _dptr(g), _dptr(h)))
    return float(value[0]), g[:ns].copy(), h[: ns * ns].reshape(ns, ns).copy()


def test_epilogue(nodes, slots, acc, sumw, grad=True):
    """Host only: raw accumulators -> (channel value, d value / d slot | None) -- test hook."""
    arr = (Node * len(nodes))(*nodes)
    slots = np.ascontiguousarray(slots, dtype=np.float64)
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    if acc.size != 2 + slots.size:
        msg = f"expected {2 + slots.size} accumulators, got {acc.size}"
        raise ValueError(msg)
    value = np.zeros(1)
    g = np.zeros(max(slots.size, 1))
    _check(lib().znll_test_epilogue(arr, len(nodes), _dptr(slots), _dptr(acc), float(sumw), _dptr(value), _dptr(g) if grad else None))
    return float(value[0]), (g[: slots.size].copy() if grad else None)
