"""Golden vectors from the REFERENCE'S OWN SOURCE for the formulas of the hot path that zfit itself owns.

zfit cannot be imported in this image (TensorFlow, tensorflow_probability, ... are not installed),
but the arithmetic that lives in zfit's own files -- the CrystalBall / DoubleCB / GeneralizedCB / GaussExpTail shapes and
their analytic integrals (``models/physics.py``), the recursive polynomials, Bernstein and their integrals
(``models/polynomials.py``), ``z.safe_where`` (``z/zextension.py``) -- is written against ``zfit.z.numpy`` (an alias of
``tf.experimental.numpy``, ``z/numpy.py``) and a dozen ``tf.*`` element-wise calls.  This script takes the *unmodified*
function definitions out of those files with ``ast``, executes them in a namespace where ``znp`` / ``tf`` / ``z`` are
numpy-backed stand-ins of the same names and semantics, and stores inputs and outputs as ``reference_formulas.npz``.

What this pins: every formula zfit owns on this path, as written in the reference (operation order included, in IEEE
fp64 through numpy).  What it does not pin: TensorFlow's and tfp's own kernels (``erf``, ``Normal.prob/cdf``,
``log_poisson_loss``, ``reduce_kahan_sum``) -- ``tf.math.erf`` is scipy's here.

Run in the build container only (needs /root/reference):   python tests/golden/make_reference_formulas.py
The tests read the committed .npz, never /root/reference.
"""

from __future__ import annotations

import ast
import sys
import types
from pathlib import Path

import numpy as np
import scipy.special

REF = Path("/root/reference/src/zfit")
OUT = Path(__file__).resolve().parent / "reference_formulas.npz"


# ----------------------------------------------------------------------------------------------------------------------
# numpy-backed stand-ins for the names the reference functions use
# ----------------------------------------------------------------------------------------------------------------------
def _f64(x):
    return np.asarray(x, dtype=np.float64)


class _Shape(tuple):  # tf.TensorShape: the reference reads `.shape.rank`
    @property
    def rank(self):
        return len(self)


class _Tensor(np.ndarray):
    """numpy array with the two tf.Tensor habits the reference relies on: ``.shape.rank`` and immutability (``a *= b``
    rebinds the name instead of writing into ``a``, so a scalar may grow to the broadcast shape)."""

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))

    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other


def _t(x):
    return np.asarray(x, dtype=np.float64).view(_Tensor)



znp = types.SimpleNamespace(
    exp=np.exp, log=np.log, abs=np.abs, square=np.square, power=np.power, sqrt=np.sqrt, where=np.where,
    zeros_like=np.zeros_like, ones_like=np.ones_like, multiply=np.multiply, maximum=np.maximum, minimum=np.minimum,
    sum=np.sum, atleast_1d=np.atleast_1d, concatenate=np.concatenate, stack=np.stack, asarray=lambda x, dtype=None: _t(x), reshape=lambda x, newshape=None, **kw: _t(np.reshape(x, newshape)),
    float64=np.float64, pi=np.pi,
)


def _tf_where(condition, x=None, y=None):
    return np.asarray(np.where(condition, x, y)).view(_Tensor)


tf = types.SimpleNamespace(
    where=_tf_where, less=np.less, greater=np.greater, less_equal=np.less_equal, greater_equal=np.greater_equal,
    ones_like=np.ones_like, sign=np.sign, square=np.square, Tensor=object, float64=np.float64,
    constant=lambda v, dtype=None: _f64(v), math=types.SimpleNamespace(erf=scipy.special.erf),
    nn=types.SimpleNamespace(log_poisson_loss=lambda *a, **k: _log_poisson_loss(*a, **k)),
)


def _function(*dargs, **dkw):  # z.function(wraps=..., ...) and bare @z.function
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]
    return lambda f: f


z = types.SimpleNamespace(
    function=_function, convert_to_tensor=lambda x, dtype=None, **kw: _t(x), constant=lambda v, dtype=None: _t(v),
    reduce_sum=lambda xs, axis=None: np.sum(np.asarray(xs, dtype=np.float64), axis=axis), square=np.square, numpy=znp,
    unstack_x=lambda x: x, exp=np.exp, assert_all_finite=lambda *a, **k: None,
)


def _top_level(path: Path, namespace: dict, only_functions=None):
    """Execute the top-level function definitions (and the plain assignments between them that resolve) of a reference
    file, unmodified, in ``namespace``.  Classes, imports and statements that need the rest of zfit are skipped."""
    tree = ast.parse(path.read_text())
    taken = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            if only_functions is not None and node.name not in only_functions:
                continue
        elif not isinstance(node, ast.Assign):
            continue
        mod = ast.Module(body=[node], type_ignores=[])
        try:
            exec(compile(mod, str(path), "exec"), namespace)  # noqa: S102 -- the point of this script
        except (NameError, AttributeError):
            continue  # e.g. `limits = Space(...)`, `Cls.register_analytic_integral(...)`
        taken.append(getattr(node, "name", None) or ast.unparse(node.targets[0]))
    return taken


def _methods(path: Path, class_name: str, names, namespace: dict):
    """The named methods of a reference class as plain functions (first argument ``self``), source unmodified."""
    tree = ast.parse(path.read_text())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in names:
                    local = dict(namespace)
                    exec(compile(ast.Module(body=[item], type_ignores=[]), str(path), "exec"), local)  # noqa: S102
                    out[item.name] = local[item.name]
    missing = set(names) - set(out)
    if missing:
        msg = f"{class_name}: methods {sorted(missing)} not found in {path}"
        raise RuntimeError(msg)
    return out


def _log_poisson_loss(targets, log_input, compute_full_loss=False):
    """tf.nn.log_poisson_loss by its published definition (third party, NOT pinned by this script): exp(c) - c*z, plus
    Stirling's z*log(z) - z + 0.5*log(2*pi*z) where z is outside [0, 1] when compute_full_loss."""
    targets, log_input = _f64(targets), _f64(log_input)
    result = np.exp(log_input) - log_input * targets
    if compute_full_loss:
        with np.errstate(all="ignore"):
            stirling = targets * np.log(targets) - targets + 0.5 * np.log(2.0 * np.pi * targets)
        result = result + np.where((targets >= 0.0) & (targets <= 1.0), 0.0, stirling)
    return _t(result)


def _kahan_sum(input_tensor, axis=0):  # stand-in for tfp.math.reduce_kahan_sum (third party, NOT pinned by this script)
    s = c = 0.0
    for v in np.asarray(input_tensor, dtype=np.float64):
        y = v - c
        t = s + y
        c = (t - s) - y
        s = t
    return _t(s), _t(c)


def load_reference():
    ns = {"znp": znp, "tf": tf, "z": z, "np": np, "typing": __import__("typing"), "annotations": None,
          "Mapping": __import__("collections.abc").abc.Mapping, "functools": __import__("functools"),
          "operator": __import__("operator"), "supports": lambda *a, **k: (lambda f: f),
          "is_container": lambda obj: isinstance(obj, (list, tuple)),
          "tfp": types.SimpleNamespace(math=types.SimpleNamespace(reduce_kahan_sum=_kahan_sum)),
          "SpecificFunctionNotImplemented": RuntimeError, "NotExtendedPDFError": RuntimeError}
    _top_level(REF / "z" / "zextension.py", ns, only_functions={"safe_where"})
    z.safe_where = ns["safe_where"]
    got = _top_level(REF / "models" / "physics.py", ns) + _top_level(REF / "models" / "polynomials.py", ns)
    # Exponential (models/basic.py): the two integral functions are top level, value and shift are methods
    got += _top_level(REF / "models" / "basic.py", ns, only_functions={"_exp_integral_from_any_to_any", "_exp_integral_func_shifting"})
    ns["exponential_methods"] = _methods(REF / "models" / "basic.py", "Exponential", {"_unnormalized_pdf", "_shift_x"}, ns)
    # NLL assembly (core/loss.py)
    got += _top_level(REF / "core" / "loss.py", ns, only_functions={"_unbinned_nll_tf", "_nll_calc_unbinned_tf"})
    # SumPDF / ProductPDF value (models/functor.py)
    ns["sum_methods"] = _methods(REF / "models" / "functor.py", "SumPDF", {"_pdf"}, ns)
    ns["prod_methods"] = _methods(REF / "models" / "functor.py", "ProductPDF", {"_pdf"}, ns)
    ns["nll_methods"] = _methods(REF / "core" / "loss.py", "UnbinnedNLL", {"_loss_func_watched"}, ns)
    ns["extnll_methods"] = _methods(REF / "core" / "loss.py", "ExtendedUnbinnedNLL", {"_loss_func"}, ns)
    got += ["Exponential._unnormalized_pdf", "Exponential._shift_x", "SumPDF._pdf", "ProductPDF._pdf",
            "UnbinnedNLL._loss_func_watched", "ExtendedUnbinnedNLL._loss_func"]
    return ns, got


class _Space:  # the few attributes of zfit.Space the integral functions read
    def __init__(self, lo, hi):
        self.lo, self.hi = float(lo), float(hi)

    @property
    def limit1d(self):
        return self.lo, self.hi

    @property
    def v1(self):
        return types.SimpleNamespace(limits=(np.array([[self.lo]]), np.array([[self.hi]])))

    @property
    def volume(self):
        return np.array([self.hi - self.lo])


class _PolyModel:  # RecursivePolynomial: degree, space, dtype and the rescaling method (polynomials.py:120-123)
    def __init__(self, ns, degree, space):
        self._ns, self.degree, self.space, self.dtype = ns, degree, space, np.float64

    def _polynomials_rescale(self, x):
        return self._ns["rescale_minus_plus_one"](x, limits=self.space)


def main():
    ns, got = load_reference()
    rng = np.random.default_rng(20240607)
    out = {}

    # ---- physics.py: shapes and analytic integrals
    x = np.concatenate([np.linspace(-9.0, 8.0, 341), rng.uniform(-9, 8, 160)])
    out["x"] = x
    cb_sets = np.array([  # mu, sigma, alpha, n
        [0.3, 1.1, 0.8, 2.0], [-0.3, 1.1, 0.8, 2.0], [0.5, 0.7, 1.5, 3.0], [0.5, 0.7, -1.5, 3.0], [0.1, 2.0, 0.4, 1.05],
        [0.1, 2.0, 2.5, 1.000001], [0.0, 1.0, 1.2, 10.0], [0.2, 1.3, -0.6, 1.000003]])
    out["cb_params"] = cb_sets
    out["cb_shape"] = np.stack([ns["crystalball_func"](x, *p) for p in cb_sets])
    cb_limits = np.array([[-8.0, 7.0], [0.5, 7.0], [-8.0, -3.0], [-1.0, 1.0], [-8.0, 0.0], [1.0, 3.0]])
    out["cb_limits"] = cb_limits
    out["cb_integral"] = np.array([[float(np.reshape(ns["crystalball_integral_func"](*p, lo, hi), ())) for lo, hi in cb_limits]
                                   for p in cb_sets])
    dcb_sets = np.array([  # mu, sigma, alphal, nl, alphar, nr
        [0.2, 1.1, 1.2, 2.5, 1.5, 3.0], [-0.4, 0.8, 0.7, 1.3, 2.0, 4.0], [0.0, 1.5, 2.0, 5.0, 0.5, 1.1]])
    out["dcb_params"] = dcb_sets
    out["dcb_shape"] = np.stack([ns["double_crystalball_func"](x, *p) for p in dcb_sets])
    out["dcb_integral"] = np.array([[float(np.reshape(ns["double_crystalball_mu_integral_func"](*p, lo, hi), ()))
                                     for lo, hi in cb_limits] for p in dcb_sets])
    gcb_sets = np.array([  # mu, sigmal, alphal, nl, sigmar, alphar, nr
        [0.2, 1.1, 1.2, 2.5, 0.9, 1.5, 3.0], [-0.4, 0.8, 0.7, 1.3, 1.6, 2.0, 4.0], [0.0, 1.5, 2.0, 5.0, 0.6, 0.5, 1.1]])
    out["gcb_params"] = gcb_sets
    out["gcb_shape"] = np.stack([ns["generalized_crystalball_func"](x, *p) for p in gcb_sets])
    out["gcb_integral"] = np.array([[float(np.reshape(ns["generalized_crystalball_mu_integral_func"](*p, lo, hi), ()))
                                     for lo, hi in cb_limits] for p in gcb_sets])
    get_sets = np.array([[0.2, 1.1, 1.2], [0.2, 1.1, -1.2], [-0.5, 0.6, 0.4], [0.0, 2.0, 2.5]])  # mu, sigma, alpha
    out["get_params"] = get_sets
    out["get_shape"] = np.stack([ns["gaussexptail_func"](x, *p) for p in get_sets])
    out["get_integral"] = np.array([[float(np.reshape(ns["gaussexptail_integral_func"](*p, lo, hi), ())) for lo, hi in cb_limits]
                                    for p in get_sets])
    gget_sets = np.array([[0.2, 1.1, 1.2, 0.9, 1.5], [-0.4, 0.8, 0.7, 1.6, 2.0], [0.0, 1.5, 2.0, 0.6, 0.5]])
    out["gget_params"] = gget_sets
    out["gget_shape"] = np.stack([ns["generalized_gaussexptail_func"](x, *p) for p in gget_sets])
    out["gget_integral"] = np.array([[float(np.reshape(ns["generalized_gaussexptail_integral_func"](*p, lo, hi), ()))
                                      for lo, hi in cb_limits] for p in gget_sets])

    # ---- polynomials.py: shapes on the rescaled observable and integrals over sub-ranges of the space
    space = _Space(-6.0, 8.0)
    xs = np.concatenate([np.linspace(-6.0, 8.0, 141), rng.uniform(-6, 8, 60)])
    zs = ns["rescale_minus_plus_one"](xs, limits=space)
    out["poly_space"] = np.array([space.lo, space.hi])
    out["poly_x"] = xs
    out["poly_z"] = zs
    coeff_sets = [np.array([1.0, 0.12]), np.array([1.0, 0.12, -0.07]), np.array([1.0, 0.12, -0.07, 0.05]),
                  np.array([0.8, -0.2, 0.1, 0.04, -0.03]), np.array([1.0, 0.3, -0.2, 0.1, 0.05, -0.02])]
    poly_limits = np.array([[-6.0, 8.0], [-2.0, 5.0], [-6.0, 0.0], [1.0, 7.5]])
    out["poly_limits"] = poly_limits
    families = {
        "chebyshev": ("chebyshev_shape", "func_integral_chebyshev1", True),
        "legendre": ("legendre_shape", "legendre_integral", True),
        "chebyshev2": ("chebyshev2_shape", "func_integral_chebyshev2", True),
        "hermite": ("hermite_shape", "func_integral_hermite", True),
        "laguerre": ("laguerre_shape", "func_integral_laguerre", True),
    }
    for k, cs in enumerate(coeff_sets):
        out[f"poly_coeffs_{k}"] = cs
    for fam, (shape, integral, _) in families.items():
        for k, cs in enumerate(coeff_sets):
            coeffs = [np.float64(c) for c in cs]
            out[f"{fam}_shape_{k}"] = _f64(ns[shape](x=zs, coeffs=coeffs))
            params = {f"c_{i}": np.float64(c) for i, c in enumerate(cs)}
            model = _PolyModel(ns, len(cs) - 1, space)
            vals = []
            for lo, hi in poly_limits:
                vals.append(float(np.ravel(ns[integral](limits=_Space(lo, hi), norm=None, params=dict(params), model=model))[0]))
            out[f"{fam}_integral_{k}"] = np.array(vals)
    # Bernstein: coefficients scale the basis polynomials directly; x rescaled to (0, 1)
    ts = ns["rescale_zero_one"](xs, space)
    for k, cs in enumerate([np.array([0.9, 1.2]), np.array([0.9, 1.2, 0.7]), np.array([0.9, 1.2, 0.7, 1.05]),
                            np.array([0.5, 1.5, 0.2, 1.1, 0.8])]):
        out[f"bernstein_coeffs_{k}"] = cs
        out[f"bernstein_shape_{k}"] = _f64(ns["bernstein_shape"](ts, [np.float64(c) for c in cs]))
        params = {f"c_{i}": np.float64(c) for i, c in enumerate(cs)}
        model = types.SimpleNamespace(space=space)
        out[f"bernstein_integral_{k}"] = np.array(
            [float(np.ravel(ns["func_integral_bernstein"](limits=_Space(lo, hi), params=params, model=model))[0]) for lo, hi in poly_limits])

    # ---- basic.py: Exponential value with its numerical shift and the shifted analytic integral
    exp_sets = np.array([[-0.003, 5300.0], [-0.6, 0.0], [0.8, 2.5], [-1e-5, 5300.0], [2.0, -1.0]])  # lambda, shift
    xe = np.concatenate([np.linspace(0.0, 5.0, 51), rng.uniform(0, 5, 30)])
    out["exp_params"], out["exp_x"] = exp_sets, xe
    exp_limits = np.array([[0.0, 5.0], [1.0, 2.0], [-3.0, 4.0]])
    out["exp_limits"] = exp_limits
    vals, ints = [], []
    for lam, shift in exp_sets:
        me = ns["exponential_methods"]
        this = types.SimpleNamespace(_calc_numerics_data_shift=lambda shift=shift: np.float64(shift))
        this._shift_x = lambda x, this=this: me["_shift_x"](this, x)
        offs = 5300.0 if shift > 1000 else 0.0  # the mass-like rows are evaluated around their shift
        xdata = types.SimpleNamespace(unstack_x=lambda offs=offs: xe + offs)
        vals.append(_f64(me["_unnormalized_pdf"](this, xdata, {"lambda": np.float64(lam)})))
        row = []
        for lo, hi in exp_limits:
            lim = types.SimpleNamespace(rect_limits=(np.array([[lo + offs]]), np.array([[hi + offs]])))
            row.append(float(np.ravel(ns["_exp_integral_from_any_to_any"](limits=lim, params={"lambda": np.float64(lam)}, model=this))[0]))
        ints.append(row)
    out["exp_shape"], out["exp_integral"] = np.stack(vals), np.array(ints)

    # ---- core/loss.py: NLL assembly from given per-event probabilities (two channels; weights; offsets; epsilon)
    n1, n2 = 4001, 2500
    p1 = np.exp(rng.normal(-2.0, 1.5, n1))
    p1[:6] = [0.0, 1e-320, 1e-310, 3e-308, 1e-300, 1e-200]  # where the + 1e-307 matters
    p2 = np.exp(rng.normal(-1.0, 0.7, n2))
    w1 = rng.normal(1.0, 0.6, n1)
    out["nll_p1"], out["nll_p2"], out["nll_w1"] = p1, p2, w1

    class _Model:
        def __init__(self, probs):
            self.probs = probs

        def pdf(self, data, norm=None):
            return _t(self.probs)

    class _Data:
        def __init__(self, weights=None):
            self.weights = None if weights is None else _t(weights)

        def with_obs(self, obs):
            return self

    cases = []
    for weights in (None, w1):
        for log_offset in (False, -1.75, 3.0):
            for kahan in (False, True):
                nll, corr = ns["_unbinned_nll_tf"](model=[_Model(p1), _Model(p2)], data=[_Data(weights), _Data()],
                                                   fit_range=[None, None], log_offset=log_offset, kahan=kahan)
                cases.append([0.0 if weights is None else 1.0, np.nan if log_offset is False else log_offset, float(kahan),
                              float(nll), float(corr)])
    out["nll_cases"] = np.array(cases)  # weighted?, log_offset (NaN = False), kahan?, nll, corr

    # the loss bodies around it: UnbinnedNLL._loss_func_watched (constraints, nll - corr) and
    # ExtendedUnbinnedNLL._loss_func (per-channel Poisson term from samplesize and yield, offset per channel)
    class _Ext(_Model):
        def __init__(self, probs, yield_):
            super().__init__(probs)
            self.is_extended, self._y = True, yield_

        def get_yield(self):
            return _t(self._y)

    class _ExtData(_Data):
        def __init__(self, n, weights=None):
            super().__init__(weights)
            self.samplesize = float(n) if weights is None else float(np.sum(weights))

    yields = np.array([3900.0, 2600.5])
    out["ext_yields"] = yields
    constraint_values = np.array([0.37, 1.25])
    out["ext_constraints"] = constraint_values
    cons = [types.SimpleNamespace(value=lambda v=v: _t(v)) for v in constraint_values]
    cases = []
    for weights in (None, w1):
        for log_offset in (False, -1.75):
            for with_cons in (False, True):
                mods = [_Ext(p1, yields[0]), _Ext(p2, yields[1])]
                dats = [_ExtData(n1, weights), _ExtData(n2)]
                ext = ns["extnll_methods"]["_loss_func"](None, mods, dats, [None, None], cons if with_cons else [], log_offset)
                plain = ns["nll_methods"]["_loss_func_watched"](None, dats, mods, [None, None], cons if with_cons else [], log_offset,
                                                               sumtype=None)
                cases.append([0.0 if weights is None else 1.0, np.nan if log_offset is False else log_offset, float(with_cons),
                              float(ext), float(plain)])
    out["loss_cases"] = np.array(cases)  # weighted?, log_offset (NaN = False), constraints?, extended NLL, plain NLL

    # ---- models/functor.py: SumPDF._pdf and ProductPDF._pdf from given daughter values
    d1, d2, d3 = p1[:2000] + 1e-3, np.exp(rng.normal(-1.0, 0.5, 2000)), rng.uniform(0.1, 2.0, 2000)
    out["functor_daughters"] = np.stack([d1, d2, d3])
    fr = np.array([0.25, 0.6, 0.15])
    out["functor_fracs"] = fr
    kids = [types.SimpleNamespace(pdf=lambda x=None, var=None, norm=None, d=d: _t(d), norm="same") for d in (d1, d2, d3)]
    this = types.SimpleNamespace(pdfs=kids, params={f"frac_{i}": np.float64(f) for i, f in enumerate(fr)})
    out["sum_pdf"] = _f64(ns["sum_methods"]["_pdf"](this, None, "same"))
    this = types.SimpleNamespace(pdfs=kids, _prod_is_same_obs_pdf=False, _prod_disjoint_obs_pdfs=kids)
    out["prod_pdf"] = _f64(ns["prod_methods"]["_pdf"](this, None, "same"))

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(out)} arrays); reference definitions executed: {len(got)}")
    print(", ".join(got))


if __name__ == "__main__":
    sys.exit(main())
