"""The oracle, the host prologue and the (host-compiled) device code against golden vectors computed by the REFERENCE'S
OWN function definitions (``tests/golden/reference_formulas.npz``, written by ``tests/golden/make_reference_formulas.py``
from the unmodified source of ``models/physics.py`` / ``models/polynomials.py`` executed over a numpy stand-in of
``zfit.z.numpy`` / ``tf``).  This pins every formula zfit itself owns on the path; TensorFlow's own kernels (erf,
Normal.prob/cdf, log_poisson_loss) are not pinned by it."""

from pathlib import Path

import numpy as np
import pytest

from oracle import pdfs as O
from oracle.backend import NumpyBackend
from tests._hostsim import HostSim
from zfit_b200 import _znll as Z

G = np.load(Path(__file__).resolve().parent / "golden" / "reference_formulas.npz")
B = NumpyBackend()
RTOL = 2e-14


def _close(got, ref, rtol=RTOL, scale=None):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    s = np.abs(ref) if scale is None else scale
    assert np.all(np.abs(got - ref) <= rtol * np.maximum(s, 1e-300)), float(np.max(np.abs(got - ref) / np.maximum(s, 1e-300)))


PHYS = {  # key -> (oracle class, parameter names, node kind, slots from a golden parameter row)
    "cb": (O.CrystalBall, ["mu", "sigma", "alpha", "n"], Z.CB, lambda p: list(p)),
    "dcb": (O.GeneralizedCB, ["mu", "sl", "al", "nl", "sr", "ar", "nr"], Z.GCB, lambda p: [p[0], p[1], p[2], p[3], p[1], p[4], p[5]]),
    "gcb": (O.GeneralizedCB, ["mu", "sl", "al", "nl", "sr", "ar", "nr"], Z.GCB, lambda p: list(p)),
    "get": (O.GaussExpTail, ["mu", "sigma", "alpha"], Z.GET, lambda p: list(p)),
    "gget": (O.GeneralizedGaussExpTail, ["mu", "sl", "al", "sr", "ar"], Z.GGET, lambda p: list(p)),
}


@pytest.mark.parametrize("key", list(PHYS))
def test_physics_shapes_and_integrals(key):
    cls, names, kind, to_slots = PHYS[key]
    x = G["x"]
    for row, shape, integrals in zip(G[f"{key}_params"], G[f"{key}_shape"], G[f"{key}_integral"]):
        slots = to_slots(row)
        P = dict(zip(names, (np.float64(v) for v in slots)))
        model = cls(*names, "x", (-8.0, 7.0))
        _close(model.unnorm(B, x, P, None), shape, scale=np.maximum(np.abs(shape), 1e-290))
        for (lo, hi), ref in zip(G["cb_limits"], integrals):
            _close(model.integral(B, P, float(lo), float(hi), None), ref)
            I, _ = Z.leaf_integral(Z.Node(kind, 0, 0, 0, float(lo), float(hi), 0, 0, 0), slots)  # the product's prologue
            # the |n - 1| < 1e-5 rows use the reference's log approximation: the same formula, but ill-conditioned in n
            _close(I, ref, rtol=1e-9 if key == "cb" and abs(row[3] - 1.0) < 1e-5 else 1e-13)


POLY = {"chebyshev": (O.Chebyshev, Z.CHEB), "legendre": (O.Legendre, Z.LEGENDRE), "chebyshev2": (O.Chebyshev2, Z.CHEB2),
        "hermite": (O.Hermite, Z.HERMITE), "laguerre": (O.Laguerre, Z.LAGUERRE)}


@pytest.mark.parametrize("fam", list(POLY))
def test_polynomial_shapes_and_integrals(fam):
    cls, kind = POLY[fam]
    lo_s, hi_s = (float(v) for v in G["poly_space"])
    x = G["poly_x"]
    for k in range(5):
        cs = G[f"poly_coeffs_{k}"]
        names = [f"c{i}" for i in range(len(cs))]
        P = dict(zip(names, (np.float64(c) for c in cs)))
        model = cls(names[1:], "x", (lo_s, hi_s), coeff0=names[0])
        shape = G[f"{fam}_shape_{k}"]
        _close(model.unnorm(B, x, P, None), shape, scale=np.maximum(np.abs(shape), 1.0))
        for (lo, hi), ref in zip(G["poly_limits"], G[f"{fam}_integral_{k}"]):
            _close(model.integral(B, P, float(lo), float(hi), None), ref, scale=max(abs(ref), 1.0))
            I, _ = Z.leaf_integral(Z.Node(kind, 0, 0, len(cs) - 1, float(lo), float(hi), lo_s, hi_s, 0), list(cs))
            _close(I, ref, rtol=1e-13, scale=max(abs(ref), 1.0))


def test_bernstein_shapes_and_integrals():
    lo_s, hi_s = (float(v) for v in G["poly_space"])
    x = G["poly_x"]
    for k in range(4):
        cs = G[f"bernstein_coeffs_{k}"]
        names = [f"b{i}" for i in range(len(cs))]
        P = dict(zip(names, (np.float64(c) for c in cs)))
        model = O.Bernstein(names, "x", (lo_s, hi_s))
        _close(model.unnorm(B, x, P, None), G[f"bernstein_shape_{k}"])
        for (lo, hi), ref in zip(G["poly_limits"], G[f"bernstein_integral_{k}"]):
            _close(model.integral(B, P, float(lo), float(hi), None), ref)
            I, _ = Z.leaf_integral(Z.Node(Z.BERNSTEIN, 0, 0, len(cs) - 1, float(lo), float(hi), lo_s, hi_s, 0), list(cs))
            _close(I, ref, rtol=1e-13)


def test_device_code_reproduces_the_reference_shapes():
    """pdf_kernel's arithmetic (host-compiled) x the prologue's integral = the reference's unnormalised shape."""
    lo_s, hi_s = (float(v) for v in G["poly_space"])
    trees = ["CBall<0>", "GCBall<0>", "GETail<0>", "GGETail<0>", "Cheb<0,3>", "Legd<0,3>", "Chb2<0,3>", "Herm<0,3>", "Lagr<0,3>",
             "Bern<0,3>"]
    sim = HostSim(trees)
    x = G["x"][(G["x"] >= -8.0) & (G["x"] <= 7.0)]
    keep = (G["x"] >= -8.0) & (G["x"] <= 7.0)
    for key, (cls, names, kind, to_slots) in PHYS.items():
        for row, shape in zip(G[f"{key}_params"], G[f"{key}_shape"]):
            slots = to_slots(row)
            node = Z.Node(kind, 0, 0, 0, -8.0, 7.0, 0, 0, 0)
            I, _ = Z.leaf_integral(node, slots)
            got = sim.pdf([node], slots, [x]) * I
            ref = shape[keep]
            big = ref > 1e-250  # below that the normalised log value leaves the fast path; compared in absolute terms
            _close(got[big], ref[big], rtol=2e-13)
            assert np.all(np.abs(got[~big] - ref[~big]) <= 1e-250)
    xs = G["poly_x"]
    for fam, (cls, kind) in POLY.items():
        cs = G["poly_coeffs_2"]  # degree 3
        node = Z.Node(kind, 0, 0, 3, lo_s, hi_s, lo_s, hi_s, 0)
        I, _ = Z.leaf_integral(node, list(cs))
        _close(sim.pdf([node], list(cs), [xs]) * I, G[f"{fam}_shape_2"], rtol=1e-13, scale=np.maximum(np.abs(G[f"{fam}_shape_2"]), 1.0))
    cs = G["bernstein_coeffs_2"]
    node = Z.Node(Z.BERNSTEIN, 0, 0, 3, lo_s, hi_s, lo_s, hi_s, 0)
    I, _ = Z.leaf_integral(node, list(cs))
    _close(sim.pdf([node], list(cs), [xs]) * I, G["bernstein_shape_2"], rtol=1e-13)


def test_exponential_value_shift_and_integral():
    """``Exponential._unnormalized_pdf`` / ``_shift_x`` and ``_exp_integral_from_any_to_any`` (models/basic.py)."""
    xe = G["exp_x"]
    for (lam, shift), shape, integrals in zip(G["exp_params"], G["exp_shape"], G["exp_integral"]):
        offs = 5300.0 if shift > 1000 else 0.0
        # the oracle derives the shift from the norm handed to the hook: a norm centred on `shift` (or none for shift 0)
        norm = False if shift == 0.0 else (shift - 7.0, shift + 7.0)
        model = O.Exponential("lam", "x", (offs + 0.0, offs + 5.0))
        P = {"lam": np.float64(lam)}
        _close(model.unnorm(B, xe + offs, P, norm), shape)
        for (lo, hi), ref in zip(G["exp_limits"], integrals):
            _close(model.integral(B, P, lo + offs, hi + offs, norm), ref)
            I, _ = Z.leaf_integral(Z.Node(Z.EXP, 0, 0, 0, lo + offs, hi + offs, 0, 0, float(shift)), [lam])
            _close(I, ref, rtol=1e-13)


class _GivenProbs:  # an oracle "model" whose pdf is a stored array, like the stand-in model of the generator
    def __init__(self, probs):
        self.probs = probs

    def pdf(self, B_, x, P, norm=None):
        return B_.asarray(self.probs)


def test_nll_assembly_from_given_probabilities():
    """``_unbinned_nll_tf`` + ``_nll_calc_unbinned_tf`` (core/loss.py:73-142): + 1e-307, weights before the offset, the
    per-event offset, the sum over channels, the sign, the (value, correction) pair."""
    from oracle import nll as ON

    p1, p2, w1 = G["nll_p1"], G["nll_p2"], G["nll_w1"]
    models = [_GivenProbs(p1), _GivenProbs(p2)]
    datas = [{"x": np.zeros(len(p1))}, {"x": np.zeros(len(p2))}]
    for weighted, off, kahan, nll, corr in G["nll_cases"]:
        weights = [w1, None] if weighted else None
        log_offset = False if np.isnan(off) else float(off)
        got = ON.unbinned_nll(models, datas, {}, weights=weights, log_offset=log_offset, kahan=bool(kahan))
        ref = nll - corr  # what UnbinnedNLL._loss_func returns (loss.py:1134-1158)
        scale = float(np.sum(np.abs(np.log(p1 + 1e-307)) * (np.abs(w1) if weighted else 1.0)) + np.sum(np.abs(np.log(p2 + 1e-307))))
        assert abs(float(got) - ref) <= 2e-14 * scale, (weighted, off, kahan, float(got), ref)


def test_functor_values_from_given_daughters():
    """``SumPDF._pdf`` (sum_i frac_i * pdf_i, functor.py:197-206) and ``ProductPDF._pdf`` (functor.py:413-421)."""
    d = G["functor_daughters"]
    fr = G["functor_fracs"]
    kids = []
    for row in d:
        k = _GivenProbs(row)
        k.obs, k.limits, k.yield_ = "x", (0.0, 1.0), None
        kids.append(k)
    s = O.SumPDF(kids, fracs=["f0", "f1", "f2"])
    P = {f"f{i}": np.float64(f) for i, f in enumerate(fr)}
    _close(s.pdf(B, {"x": np.zeros(d.shape[1])}, P), G["sum_pdf"], rtol=4e-16)
    _close(d[0] * d[1] * d[2], G["prod_pdf"], rtol=4e-16)  # ProductPDF of normalised daughters: the plain product, in order


class _GivenProbsExt(_GivenProbs):
    def __init__(self, probs, yield_):
        super().__init__(probs)
        self._y = yield_

    def get_yield(self, B_, P):
        return B_.const(self._y)


def test_loss_bodies_extended_term_and_constraints():
    """``UnbinnedNLL._loss_func_watched`` and ``ExtendedUnbinnedNLL._loss_func`` (core/loss.py:1134-1158, 1291-1318): the
    Poisson term per channel from samplesize (sum of weights) and yield, the offset added once per channel, Stirling only
    without an offset, the constraints' sum.  ``tf.nn.log_poisson_loss`` itself is the published formula on both sides."""
    from oracle import nll as ON

    p1, p2, w1 = G["nll_p1"], G["nll_p2"], G["nll_w1"]
    y = G["ext_yields"]
    models = [_GivenProbsExt(p1, y[0]), _GivenProbsExt(p2, y[1])]
    datas = [{"x": np.zeros(len(p1))}, {"x": np.zeros(len(p2))}]
    csum = float(np.sum(G["ext_constraints"]))
    for weighted, off, with_cons, ext, plain in G["loss_cases"]:
        weights = [w1, None] if weighted else None
        log_offset = False if np.isnan(off) else float(off)
        add = csum if with_cons else 0.0
        scale = float(np.sum(np.abs(np.log(p1 + 1e-307)) * (np.abs(w1) if weighted else 1.0)) + np.sum(np.abs(np.log(p2 + 1e-307))))
        got_plain = float(ON.unbinned_nll(models, datas, {}, weights=weights, log_offset=log_offset)) + add
        got_ext = float(ON.unbinned_nll(models, datas, {}, weights=weights, log_offset=log_offset, extended=True)) + add
        assert abs(got_plain - plain) <= 2e-14 * scale
        assert abs(got_ext - ext) <= 2e-14 * scale, (weighted, off, with_cons, got_ext, ext)


def test_committed_fixture_is_what_the_generator_produces_from_the_reference(tmp_path, capsys):
    """Provenance of ``tests/golden/reference_formulas.npz``: where the reference tree is present (the build container --
    never the GPU box), running the committed generator again on the reference's source gives the committed arrays bit
    for bit."""
    import importlib.util
    from pathlib import Path

    gen = Path(__file__).resolve().parent / "golden" / "make_reference_formulas.py"
    spec = importlib.util.spec_from_file_location("make_reference_formulas", gen)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not mod.REF.exists():
        pytest.skip("the reference source tree is not on this machine")
    mod.OUT = tmp_path / "regenerated.npz"
    mod.main()
    capsys.readouterr()
    new, old = np.load(mod.OUT), np.load(gen.parent / "reference_formulas.npz")
    assert sorted(new.files) == sorted(old.files) and len(new.files) > 100
    for key in new.files:
        assert new[key].shape == old[key].shape and np.array_equal(new[key], old[key], equal_nan=True), key
