"""Shared parity cases at the C-ABI level: (znll nodes, slots, columns) next to the oracle model.

Each case mirrors one of BASELINE.json's configs at a size the oracle finishes
in seconds.  The oracle is the checker only.
"""

from __future__ import annotations

import numpy as np

from oracle import pdfs as O
from zfit_b200 import _znll as Z


def _trunc(rng, draw, lo, hi, n):
    out = np.empty(0)
    while out.size < n:
        x = draw(max(1024, 2 * (n - out.size)))
        out = np.concatenate([out, x[(x >= lo) & (x <= hi)]])
    return out[:n]


def _cb_sample(rng, n, mu, sigma, alpha, nn):
    from scipy import stats

    return stats.crystalball.rvs(beta=alpha, m=nn, loc=mu, scale=sigma, size=n, random_state=rng)


class Case:
    def __init__(self, name, nodes, slot_names, slots, cols, oracle_model, obs, weights=None, kernel=None):
        self.name = name
        self.nodes = nodes
        self.slot_names = slot_names  # one oracle parameter name per slot (None = constant slot)
        self.slots = np.asarray(slots, dtype=np.float64)
        self.cols = cols
        self.oracle_model = oracle_model
        self.obs = obs
        self.weights = weights
        self.kernel = kernel

    @property
    def data(self):
        return dict(zip(self.obs, self.cols))

    @property
    def P(self):
        return {n: np.float64(v) for n, v in zip(self.slot_names, self.slots) if n is not None}

    def free(self):
        idx = [i for i, n in enumerate(self.slot_names) if n is not None]
        return idx, [self.slot_names[i] for i in idx], self.slots[idx]


def gauss_case(n=10_000, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.normal(2, 3, n)
    x = x[(x >= -2) & (x <= 3)]
    nodes = [Z.Node(Z.GAUSS, 0, 0, 0, -2, 3, 0, 0, 0)]
    m = O.Gauss("mu", "sigma", "x", (-2, 3))
    return Case("cfg1_gauss", nodes, ["mu", "sigma"], [1.2, 1.3], [x], m, ["x"], kernel="Gauss<0>")


def cb_exp_case(n=200_000, seed=2, alpha=1.2, nn=2.5):
    rng = np.random.default_rng(seed)
    lo, hi = 5000.0, 5600.0
    nsig = int(0.3 * n)
    sig = _trunc(rng, lambda k: _cb_sample(rng, k, 5280, 20, 1.5, 3), lo, hi, nsig)
    bkg = _trunc(rng, lambda k: lo + rng.exponential(1 / 0.002, k), lo, hi, n - nsig)
    x = rng.permutation(np.concatenate([sig, bkg]))
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.CB, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.EXP, 0, 0, 0, lo, hi, 0, 0, 0.5 * (lo + hi)),
    ]
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "m", (lo, hi))
    ex = O.Exponential("lam", "m", (lo, hi))
    m = O.SumPDF([cb, ex], fracs=["f0", "f1"])
    names = ["f0", "f1", "mu", "sigma", "alpha", "n", "lam"]
    slots = [0.25, 0.75, 5275.0, 25.0, alpha, nn, -0.003]
    return Case("cfg2_cb_exp", nodes, names, slots, [x], m, ["m"], kernel="Sum<CBall<0>,Expo<0>>")


def gauss_cheb_case(n=200_000, seed=3, weighted=False, frac=0.4, coeffs=(0.2, -0.1, 0.05, 0.02)):
    rng = np.random.default_rng(seed)
    lo, hi = -10.0, 10.0
    ns = int(frac * n)
    sig = _trunc(rng, lambda k: rng.normal(1, 1.5, k), lo, hi, ns)
    # background ~ chebyshev shape by accept-reject
    cc = np.array([1.0, *coeffs])
    def draw(k):
        u = rng.uniform(lo, hi, 3 * k)
        z = (2 * u - lo - hi) / (hi - lo)
        v = np.polynomial.chebyshev.chebval(z, cc)
        keep = rng.uniform(0, 1.6, 3 * k) < v
        return u[keep]
    bkg = _trunc(rng, draw, lo, hi, n - ns)
    x = rng.permutation(np.concatenate([sig, bkg]))
    w = None
    if weighted:
        w = rng.normal(1.0, 0.6, n)
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.CHEB, 0, 0, 4, lo, hi, lo, hi, 0),
    ]
    g = O.Gauss("mu", "sigma", "x", (lo, hi))
    c = O.Chebyshev(["c1", "c2", "c3", "c4"], "x", (lo, hi))
    m = O.SumPDF([g, c], fracs=["f0", "f1"])
    names = ["f0", "f1", "mu", "sigma", None, "c1", "c2", "c3", "c4"]
    slots = [frac + 0.05, 1 - frac - 0.05, 0.8, 1.7, 1.0, 0.15, -0.05, 0.03, 0.01]
    return Case(
        "cfg5_weighted_gauss_cheb" if weighted else "cfg3_gauss_cheb",
        nodes, names, slots, [x], m, ["x"], weights=w, kernel="Sum<Gauss<0>,Cheb<0,4>>",
    )


def product_case(n=200_000, seed=5):
    rng = np.random.default_rng(seed)
    x = _trunc(rng, lambda k: rng.normal(0.3, 1.2, k), -4, 4, n)
    y = _trunc(rng, lambda k: rng.exponential(1 / 0.7, k), 0, 5, n)
    cc = np.array([1.0, 0.2, -0.1, 0.05, 0.02])
    def draw(k):
        u = rng.uniform(-2, 4, 3 * k)
        zz = (2 * u + 2 - 4) / 6
        v = np.polynomial.chebyshev.chebval(zz, cc)
        return u[rng.uniform(0, 1.6, 3 * k) < v]
    z = _trunc(rng, draw, -2, 4, n)
    nodes = [
        Z.Node(Z.PROD, 3, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, -4, 4, 0, 0, 0),
        Z.Node(Z.EXP, 0, 1, 0, 0, 5, 0, 0, 0.0),
        Z.Node(Z.CHEB, 0, 2, 4, -2, 4, -2, 4, 0),
    ]
    g = O.Gauss("mu", "sigma", "x", (-4, 4))
    e = O.Exponential("lam", "y", (0, 5))
    c = O.Chebyshev(["c1", "c2", "c3", "c4"], "z", (-2, 4))
    m = O.ProductPDF([g, e, c])
    names = ["mu", "sigma", "lam", None, "c1", "c2", "c3", "c4"]
    slots = [0.2, 1.3, -0.6, 1.0, 0.15, -0.05, 0.03, 0.01]
    return Case("cfg4_product3d", nodes, names, slots, [x, y, z], m, ["x", "y", "z"],
                kernel="Prod<Gauss<0>,Expo<1>,Cheb<2,4>>")


def rtc_case(n=50_000, seed=7):
    """A tree shape that is NOT in the ahead-of-time catalogue: exercised through NVRTC."""
    rng = np.random.default_rng(seed)
    lo, hi = -8.0, 9.0
    x = _trunc(rng, lambda k: np.concatenate([rng.normal(-1, 1.0, k // 3), rng.normal(2, 2.5, k // 3), rng.uniform(lo, hi, k // 3)]), lo, hi, n)
    nodes = [
        Z.Node(Z.SUM, 3, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.CHEB, 0, 0, 2, lo, hi, lo, hi, 0),
    ]
    g1 = O.Gauss("mu1", "s1", "x", (lo, hi))
    g2 = O.Gauss("mu2", "s2", "x", (lo, hi))
    c = O.Chebyshev(["c1", "c2"], "x", (lo, hi))
    m = O.SumPDF([g1, g2, c], fracs=["f0", "f1", "f2"])
    names = ["f0", "f1", "f2", "mu1", "s1", "mu2", "s2", None, "c1", "c2"]
    slots = [0.3, 0.35, 0.35, -0.8, 1.1, 1.7, 2.2, 1.0, 0.1, -0.2]
    return Case("rtc_gauss_gauss_cheb2", nodes, names, slots, [x], m, ["x"], kernel="Sum<Gauss<0>,Gauss<0>,Cheb<0,2>>")


def nested_case(n=50_000, seed=8):
    """Sum of two 2-D products (signal x signal + background x background), in the catalogue."""
    rng = np.random.default_rng(seed)
    x = _trunc(rng, lambda k: np.concatenate([rng.normal(0.5, 1.0, k // 2), rng.uniform(-5, 5, k // 2)]), -5, 5, n)
    y = _trunc(rng, lambda k: rng.exponential(1 / 0.5, k), 0, 6, n)
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.PROD, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, -5, 5, 0, 0, 0),
        Z.Node(Z.EXP, 0, 1, 0, 0, 6, 0, 0, 0.0),
        Z.Node(Z.PROD, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.CHEB, 0, 0, 1, -5, 5, -5, 5, 0),
        Z.Node(Z.EXP, 0, 1, 0, 0, 6, 0, 0, 0.0),
    ]
    g = O.Gauss("mu", "sigma", "x", (-5, 5))
    e1 = O.Exponential("l1", "y", (0, 6))
    c = O.Chebyshev(["c1"], "x", (-5, 5))
    e2 = O.Exponential("l2", "y", (0, 6))
    m = _SumOfProducts([O.ProductPDF([g, e1]), O.ProductPDF([c, e2])], ["f0", "f1"])
    names = ["f0", "f1", "mu", "sigma", "l1", None, "c1", "l2"]
    slots = [0.45, 0.55, 0.4, 1.1, -0.6, 1.0, 0.1, -0.3]
    return Case("nested_sum_of_products", nodes, names, slots, [x, y], m, ["x", "y"],
                kernel="Sum<Prod<Gauss<0>,Expo<1>>,Prod<Cheb<0,1>,Expo<1>>>")


def gcb_exp_case(n=100_000, seed=9):
    """Generalized (two-sided) CrystalBall + exponential background."""
    rng = np.random.default_rng(seed)
    lo, hi = 5000.0, 5600.0
    sig = _trunc(rng, lambda k: np.concatenate([_cb_sample(rng, k // 2, 5280, 18, 1.4, 3), 2 * 5280 - _cb_sample(rng, k // 2, 5280, 24, 1.1, 4)]), lo, hi, int(0.4 * n))
    bkg = _trunc(rng, lambda k: lo + rng.exponential(1 / 0.002, k), lo, hi, n - sig.size)
    x = rng.permutation(np.concatenate([sig, bkg]))
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GCB, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.EXP, 0, 0, 0, lo, hi, 0, 0, 0.5 * (lo + hi)),
    ]
    names = ["f0", "f1", "mu", "sl", "al", "nl", "sr", "ar", "nr", "lam"]
    g = O.GeneralizedCB("mu", "sl", "al", "nl", "sr", "ar", "nr", "m", (lo, hi))
    m = O.SumPDF([g, O.Exponential("lam", "m", (lo, hi))], fracs=["f0", "f1"])
    slots = [0.35, 0.65, 5277.0, 20.0, 1.2, 2.5, 22.0, 1.3, 3.5, -0.0025]
    return Case("gcb_exp", nodes, names, slots, [x], m, ["m"], kernel="Sum<GCBall<0>,Expo<0>>")


def poly_case(kind, n=100_000, seed=10):
    """Gauss + Legendre / Chebyshev2 / Hermite / Laguerre background (degree 3: outside the catalogue -> NVRTC; degree 2:
    catalogue)."""
    rng = np.random.default_rng(seed)
    lo, hi = -6.0, 8.0
    x = _trunc(rng, lambda k: np.concatenate([rng.normal(1, 1.2, k // 3), rng.uniform(lo, hi, k - k // 3)]), lo, hi, n)
    zk, ocls, tname, deg = {"legendre": (Z.LEGENDRE, O.Legendre, "Legd", 3), "cheb2": (Z.CHEB2, O.Chebyshev2, "Chb2", 2),
                            "hermite": (Z.HERMITE, O.Hermite, "Herm", 3), "laguerre": (Z.LAGUERRE, O.Laguerre, "Lagr", 3)}[kind]
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(zk, 0, 0, deg, lo, hi, lo, hi, 0),
    ]
    cn = [f"c{k + 1}" for k in range(deg)]
    m = O.SumPDF([O.Gauss("mu", "sigma", "x", (lo, hi)), ocls(cn, "x", (lo, hi))], fracs=["f0", "f1"])
    names = ["f0", "f1", "mu", "sigma", None, *cn]
    slots = [0.3, 0.7, 0.8, 1.3, 1.0, *([0.12, -0.07, 0.05][:deg])]
    return Case(f"gauss_{kind}", nodes, names, slots, [x], m, ["x"], kernel=f"Sum<Gauss<0>,{tname}<0,{deg}>>")


def get_exp_case(generalized, n=100_000, seed=13):
    """(Generalized) GaussExpTail + exponential background."""
    rng = np.random.default_rng(seed)
    lo, hi = -12.0, 10.0
    sig = _trunc(rng, lambda k: np.concatenate([rng.normal(0.5, 1.2, k // 2), 0.5 - rng.exponential(2.0, k // 4), 0.5 + rng.exponential(2.5, k // 4)]), lo, hi, int(0.5 * n))
    bkg = _trunc(rng, lambda k: lo + rng.exponential(1 / 0.1, k), lo, hi, n - sig.size)
    x = rng.permutation(np.concatenate([sig, bkg]))
    if generalized:
        leaf = Z.Node(Z.GGET, 0, 0, 0, lo, hi, 0, 0, 0)
        o = O.GeneralizedGaussExpTail("mu", "sl", "al", "sr", "ar", "x", (lo, hi))
        names, vals, kern = ["mu", "sl", "al", "sr", "ar"], [0.4, 1.3, 1.1, 1.0, 1.4], "GGETail<0>"
    else:
        leaf = Z.Node(Z.GET, 0, 0, 0, lo, hi, 0, 0, 0)
        o = O.GaussExpTail("mu", "s", "a", "x", (lo, hi))
        names, vals, kern = ["mu", "s", "a"], [0.4, 1.3, 1.1], "GETail<0>"
    nodes = [Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0), leaf, Z.Node(Z.EXP, 0, 0, 0, lo, hi, 0, 0, 0.5 * (lo + hi))]
    m = O.SumPDF([o, O.Exponential("lam", "x", (lo, hi))], fracs=["f0", "f1"])
    return Case("gget_exp" if generalized else "get_exp", nodes, ["f0", "f1", *names, "lam"], [0.45, 0.55, *vals, -0.08],
                [x], m, ["x"], kernel=f"Sum<{kern},Expo<0>>")


def bernstein_case(n=100_000, seed=15):
    """Gauss + Bernstein background (degree 3: in the catalogue)."""
    rng = np.random.default_rng(seed)
    lo, hi = -6.0, 8.0
    x = _trunc(rng, lambda k: np.concatenate([rng.normal(1, 1.2, k // 3), rng.uniform(lo, hi, k - k // 3)]), lo, hi, n)
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.BERNSTEIN, 0, 0, 3, lo, hi, lo, hi, 0),
    ]
    cn = ["b0", "b1", "b2", "b3"]
    m = O.SumPDF([O.Gauss("mu", "sigma", "x", (lo, hi)), O.Bernstein(cn, "x", (lo, hi))], fracs=["f0", "f1"])
    return Case("gauss_bernstein", nodes, ["f0", "f1", "mu", "sigma", *cn], [0.3, 0.7, 0.8, 1.3, 0.9, 1.2, 0.7, 1.05], [x], m, ["x"],
                kernel="Sum<Gauss<0>,Bern<0,3>>")


def tgauss_case(n=100_000, seed=16):
    """TruncatedGauss + exponential: a third of the events lie outside [low, high], where only the background contributes."""
    rng = np.random.default_rng(seed)
    lo, hi = -6.0, 8.0
    sig = _trunc(rng, lambda k: rng.normal(1.0, 1.5, k), -1.0, 3.5, int(0.4 * n))
    bkg = _trunc(rng, lambda k: lo + rng.exponential(1 / 0.15, k), lo, hi, n - sig.size)
    x = rng.permutation(np.concatenate([sig, bkg]))
    nodes = [
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.TGAUSS, 0, 0, 0, lo, hi, 0, 0, 0),
        Z.Node(Z.EXP, 0, 0, 0, lo, hi, 0, 0, 0.5 * (lo + hi)),
    ]
    tg = O.TruncatedGauss("mu", "sigma", "low", "high", "x", (lo, hi))
    m = O.SumPDF([tg, O.Exponential("lam", "x", (lo, hi))], fracs=["f0", "f1"])
    return Case("tgauss_exp", nodes, ["f0", "f1", "mu", "sigma", "low", "high", "lam"], [0.35, 0.65, 0.8, 1.7, -1.0, 3.5, -0.12],
                [x], m, ["x"], kernel="Sum<TGauss<0>,Expo<0>>")


class _SumOfProducts:
    """SumPDF over N-D ProductPDFs (functor.py:197-206: sum_i frac_i * pdf_i.pdf(x))."""

    def __init__(self, pdfs, fracs):
        self.pdfs, self.fr = pdfs, fracs

    def pdf(self, B, x, P, norm=None):
        tot = None
        for p, f in zip(self.pdfs, self.fr):
            t = p.pdf(B, x, P) * P[f]
            tot = t if tot is None else tot + t
        return tot


ALL_CASES = {
    "gauss": gauss_case,
    "cb_exp": cb_exp_case,
    "gauss_cheb": gauss_cheb_case,
    "gauss_cheb_weighted": lambda **kw: gauss_cheb_case(weighted=True, seed=6, **kw),
    "product": product_case,
    "rtc": rtc_case,
    "nested": nested_case,
    "gcb_exp": gcb_exp_case,
    "legendre": lambda **kw: poly_case("legendre", **kw),
    "cheb2": lambda **kw: poly_case("cheb2", **kw),
    "hermite": lambda **kw: poly_case("hermite", **kw),
    "laguerre": lambda **kw: poly_case("laguerre", **kw),
    "get_exp": lambda **kw: get_exp_case(False, **kw),
    "gget_exp": lambda **kw: get_exp_case(True, **kw),
    "bernstein": bernstein_case,
    "tgauss": tgauss_case,
}
