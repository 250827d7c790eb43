"""Random model trees shared by the GPU parity test (``test_random_trees_gpu.py``) and its CPU twin on the host-compiled
device code (``test_random_trees_cpu.py``): the same seeds build the same trees in both."""

import numpy as np

import zfit_b200 as zfit

OBS = {"x": (-3.0, 4.0), "y": (0.0, 5.0), "z": (-2.0, 2.0)}


class _Builder:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.tag = f"rt{seed}"
        self.k = 0

    def par(self, value, lo, hi):
        self.k += 1
        return zfit.Parameter(f"p{self.k}_{self.tag}", value, lo, hi)

    def leaf(self, obs):
        lo, hi = OBS[obs]
        sp = zfit.Space(obs, lo, hi)
        r = self.rng
        mid, wid = 0.5 * (lo + hi), hi - lo
        kind = r.choice(["gauss", "expo", "cb", "cheb", "legendre", "get", "gcb", "bernstein", "tgauss"])
        if kind in ("expo", "cheb", "legendre", "bernstein"):
            mu = sg = None
        else:
            mu = self.par(mid + 0.2 * wid * r.uniform(-1, 1), lo, hi)
            sg = self.par(wid * r.uniform(0.08, 0.3), 0.01 * wid, wid)
        if kind == "gauss":
            return zfit.pdf.Gauss(mu, sg, sp)
        if kind == "expo":
            return zfit.pdf.Exponential(self.par(r.uniform(-0.8, 0.8), -3, 3), sp)
        if kind == "cb":
            return zfit.pdf.CrystalBall(mu, sg, self.par(r.uniform(0.5, 2.0), 0.1, 5), self.par(r.uniform(1.5, 4), 1.05, 20), sp)
        if kind in ("cheb", "legendre"):
            deg = int(r.integers(1, 4))
            cs = [self.par(r.uniform(-0.15, 0.15), -1, 1) for _ in range(deg)]
            return (zfit.pdf.Chebyshev if kind == "cheb" else zfit.pdf.Legendre)(sp, cs)
        if kind == "bernstein":
            return zfit.pdf.Bernstein(sp, [self.par(r.uniform(0.5, 1.5), 0.01, 5) for _ in range(int(r.integers(2, 5)))])
        if kind == "tgauss":  # the truncation keeps most of the range, so part of the events fall outside it
            return zfit.pdf.TruncatedGauss(mu, sg, self.par(lo + 0.15 * wid, lo, mid), self.par(hi - 0.1 * wid, mid, hi), sp)
        if kind == "get":
            return zfit.pdf.GaussExpTail(mu, sg, self.par(r.uniform(0.5, 2.0), 0.1, 5), sp)
        return zfit.pdf.GeneralizedCB(
            mu, sg, self.par(r.uniform(0.5, 2), 0.1, 5), self.par(r.uniform(1.5, 4), 1.05, 20),
            self.par(wid * r.uniform(0.08, 0.3), 0.01 * wid, wid), self.par(r.uniform(0.5, 2), 0.1, 5),
            self.par(r.uniform(1.5, 4), 1.05, 20), sp)

    def tree(self, obs, depth):
        r = self.rng
        if depth == 0 or r.uniform() < 0.25:
            if len(obs) == 1:
                return self.leaf(obs[0])
            return zfit.pdf.ProductPDF([self.tree([o], 0) for o in obs])
        if len(obs) > 1 and r.uniform() < 0.5:  # product over a random split of the observables
            cut = int(r.integers(1, len(obs)))
            return zfit.pdf.ProductPDF([self.tree(obs[:cut], depth - 1), self.tree(obs[cut:], depth - 1)])
        k = int(r.integers(2, 4))
        kids = [self.tree(obs, depth - 1) for _ in range(k)]
        raw = r.dirichlet(np.ones(k))
        fracs = [self.par(float(f), 0, 1) for f in raw[:-1]]
        return zfit.pdf.SumPDF(kids, fracs=fracs)


def build_case(seed, n):
    """-> (builder, loss): a random tree over 1-3 observables with uniform events (odd n exercises the scalar tail event)
    and, for odd seeds, weights."""
    b = _Builder(seed)
    obs = list(OBS)[: 1 + seed % 3]
    model = b.tree(obs, depth=2)
    cols = {o: b.rng.uniform(*OBS[o], n) for o in obs}
    space = zfit.Space(obs[0], *OBS[obs[0]])
    for o in obs[1:]:
        space = space * zfit.Space(o, *OBS[o])
    weights = b.rng.normal(1.0, 0.5, n) if seed % 2 else None
    data = zfit.Data(cols, obs=space, weights=weights)
    return b, zfit.loss.UnbinnedNLL(model, data)
