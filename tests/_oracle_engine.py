"""Test double with the ``CudaEngine`` interface, backed by the CPU oracle (tests only).

It lets the host layer (parameter ordering, chain rule through composed parameters, extended term,
offsets, minimizer protocol) run on a CPU box, and gives the reference-side numbers for the GPU parity
tests.  The lowered ``znll_node`` trees are translated back into oracle PDF trees whose parameters are
the slots ``s0, s1, ...`` -- so the oracle differentiates with respect to exactly what the kernel does.
"""

from __future__ import annotations

import numpy as np

from oracle import nll as ON
from oracle import pdfs as O
from zfit_b200 import _znll as Z
from zfit_b200 import distributed as D


def _translate(nodes, obs, i, slot, in_prod):
    nd = nodes[i]
    if nd.kind == Z.SUM:
        fr = [f"s{slot + k}" for k in range(nd.n_children)]
        slot += nd.n_children
        kids = []
        j = i + 1
        for _ in range(nd.n_children):
            kid, j, slot = _translate(nodes, obs, j, slot, False)
            kids.append(kid)
        return _Sum(kids, fr), j, slot
    if nd.kind == Z.PROD:
        kids = []
        j = i + 1
        for _ in range(nd.n_children):
            kid, j, slot = _translate(nodes, obs, j, slot, True)
            kids.append(kid)
        return O.ProductPDF(kids), j, slot
    o = obs[nd.column]
    lim = (nd.norm_lo, nd.norm_hi)
    if nd.kind == Z.GAUSS:
        return O.Gauss(f"s{slot}", f"s{slot + 1}", o, lim), i + 1, slot + 2
    if nd.kind == Z.EXP:
        return O.Exponential(f"s{slot}", o, lim), i + 1, slot + 1
    if nd.kind == Z.CB:
        return O.CrystalBall(f"s{slot}", f"s{slot + 1}", f"s{slot + 2}", f"s{slot + 3}", o, lim), i + 1, slot + 4
    if nd.kind in (Z.CHEB, Z.LEGENDRE, Z.CHEB2, Z.HERMITE, Z.LAGUERRE):
        assert (nd.space_lo, nd.space_hi) == lim, "oracle translation needs norm == space for polynomials"
        names = [f"s{slot + k}" for k in range(nd.degree + 1)]
        cls = {Z.CHEB: O.Chebyshev, Z.LEGENDRE: O.Legendre, Z.CHEB2: O.Chebyshev2, Z.HERMITE: O.Hermite,
               Z.LAGUERRE: O.Laguerre}[nd.kind]
        return cls(names[1:], o, lim, coeff0=names[0]), i + 1, slot + nd.degree + 1
    if nd.kind == Z.BERNSTEIN:
        assert (nd.space_lo, nd.space_hi) == lim, "oracle translation needs norm == space for polynomials"
        return O.Bernstein([f"s{slot + k}" for k in range(nd.degree + 1)], o, lim), i + 1, slot + nd.degree + 1
    if nd.kind == Z.TGAUSS:
        return O.TruncatedGauss(*[f"s{slot + k}" for k in range(4)], o, lim), i + 1, slot + 4
    if nd.kind == Z.GET:
        return O.GaussExpTail(f"s{slot}", f"s{slot + 1}", f"s{slot + 2}", o, lim), i + 1, slot + 3
    if nd.kind == Z.GGET:
        names = [f"s{slot + k}" for k in range(5)]
        return O.GeneralizedGaussExpTail(*names, o, lim), i + 1, slot + 5
    if nd.kind == Z.GCB:
        names = [f"s{slot + k}" for k in range(7)]
        return O.GeneralizedCB(*names, o, lim), i + 1, slot + 7
    raise AssertionError(nd.kind)


class _Sum:
    def __init__(self, pdfs, fracs):
        self.pdfs, self.fr = pdfs, fracs
        self.limits = getattr(pdfs[0], "limits", None)

    def pdf(self, B, x, P, norm=None):
        tot = None
        for p, f in zip(self.pdfs, self.fr):
            t = p.pdf(B, x, P) * P[f]
            tot = t if tot is None else tot + t
        return tot

    def integrate(self, B, P, limits, norm=False):  # fractions sum to one in the lowered cases
        tot = None
        for f in self.fr:
            tot = P[f] if tot is None else tot + P[f]
        return tot


class OracleEngine:
    def __init__(self, models, datas, fit_ranges, chunk=None):
        """``chunk``: evaluate at most this many events per tape and sum the pieces (the NLL and its gradient are
        plain sums over events) -- lets ``bench.py --impl reference`` run the full 1e8 events within host RAM."""
        self.chunk = None if not chunk else int(chunk)
        self.slot_params = []
        self.channels = []
        datas = [d if r is None else d.with_obs(r, guarantee_limits=False) for d, r in zip(datas, fit_ranges)]
        for model, data, rng in zip(models, datas, fit_ranges):
            nodes, slots = model._lower(data.obs, rng)
            tree, end, ns = _translate(nodes, data.obs, 0, 0, False)
            assert end == len(nodes) and ns == len(slots)
            cols = {o: np.asarray(data.value(o)) for o in data.obs}
            self.channels.append((tree, cols, data.weights, len(self.slot_params), ns))
            self.slot_params.extend(slots)
        self.n_local = [d.num_entries for d in datas]
        self.samplesize = [d.samplesize for d in datas]
        self.n_entries = list(self.n_local)
        self.distributed = D.is_distributed()
        if self.distributed:
            import torch

            t = torch.tensor(self.samplesize + [float(n) for n in self.n_local], dtype=torch.float64)
            D.all_reduce_sum_(t, deterministic=True)
            k = len(models)
            self.samplesize = [float(v) for v in t[:k]]
            self.n_entries = [int(round(float(v))) for v in t[k:]]
        self.kernel_names = ["oracle"] * len(models)
        self.ncalls = 0

    def eval(self, slots, grad=True, log_offset=0.0, kahan=False):
        self.ncalls += 1
        values, gs = [], np.zeros(len(self.slot_params))
        for tree, cols, w, off, ns in self.channels:
            names = [f"s{k}" for k in range(ns)]
            theta = np.asarray(slots[off : off + ns], dtype=np.float64)
            ww = None if w is None else [w]
            if len(next(iter(cols.values()))) == 0:
                values.append(0.0)
                continue
            n = len(next(iter(cols.values())))
            step = n if not self.chunk else self.chunk
            v = 0.0
            for lo in range(0, n, step):
                if step >= n:
                    cc, cw = cols, ww
                else:
                    cc = {k: a[lo : lo + step] for k, a in cols.items()}
                    cw = None if w is None else [np.asarray(w)[lo : lo + step]]
                if grad:
                    vi, g = ON.value_and_gradient([tree], [cc], theta, names, weights=cw, log_offset=log_offset)
                    gs[off : off + ns] += g
                else:
                    vi = float(ON.unbinned_nll([tree], [cc], dict(zip(names, theta)), weights=cw, log_offset=log_offset))
                v += vi
            values.append(v)
        values = np.asarray(values, dtype=np.float64)
        if self.distributed:
            import torch

            t = torch.from_numpy(np.concatenate([values, gs]))
            D.all_reduce_sum_(t, deterministic=True)
            values, gs = t[: len(values)].numpy().copy(), t[len(values) :].numpy().copy()
        return values, (gs if grad else None)

    def eval_cov(self, slots):
        """Reference-style per-event Jacobian (errors.py:406-452) -> sum_i w_i^2 [g_i;1][g_i;1]^T per channel."""
        import torch

        from oracle.backend import TorchBackend

        B = TorchBackend()
        mats = []
        for tree, cols, w, off, ns in self.channels:
            names = [f"s{k}" for k in range(ns)]
            theta = torch.tensor(np.asarray(slots[off : off + ns], dtype=np.float64), requires_grad=True)
            x = {k: B.asarray(v) for k, v in cols.items()}

            def logp(th):
                return torch.log(tree.pdf(B, x, {n: th[i] for i, n in enumerate(names)}))

            # forward mode: ns tangent passes, O(N * ns) memory (reverse mode would need an N x N basis)
            J = torch.autograd.functional.jacobian(logp, theta, vectorize=True, strategy="forward-mode").detach().numpy()  # (N, ns)
            ww = np.ones(J.shape[0]) if w is None else np.asarray(w)
            G = np.concatenate([J, np.ones((J.shape[0], 1))], axis=1)
            mats.append((G * (ww**2)[:, None]).T @ G)
        if self.distributed:
            t = torch.from_numpy(np.concatenate([m.ravel() for m in mats]))
            D.all_reduce_sum_(t, deterministic=True)
            out, o = [], 0
            for m in mats:
                out.append(t[o : o + m.size].numpy().reshape(m.shape).copy())
                o += m.size
            mats = out
        return mats

    def launch_count(self):
        return 0

    def close(self):
        pass


def use_oracle(loss):
    loss._set_engine_factory(OracleEngine)
    return loss
