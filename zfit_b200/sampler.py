"""On-device toy generation (reference: ``accept_reject_sample`` / ``UniformSampleAndWeights``,
``src/zfit/core/sample.py:148-566``, driven by ``BasePDF.create_sampler`` / ``SamplerData.resample``,
``core/basepdf.py`` / ``core/data.py:1563-1599``).

Accept-reject with uniform proposals, entirely on the GPU: torch's generator fills a fixed proposal buffer per observable,
the model tree's ``pdf_kernel`` (the same compiled tree the likelihood uses) writes the densities into a device buffer
(``znll_pdf_device``), and the accepted events are compacted by the library's cut pass (``znll_cut_events`` with the
``u < p`` predicate).  The only host traffic per batch is the
accepted count.  The envelope starts at 1.2 x the largest density seen and is raised -- and the sample restarted -- if a
proposal ever exceeds it, so the result is an unbiased sample (the reference re-weights instead, ``sample.py:440-505``).
"""

from __future__ import annotations

import numpy as np

from . import _znll as Z


class DeviceSampler:
    MIN_PROPOSALS = 1 << 14
    MAX_PROPOSALS = 1 << 26  # 512 MB per buffer

    def __init__(self, pdf, limits, *, device=None, batch=None):
        import torch

        from . import distributed as _dist

        self.torch = torch
        self.pdf = pdf
        self.limits = limits
        self.device = _dist.local_device() if device is None else device
        self.fixed_batch = None if batch is None else int(batch)  # a caller-chosen proposal count (tests); else adaptive
        self.lower = [float(v) for v in np.atleast_1d(limits._lower).ravel()]
        self.upper = [float(v) for v in np.atleast_1d(limits._upper).ravel()]
        nodes, self.slot_params = pdf._lower(limits.obs, None)
        self.plan = Z.Plan(1, device=self.device)
        self.plan.set_model(0, nodes)
        self.batch = 0  # proposals per pass = size of the buffers below
        self.cols, self.pvals, self.unif = [], None, None
        self.generator = torch.Generator(device=torch.device("cuda", self.device))
        self.generator.manual_seed(int(np.random.randint(0, 2**31 - 1)))
        self.pmax = None
        self.accept_rate = None  # accepted / proposed of the last draw

    def seed(self, seed):
        self.generator.manual_seed(int(seed))

    def _slots(self, values):
        from .pdf import _maybe_set

        with _maybe_set(dict(values), self.pdf):
            return np.array([p.value() for p in self.slot_params], dtype=np.float64)

    def _size_buffers(self, n):
        """Proposal buffers for draws of about n events: one pass should be enough (n / acceptance, +15 %), so that a toy
        costs a handful of launches whatever its size; buffers only grow."""
        torch = self.torch
        if self.fixed_batch is not None:
            want = self.fixed_batch
        else:
            rate = self.accept_rate if self.accept_rate else 0.25
            want = int(min(self.MAX_PROPOSALS, max(self.MIN_PROPOSALS, 1.15 * n / max(rate, 1e-3) + 1024)))
            step = 1 << max(14, want.bit_length() - 4)  # sixteenths of the next power of two: regrowth stays rare
            want = -(-want // step) * step
        if want <= self.batch:
            return
        dev = torch.device("cuda", self.device)
        self.batch = want
        self.cols = [torch.empty(want, dtype=torch.float64, device=dev) for _ in self.limits.obs]
        self.pvals = torch.empty(want, dtype=torch.float64, device=dev)
        self.unif = torch.empty(want, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(self.device)
        self.plan._keep.clear()  # the previous, smaller buffers
        self.plan.bind_device_ptrs(0, [c.data_ptr() for c in self.cols], 0, want, keepalive=self.cols)

    def draw(self, n, values):
        """-> dict obs -> device column with exactly n events drawn at the parameter values ``values``."""
        torch = self.torch
        if n is None:
            if not self.pdf.is_extended:
                msg = "n has to be given for a non-extended pdf"
                raise ValueError(msg)
            from .pdf import _maybe_set

            with _maybe_set(dict(values), self.pdf):
                n = int(np.random.poisson(self.pdf.get_yield().value()))
        n = int(n)
        self._size_buffers(n)
        slots = self._slots(values)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pmax = self.pmax
        while True:  # restarted only if the envelope turned out too low
            kept = [[] for _ in self.cols]
            have, proposed, restart = 0, 0, False
            while have < n:
                for c, lo, up in zip(self.cols, self.lower, self.upper):
                    c.uniform_(lo, up, generator=self.generator)
                self.plan.pdf_values_device(0, slots, self.pvals.data_ptr(), stream=stream)
                top = self.pvals.max()  # stays on the device; read after the compaction's own synchronisation
                if pmax is None:
                    pmax, top = 1.2 * float(top), None
                self.unif.uniform_(0.0, pmax, generator=self.generator)
                # accept u < p and compact the accepted proposals in one pass (znll_cut_events; only the count comes back)
                out = [torch.empty_like(c) for c in self.cols]
                n_acc = Z.cut_events(self.device, [c.data_ptr() for c in self.cols], [c.data_ptr() for c in out], 0, 0,
                                     self.lower, self.upper, self.batch, less_a=self.unif.data_ptr(),
                                     less_b=self.pvals.data_ptr(), stream=stream)
                if top is not None and float(top) > pmax:  # a proposal above the envelope: raise it and start over
                    pmax, restart = 1.2 * float(top), True
                    break
                for k, c in enumerate(out):
                    kept[k].append(c[:n_acc])
                have += n_acc
                proposed += self.batch
            if not restart:
                break
        self.pmax = pmax
        if proposed:
            self.accept_rate = max(have / proposed, 1e-6)
        if not kept[0]:  # n == 0
            dev = torch.device("cuda", self.device)
            return {o: torch.empty(0, dtype=torch.float64, device=dev) for o in self.limits.obs}
        if len(kept[0]) == 1:  # the usual case: one pass was enough -- its compacted buffer (copied when much larger than n)
            take = (lambda t: t[:n].clone()) if self.batch > 2 * n else (lambda t: t[:n])
            return {o: take(parts[0]) for o, parts in zip(self.limits.obs, kept)}
        return {o: torch.cat(parts)[:n].contiguous() for o, parts in zip(self.limits.obs, kept)}

    def close(self):
        self.plan.close()
