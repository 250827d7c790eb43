"""CPU oracle for the unbinned-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / torch-CPU fp64) of the reference's
algorithm for ``UnbinnedNLL`` / ``ExtendedUnbinnedNLL`` value and gradient.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product
(``zfit_b200``) never imports it and never falls back to it.

Parity status
-------------
The reference (zfit 0.28.0) is pure Python on top of TensorFlow and
TensorFlow-Probability, neither of which is installed in this image, so the
reference itself cannot be executed here ("unbuildable": no TF wheel, no
network).  The oracle is therefore pinned against

* outputs of the reference's OWN function definitions for every formula zfit
  itself owns on this path: ``tests/golden/make_reference_formulas.py`` lifts the
  unmodified source of the CrystalBall / DoubleCB / GeneralizedCB / GaussExpTail
  shapes and integrals (``models/physics.py``), the recursive polynomials,
  Bernstein and their integrals (``models/polynomials.py``), the Exponential's
  value, shift and integral (``models/basic.py``), the NLL assembly
  ``_unbinned_nll_tf`` / ``_nll_calc_unbinned_tf`` (``core/loss.py``), ``SumPDF._pdf``
  / ``ProductPDF._pdf`` (``models/functor.py``) and ``z.safe_where`` out of
  ``/root/reference`` with ``ast`` and runs it over numpy stand-ins of ``znp`` /
  ``tf`` (60 definitions, the two loss bodies ``UnbinnedNLL._loss_func_watched`` and
  ``ExtendedUnbinnedNLL._loss_func`` included); the vectors are committed as
  ``tests/golden/reference_formulas.npz`` and ``tests/test_reference_formulas_cpu.py``
  holds the oracle (<= 2e-14), the product's host prologue (<= 1e-13) and the
  host-compiled device code (<= 2e-13) to them;
* the reference's only stored numeric golden on this path,
  ``tests/truth/data/deterministic_fit_data.npz`` (copied to
  ``tests/golden/``): at the stored fitted parameters the oracle's gradient
  must vanish to Minuit's EDM tolerance and the yields must sum to N;
* the known-answer checks the reference's own tests use
  (``scipy.stats.crystalball`` shape, ``tests/test_pdf_cb.py:129-147``; the
  hand formula for a Gauss sum, ``tests/test_pdfs.py:53-72``; analytic vs
  numeric integrals, ``tests/test_pdf_cb.py:53-78``,
  ``tests/test_pdf_polynomials.py``);
* independent re-derivations (scipy.integrate.quad, numpy.polynomial.chebyshev,
  central differences, torch autograd on the same op graph).

Arithmetic that lives in un-vendored third-party code (``tfp.distributions.Normal``,
``tfp.math.reduce_kahan_sum``, ``tf.nn.log_poisson_loss``, ``tf.math.erf``; pins are
ranges: tensorflow>=2.16.2,<3.0, tensorflow-probability>=0.24,<1.0, no lock file)
is restated from its published formulas: for those kernels -- and only those --
**parity with the TF implementation is unpinned** at the 1e-12 level (no stored
NLL values exist in the reference); it is adjudicated against an
extended-precision evaluation instead.
"""
