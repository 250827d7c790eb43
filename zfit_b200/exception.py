"""Exception types that are part of the zfit API on this path.

The reference keeps them in ``src/zfit/util/exception.py``; several are used as control flow by the model dispatch
(``NormNotImplemented`` & co., ``util/exception.py:192-248``) and by the minimizers (``MaximumIterationReached``).
Only the names the hot path, its input validation and the minimizer protocol actually raise or catch exist here;
they are declared from one table so that the hierarchy is visible at a glance.
"""

from __future__ import annotations

_TABLE = (
    # name, base, one-line meaning
    ("ZfitNotImplementedError", NotImplementedError, "base of the 'not implemented, try the fallback' family"),
    ("StandardControlFlow", Exception, "raised and caught as ordinary control flow, never an error by itself"),
    ("FunctionNotImplemented", "ZfitNotImplementedError", "a model function is missing; the caller falls back"),
    ("SpecificFunctionNotImplemented", "FunctionNotImplemented", "a specific (user-facing) function is missing"),
    ("AnalyticNotImplemented", "ZfitNotImplementedError", "no analytic expression available"),
    ("AnalyticIntegralNotImplemented", "AnalyticNotImplemented", "no analytic integral registered"),
    ("NormNotImplemented", "StandardControlFlow", "the function does not handle the `norm` argument itself"),
    ("MultipleLimitsNotImplemented", "StandardControlFlow", "the function does not handle several limits"),
    ("MaximumIterationReached", "StandardControlFlow", "LossEval exceeded its evaluation budget"),
    ("ExtendedPDFError", Exception, "problem with the yield of a PDF"),
    ("AlreadyExtendedPDFError", "ExtendedPDFError", "the PDF already has a yield"),
    ("NotExtendedPDFError", "ExtendedPDFError", "an extended quantity was requested from a PDF without yield"),
    ("IntentionAmbiguousError", Exception, "the input could mean two things; be explicit"),
    ("IncompatibleError", Exception, "two objects do not fit together"),
    ("ShapeIncompatibleError", "IncompatibleError", "array shapes do not match"),
    ("ObsIncompatibleError", "IncompatibleError", "observables do not match"),
    ("SpaceIncompatibleError", "IncompatibleError", "spaces do not match"),
    ("LimitsIncompatibleError", "IncompatibleError", "limits do not match or are missing"),
    ("ModelIncompatibleError", "IncompatibleError", "models cannot be combined this way"),
    ("ParamNameNotUniqueError", Exception, "two different parameters with one name inside one model / loss"),
    ("ParameterNotIndependentError", Exception, "tried to set a dependent (composed) parameter"),
    ("NotMinimizedError", Exception, "a fit result was required but the minimisation did not run"),
    ("BreakingAPIChangeError", Exception, "the call uses an API that was removed on purpose"),
    ("DerivativeCalculationError", ValueError, "a gradient entry is None"),
    ("OutsideLimitsError", Exception, "raw data lies outside the limits it was promised to respect"),
    ("NameAlreadyTakenError", Exception, "the name is in use"),
    ("FailMinimizeNaN", Exception, "too many NaNs in a row: the NaN strategy gives up"),
)

__all__ = [row[0] for row in _TABLE] + ["ModelNotOnDeviceError"]

for _name, _base, _doc in _TABLE:
    _parent = globals()[_base] if isinstance(_base, str) else _base
    globals()[_name] = type(_name, (_parent,), {"__doc__": _doc, "__module__": __name__})
del _name, _base, _doc, _parent


class ModelNotOnDeviceError(NotImplementedError):
    """The model tree contains something no fused CUDA kernel exists for.  There is no CPU fallback on purpose."""
