"""Exception names used as part of the zfit API on this path (``src/zfit/util/exception.py``).

Only the names the hot path, its validation and the minimizers' control flow use are provided.
"""


class ZfitNotImplementedError(NotImplementedError):
    pass


class StandardControlFlow(Exception):
    pass


class FunctionNotImplemented(ZfitNotImplementedError):
    pass


class SpecificFunctionNotImplemented(FunctionNotImplemented):
    pass


class AnalyticNotImplemented(ZfitNotImplementedError):
    pass


class AnalyticIntegralNotImplemented(AnalyticNotImplemented):
    pass


class NormNotImplemented(StandardControlFlow):
    pass


class MultipleLimitsNotImplemented(StandardControlFlow):
    pass


class MaximumIterationReached(StandardControlFlow):
    pass


class ExtendedPDFError(Exception):
    pass


class AlreadyExtendedPDFError(ExtendedPDFError):
    pass


class NotExtendedPDFError(ExtendedPDFError):
    pass


class IntentionAmbiguousError(Exception):
    pass


class IncompatibleError(Exception):
    pass


class ShapeIncompatibleError(IncompatibleError):
    pass


class ObsIncompatibleError(IncompatibleError):
    pass


class SpaceIncompatibleError(IncompatibleError):
    pass


class LimitsIncompatibleError(IncompatibleError):
    pass


class ModelIncompatibleError(IncompatibleError):
    pass


class ParamNameNotUniqueError(Exception):
    pass


class ParameterNotIndependentError(Exception):
    pass


class NotMinimizedError(Exception):
    pass


class BreakingAPIChangeError(Exception):
    pass


class DerivativeCalculationError(ValueError):
    pass


class OutsideLimitsError(Exception):
    pass


class NameAlreadyTakenError(Exception):
    pass


class ModelNotOnDeviceError(NotImplementedError):
    """The model tree contains a PDF for which no fused CUDA kernel exists (no CPU fallback is provided)."""


class FailMinimizeNaN(Exception):
    pass
