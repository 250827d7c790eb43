"""zfit_b200 -- B200-native unbinned-likelihood hot path behind zfit's API.

``import zfit_b200 as zfit`` gives the subset of the zfit surface that the unbinned NLL path touches
(``Space``, ``Parameter``, ``Data``, ``zfit.pdf.*``, ``zfit.loss.*``, ``zfit.minimize.*``, ``zfit.param``),
with ``loss.value`` / ``value_gradient`` evaluated by hand-written sm_100a CUDA kernels (``csrc/``)
through the C ABI in ``include/znll.h``.
"""

from __future__ import annotations

import types as _types

from . import data as _data_mod
from . import distributed, exception, settings
from . import loss as _loss_mod
from . import minimize as _min_mod
from . import parameter as _param_mod
from . import pdf as _pdf_mod
from .constraint import GaussianConstraint, LogNormalConstraint, PoissonConstraint, SimpleConstraint
from .data import Data, SamplerData, convert_to_data
from .parameter import ComposedParameter, ConstantParameter, Parameter, convert_to_parameter
from .settings import run, ztypes
from .space import Space, combine_spaces, convert_to_space

__version__ = "0.1.0"

pdf = _types.SimpleNamespace(
    BasePDF=_pdf_mod.BasePDF,
    Gauss=_pdf_mod.Gauss,
    Exponential=_pdf_mod.Exponential,
    CrystalBall=_pdf_mod.CrystalBall,
    Chebyshev=_pdf_mod.Chebyshev,
    Legendre=_pdf_mod.Legendre,
    Chebyshev2=_pdf_mod.Chebyshev2,
    Hermite=_pdf_mod.Hermite,
    Laguerre=_pdf_mod.Laguerre,
    Bernstein=_pdf_mod.Bernstein,
    TruncatedGauss=_pdf_mod.TruncatedGauss,
    DoubleCB=_pdf_mod.DoubleCB,
    GeneralizedCB=_pdf_mod.GeneralizedCB,
    GaussExpTail=_pdf_mod.GaussExpTail,
    GeneralizedGaussExpTail=_pdf_mod.GeneralizedGaussExpTail,
    SumPDF=_pdf_mod.SumPDF,
    ProductPDF=_pdf_mod.ProductPDF,
)
loss = _types.SimpleNamespace(
    BaseLoss=_loss_mod.BaseLoss,
    UnbinnedNLL=_loss_mod.UnbinnedNLL,
    ExtendedUnbinnedNLL=_loss_mod.ExtendedUnbinnedNLL,
    SimpleLoss=_loss_mod.SimpleLoss,
)
minimize = _types.SimpleNamespace(
    **{k: getattr(_min_mod, k) for k in dir(_min_mod) if k.startswith("Scipy") and k != "ScipyBaseMinimizer"},
    Minuit=_min_mod.Minuit,
    BaseMinimizer=_min_mod.BaseMinimizer,
    EDM=_min_mod.EDM,
    DefaultStrategy=_min_mod.DefaultStrategy,
    PushbackStrategy=_min_mod.PushbackStrategy,
    DefaultToyStrategy=_min_mod.DefaultToyStrategy,
    LossEval=_min_mod.LossEval,
)
result = _types.SimpleNamespace(FitResult=_min_mod.FitResult)
param = _types.SimpleNamespace(
    Parameter=Parameter,
    ComposedParameter=ComposedParameter,
    ConstantParameter=ConstantParameter,
    convert_to_parameter=convert_to_parameter,
    set_values=_param_mod.set_values,
    assign_values=_param_mod.assign_values,
)
data = _types.SimpleNamespace(Data=Data, SamplerData=SamplerData, concat=_data_mod.concat, convert_to_data=convert_to_data)
constraint = _types.SimpleNamespace(
    GaussianConstraint=GaussianConstraint,
    PoissonConstraint=PoissonConstraint,
    LogNormalConstraint=LogNormalConstraint,
    SimpleConstraint=SimpleConstraint,
)
dimension = _types.SimpleNamespace(combine_spaces=combine_spaces)
