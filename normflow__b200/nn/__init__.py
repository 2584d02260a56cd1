"""Flow layers (reference src/nn/__init__.py): trailing underscore = forward/backward
carry the log-Jacobian."""
from ._core import Module_, ModuleList_  # noqa: F401

from .scalar.modules import ConvAct, SplineNet, Conv4d  # noqa: F401
from .scalar.modules_ import DistConvertor_, Identity_, Clone_  # noqa: F401
from .scalar.modules_ import Expit_, Logit_, SplineNet_, ScaleNet_, SgnBiasNet_  # noqa: F401
from .scalar.modules_ import UnityDistConvertor_, PhaseDistConvertor_  # noqa: F401

from .scalar.couplings_ import Coupling_, ShiftCoupling_, AffineCoupling_  # noqa: F401
from .scalar.couplings_ import RQSplineCoupling_, MultiRQSplineCoupling_  # noqa: F401
from .scalar.cntr_couplings_ import DirectCntrCoupling_, CntrCoupling_  # noqa: F401
from .scalar.cntr_couplings_ import CntrShiftCoupling_, CntrAffineCoupling_  # noqa: F401
from .scalar.cntr_couplings_ import CntrRQSplineCoupling_, CntrMultiRQSplineCoupling_  # noqa: F401

from .scalar.fftflow_ import FFTNet_, IPSD, FreeScalar  # noqa: F401
from .scalar.meanfield_ import MeanFieldNet_  # noqa: F401
from .scalar.psd_ import PSDBlock_  # noqa: F401
