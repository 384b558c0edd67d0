from .etkf import ETKF  # noqa: F401
from .letkf import LETKF  # noqa: F401
from .ketkf import KETKF, LKETKF  # noqa: F401
from .base import StateError, ObservationError  # noqa: F401
from .ienks import VarAssimilation, IEnKSTransform, IEnKSBundle, LocalizedIEnKSTransform, LocalizedIEnKSBundle  # noqa: F401
