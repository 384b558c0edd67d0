from .gaspari_cohn import GaspariCohn, GaspariCohnInf, BaseLocalization  # noqa: F401
from .metrics import AbsDistance1D, PeriodicDistance1D, EuclideanDistance, HaversineDistance, ZeroDistance, ProductDistance  # noqa: F401
