"""Score-network side of the INDM hot path (drop-in for the reference's `models` package on that path)."""
from . import utils  # noqa: F401
from . import ncsnpp  # noqa: F401  (registers 'ncsnpp')
