from .chomp import CHOMP  # noqa: F401
from .gpmp2 import GPMP2  # noqa: F401
from .stoch_gpmp import StochGPMP  # noqa: F401
from .stomp import STOMP  # noqa: F401
from .mppi import MPPI  # noqa: F401
from .hybrid_planner import HybridPlanner  # noqa: F401
