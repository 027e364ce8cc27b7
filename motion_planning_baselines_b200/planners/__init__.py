from .stoch_gpmp import StochGPMP  # noqa: F401
