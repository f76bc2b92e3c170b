"""grid2op_b200 — a B200-native batched power-flow Backend behind ``grid2op.Backend.Backend``."""
__version__ = "0.1.0"
