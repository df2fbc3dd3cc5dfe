"""passport-zk-circuits_b200: B200-native batched witness generator and R1CS checker
for the passport-zk-circuits circom circuits (hot path only)."""
__version__ = "0.1.0"
