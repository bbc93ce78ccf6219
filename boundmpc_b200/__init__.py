"""boundmpc_b200: B200-native batched solver for BoundMPC's per-step OCP."""
__version__ = "0.1.0"
