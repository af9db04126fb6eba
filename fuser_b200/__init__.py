"""fuser_b200 — B200-native normalization / reduction kernels behind nvFuser's executor boundary.

Scope: the ONE hot path named in BASELINE.json — LayerNorm / RMSNorm / softmax
forward and backward segments (InnerPersistent / InnerOuterPersistent / Reduction
in the reference's scheduler terms) as hand-written sm_100a kernels reached
through a C ABI (``include/b200norm.h``). See DESIGN.md.
"""
__version__ = "0.1.0"
