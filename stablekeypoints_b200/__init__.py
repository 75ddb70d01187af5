"""stablekeypoints_b200: B200-native (sm_100a) implementation of the StableKeypoints per-image hot path.

Host side mirrors the reference's Python surface (ptp_utils / optimize / optimize_token / eval /
invertable_transform); tensor math is hand-written CUDA behind the C ABI of include/skp_b200.h.
"""
__all__ = ["ptp_utils", "optimize", "optimize_token", "eval", "invertable_transform", "ops", "sd15_engine", "compat"]
