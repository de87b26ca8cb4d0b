"""rtgs — B200-native drop-in for the per-ray render path of fangjunzhou/rt-gaussian-splat-renderer.

Same Python surface as the reference package (``rtgs.scene.Scene``, ``rtgs.camera.Camera``,
``rtgs.ray_tracer.RayTracer``, ``rtgs.gaussian``, ``rtgs.ray``, ``rtgs.utils``); the device work is
hand-written sm_100a CUDA behind the C-ABI in include/rtgs_b200.h (no Taichi, no CPU fallback).
"""
import rtgs.utils  # noqa: F401  (the reference's __init__ does the same)

__version__ = "0.1.0"
