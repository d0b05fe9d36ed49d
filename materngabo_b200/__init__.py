"""materngabo_b200: the MaternGaBO kernel-matrix / GP-posterior hot path on B200 (sm_100a).

Drop-in module names mirror BoManifolds/kernel_utils: ``kernels_sphere``, ``kernels_so``, ``kernels_torus``,
``kernels_hyperbolic``, ``kernels_spd``; ``posterior`` holds the fused exact-GP posterior and its
multi-GPU row sharding.  Everything computes in float64 on CUDA through libmgb.so (include/mgb.h); there is
no CPU fallback.
"""
__all__ = ["kernels_sphere", "kernels_so", "kernels_torus", "kernels_hyperbolic", "kernels_spd", "posterior"]
