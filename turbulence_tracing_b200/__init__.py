"""turbulence_tracing_b200 -- B200-native (sm_100a CUDA) implementation of the ray-integration hot
path of jdhare/turbulence_tracing behind the reference's own Python API.

    from turbulence_tracing_b200 import particle_tracker as pt
    from turbulence_tracing_b200 import ray_transfer_matrix as rtm
    from turbulence_tracing_b200 import turboGen as tg
"""
from . import _lib                                   # noqa: F401
from ._lib import DeviceArray, TTError               # noqa: F401
from . import particle_tracker, ray_transfer_matrix, turboGen, calculate_spectrum_3d   # noqa: F401

__all__ = ["particle_tracker", "ray_transfer_matrix", "turboGen", "calculate_spectrum_3d", "DeviceArray", "TTError"]
