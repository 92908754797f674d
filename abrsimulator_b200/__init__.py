"""abrsimulator_b200 — B200-native batched ABR environment and MPC controller.

Drop-in for the hot path of Elliotshui/ABRSimulator (``Simulator.py`` environment loop and ``mpc.py``
controller) behind the C-ABI of ``include/abr_b200.h``.  Importing the package does not need a GPU;
every compute call does (there is no CPU fallback).
"""
from .datamodel import Chunk, MPD, QOEMetric, ChunkInfo, NetworkInfo, load_network_trace, load_mpd_file  # noqa: F401
from ._lib import AbrError, default_params, launch_count, library_path  # noqa: F401

__all__ = ["Chunk", "MPD", "QOEMetric", "ChunkInfo", "NetworkInfo", "BatchedABREnv", "MPCBitrateController",
           "Simulator", "RandomPolicy", "BufferBasedPolicy", "FixedPolicy", "AbrError", "default_params",
           "launch_count", "library_path", "load_network_trace", "load_mpd_file"]


def __getattr__(name):   # torch-dependent modules are imported lazily
    if name == "BatchedABREnv":
        from .env import BatchedABREnv
        return BatchedABREnv
    if name == "MPCBitrateController":
        from .mpc import MPCBitrateController
        return MPCBitrateController
    if name in ("Simulator", "RandomPolicy", "BufferBasedPolicy", "FixedPolicy"):
        from . import simulator
        return getattr(simulator, name)
    raise AttributeError(name)
