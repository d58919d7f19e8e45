"""museinference.jl_b200 — B200-native backend for the per-simulation hot path of MUSE.

Public surface mirrors ``MuseInference.jl``'s exports (/root/reference/src/MuseInference.jl:30):
``SimpleMuseProblem, MuseResult, muse, muse!, get_J!, get_H!`` (Python: ``muse_``, ``get_J_``,
``get_H_``).  The directory name contains a dot, so import it through the top-level shim
``museinference_jl_b200`` (repo root).
"""
from ._capi import MuseBackendError, load_library, library_path          # noqa: F401
from ._build import build_library                                        # noqa: F401
from .backend import B200Backend                                         # noqa: F401
from .parallel import LocalPool, ShardPool, block_partition              # noqa: F401
from .problem import AbstractMuseProblem, SimpleMuseProblem, BaseDraws, FlatPrior, NormalPrior   # noqa: F401
from .muse import MuseResult, muse, muse_, get_J_, get_H_, finalize_result_, central_fdm, AdaptedFDM, SimpleCovariance   # noqa: F401

globals()["muse!"] = muse_
globals()["get_J!"] = get_J_
globals()["get_H!"] = get_H_
