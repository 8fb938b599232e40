"""crdmodel_b200 — B200 (sm_100a) implementation of CRDModel's time-stepping hot path.

The product is crdmodel_b200/lib/libcrd_b200.so (C ABI: include/crd_b200.h); this package is its
ctypes mirror.  Importing it never needs a GPU; creating a Context does, and there is no CPU path.
"""
from ._lib import CrdError, IcParams, Params, LIB_PATH, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (ARITH_EXACT, ARITH_FAST, ARK_NORMAL, ARK_ONE_STEP, MODELS, ARKodeSolver, Context, Grid,  # noqa: F401
                  NVector, Snapshot, decomp_phi, make_params)
