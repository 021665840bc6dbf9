"""fdfd-b200: B200-native matrix-free FDFD operator + Krylov solve behind the MaxwellFDFD.jl model API.

Import as `maxwellfdm_jl_b200` (shim at the repo root).  The compute path is libfdfd_b200.so
(hand-written sm_100a CUDA behind a C ABI, include/fdfd_b200.h); importing this package does not
load it, the first operator construction does and raises if it is missing.
"""
from .grid import EE, HH, PRIM, DUAL, Grid, PMLParam, create_stretched_dl
from .sources import PointSrc, PlaneSrc, distweights
from .model import (Model, ModelFull, ModelTE, ModelTM, ModelTEM, set_wpml, set_boundft, set_Npml, set_kbloch, create_e_mikL, clear_srcs,
                    add_srce, add_srcm, create_srcs, create_stretched_dls, create_paramops, create_curls, ParamOp, CurlOp,
                    create_A, create_b, create_linsys,
                    h_from_e, e_from_h, create_Mcs, solve, field_arr2vec, field_vec2arr)
from .operator import FdfdOperator, MultiGpuOperator, comm_unique_id, partition, halo_plan
from .reduced import ReducedOperator
from .shapes import Box, Ball, Sphere, Cylinder, add_obj, clear_objs, calc_matparams, calc_matparams_array
from . import _lib

__all__ = [n for n in dir() if not n.startswith("_")]
