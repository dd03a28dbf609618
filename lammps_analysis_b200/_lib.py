"""ctypes binding of libmdk.so (include/mdk.h).

The library is built in-tree by ``lammps_analysis_b200/csrc/build.sh`` (see
``__graft_entry__.build``).  There is no CPU fallback: if the shared object is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDK_LIB_PATH: load another build of the same library (A/B runs of compile-time tuning switches)
LIB_PATH = os.environ.get("MDK_LIB_PATH") or os.path.join(_HERE, "libmdk.so")

MDK_RDF_EXACT_DIV = 1
MDK_RDF_WRAPPED = 2
MDK_MAX_SPECIES = 8


class MdkError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_LL = C.c_longlong
_I = C.c_int
_F = C.c_float
_D = C.c_double

# name -> (argtypes) ; every function returns int except the two noted below
PROTOTYPES = {
    "mdk_version": [],
    "mdk_sm_count": [],
    "mdk_rdf_tile": [],
    "mdk_rdf_thresholds": [_F, _I, _P, _P],
    "mdk_rdf_pack": [_P, _LL, _LL, _LL, _LL, _P, _I, _P, _LL, _LL, _LL, _P],
    "mdk_gather_frames": [_P, _LL, _LL, _P, _I, _P, _P],
    "mdk_coord_extent": [_P, _I, _LL, _P, _P],
    "mdk_rdf_hist": [_P, _I, _LL, _P, _P, _I, _P, _F, _F, _I, _P, _P, _P, _P, _I, _P],
    "mdk_rdf_tie_count": [_P, _LL, _LL, _P, _F, _F, _I, _P, _I, _P, _P],
    "mdk_rdf_sort_workspace": [_I],
    "mdk_rdf_pack_sorted": [_P, _LL, _LL, _LL, _I, _LL, _P, _LL, _LL, _I, _P, _P, _LL, _P],
    "mdk_rdf_sort_batch_workspace": [_I, _I],
    "mdk_rdf_pack_sorted_batch": [_P, _LL, _LL, _LL, _I, _P, _I, _P, _LL, _LL, _I, _P, _P, _LL, _P],
    "mdk_rdf_bbox": [_P, _I, _LL, _P, _P],
    "mdk_adf_workspace": [_LL, _I, _P, _F],
    "mdk_adf_hist": [_P, _I, _LL, _P, _I, _P, _F, _I, _D, _D, _I, _P, _P, _P, _P, _LL, _P],
    "mdk_msd_windowed": [_P, _LL, _LL, _LL, _LL, _LL, _I, _I, _P, _I, _I, _P, _P],
    "mdk_msd_dense": [_P, _LL, _LL, _LL, _LL, _LL, _I, _I, _P, _P],
    "mdk_acf_lagprod": [_P, _LL, _LL, _LL, _LL, _LL, _I, _I, _P, _P],
    "mdk_acf_windows": [_P, _I, _I, _I, _I, _P, _P, _P],
    "mdk_unwrap": [_P, _LL, _LL, _P, _P, _I, _P, _P, _P],
    "mdk_unwrap_indices": [_P, _P, _LL, _P, _P, _P],
    "mdk_velocity_from_positions": [_P, _LL, _LL, _F, _P, _P],
    "mdk_ionic_current": [_P, _LL, _LL, _P, _I, _P, _P],
    "mdk_flux_sum": [_P, _LL, _LL, _I, _I, _P, _P, _P, _P],
    "mdk_thermal_flux": [_P, _P, _P, _P, _LL, _LL, _P, _P],
    "mdk_peak_fp32": [_I, _I, _P],
    "mdk_store_mapped": [_P, _P, _LL, _P],
    "mdk_lammps_scan": [C.c_char_p, _P, _P, _P, _P, _P, _I],
    "mdk_lammps_read": [C.c_char_p, _LL, _I, _I, _I, _LL, _P, _P],
}


def load():
    """Load libmdk.so once; raise MdkError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdkError(
            f"{LIB_PATH} not found: build it with lammps_analysis_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = _LL if name in ("mdk_rdf_sort_workspace", "mdk_rdf_sort_batch_workspace",
                                     "mdk_adf_workspace") else _I
    lib.mdk_last_error.argtypes = []
    lib.mdk_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mdk_last_error().decode(errors="replace")
        raise MdkError(f"{what} failed with code {rc}: {msg}")
