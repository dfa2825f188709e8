"""ctypes binding of libphoregen_b200.so (C ABI declared in include/phoregen_b200.h).

There is no CPU fallback: if the library is missing the import raises, loudly.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHOREGEN_B200_LIB: load another build of the same library (A/B comparisons of kernel variants on one box)
LIB_PATH = os.environ.get("PHOREGEN_B200_LIB") or os.path.join(_HERE, "libphoregen_b200.so")


class PhoreGenLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise PhoreGenLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m phoregen_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    return ctypes.CDLL(LIB_PATH)


class _Library:
    """The shared library, mapped on first use.  Importing the package (e.g. to enumerate a PhoreDiff state_dict, as
    bench.py's CPU reference arm does) does not map the CUDA library into the process; the first call of any entry point
    does, binds every signature below, and raises loudly if the library or a symbol is missing."""

    def __init__(self):
        self._cdll = None

    def _bind(self):
        if self._cdll is None:
            cdll = _load()
            for name, (res, args) in _SIGS.items():
                fn = getattr(cdll, name)          # AttributeError here = header/library drift; never silently ignored
                fn.restype = res
                fn.argtypes = args
            self._cdll = cdll
        return self._cdll

    @property
    def loaded(self):
        return self._cdll is not None

    def __getattr__(self, name):
        return getattr(self._bind(), name)


lib = _Library()

_P = c_void_p
_SIGS = {
    "pg_version": (c_int, []),
    "pg_last_error": (c_char_p, []),
    "pg_weight_slot_count": (c_int, []),
    "pg_weight_slot_name": (c_char_p, [c_int]),
    "pg_weight_slot_numel": (c_int64, [c_int]),
    "pg_model_create": (c_int, [POINTER(_P), _P, POINTER(c_int64), c_int]),
    "pg_model_destroy": (None, [_P]),
    "pg_plan_workspace_bytes": (c_int64, [c_int, POINTER(c_int32), POINTER(c_int32)]),
    "pg_plan_create": (c_int, [POINTER(_P), c_int, POINTER(c_int32), POINTER(c_int32), c_int, _P, _P, c_int64, _P]),
    "pg_plan_destroy": (None, [_P]),
    "pg_plan_num_ligand_atoms": (c_int64, [_P]),
    "pg_plan_num_phore_nodes": (c_int64, [_P]),
    "pg_plan_num_bond_edges": (c_int64, [_P]),
    "pg_plan_num_knn_edges": (c_int64, [_P]),
    "pg_plan_num_triplets": (c_int64, [_P]),
    "pg_plan_export_bond_edges": (c_int, [_P, _P, _P, _P]),
    "pg_plan_export_triplets": (c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "pg_knn_graph": (c_int, [_P, _P, c_int, _P, _P]),
    "pg_denoiser_forward": (c_int, [_P] * 10),
    "pg_phore_encode": (c_int, [_P] * 6),
    "pg_phorediff_forward": (c_int, [_P] * 13),
    "pg_categorical_step": (c_int, [c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_uint64, c_uint32, _P, _P, _P, _P, _P, _P, _P]),
    "pg_position_step": (c_int, [c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_uint64, c_uint32, _P, _P, _P, _P, c_int, _P, _P, _P]),
    "pg_sample_init": (c_int, [c_int, c_int, _P, _P, _P, c_uint64, c_uint32, _P, _P, _P, _P, _P, _P]),
    "pg_position_init": (c_int, [c_int, _P, _P, c_uint64, c_uint32, _P, c_int, _P, _P, _P, _P]),
    "pg_atom_count": (c_int, [_P, _P, _P, _P, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "pg_guidance_grad": (c_int, [_P, _P, _P, c_int, c_float, c_float, _P, _P, c_int, _P]),
    "pg_gemm_k128": (c_int, [c_int, c_int, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int, _P]),
    "pg_plan_ligand_graph": (_P, [_P]),
    "pg_plan_edge_graph": (_P, [_P]),
    "pg_plan_kernel_launches": (c_int64, [_P]),
    "pg_plan_timing_enable": (c_int, [_P, c_int]),
    "pg_plan_timing_read": (c_int, [_P, c_int, POINTER(ctypes.c_double), POINTER(c_int64)]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def last_error():
    msg = lib.pg_last_error()
    return msg.decode() if msg else ""


def check(rc, what):
    if rc != 0:
        raise PhoreGenLibraryError(f"{what} failed (code {rc}): {last_error()}")


def slot_table():
    n = lib.pg_weight_slot_count()
    return [(lib.pg_weight_slot_name(i).decode(), int(lib.pg_weight_slot_numel(i))) for i in range(n)]
