"""ctypes binding of include/picsp_b200.h (one function per ABI entry point)."""
from __future__ import annotations

import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
# PICSP_B200_LIB: another build of the same library (profiles/build_variant.sh), for same-box A/B measurements only
LIB_PATH = os.environ.get("PICSP_B200_LIB") or os.path.join(PKG, "libpicsp_b200.so")
HEADER = os.path.join(ROOT, "include", "picsp_b200.h")
HOST_HEADER = os.path.join(ROOT, "include", "picsp_b200_host.h")

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


class PicspError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"picsp_b200 error {code}: {msg}")
        self.code = code


class CRunConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nTimeSteps", "numxCells", "numyCells", "nParticlesI", "nParticlesE",
                                         "dumpPeriod", "solverType", "loadType")] + \
               [(n, C.c_double) for n in ("timeStep", "stepSize", "massI", "massE", "chargeE", "density", "vthE", "vthI",
                                          "driftE", "driftI", "ion_spwt", "electron_spwt", "omega_pe", "Lambda_D")]


class CParams(C.Structure):
    _fields_ = [("numxCells", C.c_int32), ("numyCells", C.c_int32), ("stepSize", C.c_double), ("timeStep", C.c_double),
                ("solverType", C.c_int32), ("flags", C.c_int32), ("charge", C.c_double * 2), ("mass", C.c_double * 2),
                ("spwt", C.c_double * 2), ("capacity", C.c_int64 * 2), ("device", C.c_int32), ("parts", C.c_int32)]


def abi_symbols():
    """Every function name declared in include/*.h."""
    names = set()
    for h in (HEADER, HOST_HEADER):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(picsp_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


_lib = None


def load_library():
    """Loads the in-tree CUDA library.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `python -m picsp_b200.build` "
                                "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    ctx = C.c_void_p
    sig = {
        "picsp_abi_version": ([], C.c_int),
        "picsp_create": ([C.POINTER(CParams), C.POINTER(ctx)], C.c_int),
        "picsp_destroy": ([ctx], None),
        "picsp_last_error": ([], C.c_char_p),
        "picsp_sync": ([ctx], C.c_int),
        "picsp_species_upload": ([ctx, C.c_int, _dp, _dp, _dp, _dp, C.c_int64], C.c_int),
        "picsp_species_download": ([ctx, C.c_int, _dp, _dp, _dp, _dp], C.c_int),
        "picsp_species_count": ([ctx, C.c_int, _i64p], C.c_int),
        "picsp_species_download_rows": ([ctx, C.c_int, _dp], C.c_int),
        "picsp_grid_upload": ([ctx, C.c_int, _dp], C.c_int),
        "picsp_grid_download": ([ctx, C.c_int, _dp], C.c_int),
        "picsp_dump_begin": ([ctx, _dp, _dp, _dp, _dp, _dp, _dp], C.c_int),
        "picsp_dump_wait": ([ctx], C.c_int),
        "picsp_host_alloc": ([C.c_size_t], C.c_void_p),
        "picsp_host_free": ([C.c_void_p], None),
        "picsp_deposit": ([ctx, C.c_int], C.c_int),
        "picsp_compute_rho": ([ctx], C.c_int),
        "picsp_solve": ([ctx], C.c_int),
        "picsp_solve_spectral": ([ctx], C.c_int),
        "picsp_solve_sor": ([ctx, _i64p, _dp], C.c_int),
        "picsp_solve_status": ([ctx, _i64p, _dp], C.c_int),
        "picsp_compute_ef": ([ctx], C.c_int),
        "picsp_push": ([ctx, C.c_int], C.c_int),
        "picsp_rewind": ([ctx, C.c_int], C.c_int),
        "picsp_bootstrap": ([ctx], C.c_int),
        "picsp_step": ([ctx, C.c_int], C.c_int),
        "picsp_compute_ke": ([ctx, C.c_int, _dp], C.c_int),
        "picsp_delta_phi": ([ctx, _dp, _dp], C.c_int),
        "picsp_repush_count": ([ctx, C.c_int, _i64p], C.c_int),
        "picsp_straggler_count": ([ctx, C.c_int, _i64p], C.c_int),
        "picsp_set_sort_period": ([ctx, C.c_int, C.c_int], C.c_int),
        "picsp_set_cell_sort_period": ([ctx, C.c_int, C.c_int], C.c_int),
        "picsp_set_bank_order": ([ctx, C.c_int, C.c_int], C.c_int),
        "picsp_set_deposit_aggregation": ([ctx, C.c_int, C.c_int], C.c_int),
        "picsp_comm_unique_id": ([C.c_void_p], C.c_int),
        "picsp_comm_attach": ([ctx, C.c_void_p, C.c_int, C.c_int], C.c_int),
        "picsp_comm_barrier": ([ctx], C.c_int),
        "picsp_comm_peer_reduction": ([ctx], C.c_int),
        "picsp_species_fill_synthetic": ([ctx, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_double], C.c_int),
        "picsp_profile_enable": ([ctx, C.c_int], C.c_int),
        "picsp_profile_get": ([ctx, C.c_int, _dp, _i64p], C.c_int),
        "picsp_profile_reset": ([ctx], C.c_int),
        "picsp_kernel_launches": ([ctx, _i64p], C.c_int),
        "picsp_parts": ([ctx, C.POINTER(C.c_int)], C.c_int),
        "picsp_spectral_engine": ([ctx, C.POINTER(C.c_int)], C.c_int),
        "picsp_fft_plan_query": ([C.c_int, C.POINTER(C.c_int32)], C.c_int),
        "picsp_host_parse_ini": ([C.c_char_p, C.POINTER(CRunConfig), C.c_int], C.c_int),
        "picsp_host_ini_dump": ([C.c_char_p, C.c_char_p, C.c_int64], C.c_int64),
        "picsp_host_loader_create": ([C.c_uint32], C.c_void_p),
        "picsp_host_loader_destroy": ([C.c_void_p], None),
        "picsp_host_loader_fill": ([C.c_void_p, C.POINTER(CRunConfig), C.c_int, _dp, _dp, _dp, _dp], C.c_int),
        "picsp_host_run": ([C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int], C.c_int),
        "picsp_host_run_ranked": ([C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "picsp_host_h5_open": ([C.c_char_p], C.c_void_p),
        "picsp_host_h5_group": ([C.c_void_p, C.c_char_p], C.c_int),
        "picsp_host_h5_dataset_f64": ([C.c_void_p, C.c_char_p, _dp, C.c_uint64, C.c_uint64], C.c_int),
        "picsp_host_h5_attr_f64": ([C.c_void_p, C.c_char_p, C.c_double], C.c_int),
        "picsp_host_h5_attr_i32": ([C.c_void_p, C.c_char_p, C.c_int32], C.c_int),
        "picsp_host_h5_close": ([C.c_void_p], C.c_int),
    }
    for name, (args, res) in sig.items():
        if os.environ.get("PICSP_B200_LIB") and not hasattr(L, name):
            continue                      # an older build under A/B test: entry points added since are simply absent
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    L._picsp_signatures = sig
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PicspError(rc, load_library().picsp_last_error().decode(errors="replace"))
