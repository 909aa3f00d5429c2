"""ctypes binding of the C ABI in ``include/rscm_b200.h``.

This is the only place the Python host layer touches native code.  There is no
CPU fallback: if ``librscm_b200.so`` has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C rscm_b200/csrc``) importing this
module raises, and creating an ensemble without a CUDA device raises
``RuntimeError`` with the engine's message.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librscm_b200.so")

ABI_VERSION = 1

OK = 0
EINVAL, EUNSUPPORTED, ENODEVICE, ECUDA, ENOMEM, ECOMM = -1, -2, -3, -4, -5, -6
UNIQUE_ID_BYTES = 128

# component kinds
TWO_LAYER, CARBON_CYCLE, CO2_ERF, GHG_FORCING = 1, 2, 3, 5
OZONE_FORCING, AEROSOL_DIRECT, AEROSOL_INDIRECT, CLIMATE_UDEB = 6, 7, 8, 9
FOUR_BOX_OHU, OCEAN_SURFACE_PP, CO2_BUDGET, TERRESTRIAL_CARBON, CH4_CHEMISTRY, N2O_CHEMISTRY = 10, 11, 12, 13, 14, 15
OCEAN_CARBON = 16
HALOCARBON_CHEMISTRY = 17
# grids / aggregate ops / sources
SCALAR, FOUR_BOX, HEMISPHERIC = 0, 1, 2
AGG_SUM, AGG_MEAN, AGG_WEIGHTED = 0, 1, 2
SRC_EXOGENOUS, SRC_OWN_STATE, SRC_UPSTREAM = 0, 1, 2
# priors
PRIOR_NONE, PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_LOGNORMAL = 0, 1, 2, 3
PRIOR_BOUND_NORMAL, PRIOR_BOUND_LOGNORMAL, PRIOR_BOUND_UNIFORM = 4, 5, 6


class ComponentDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_params", C.c_int32), ("params", C.POINTER(C.c_double))]


class SchemaVariable(C.Structure):
    _fields_ = [("name", C.c_char_p), ("grid", C.c_int32)]


class AggregateDesc(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("op", C.c_int32),
        ("grid", C.c_int32),
        ("n_contributors", C.c_int32),
        ("contributors", C.POINTER(C.c_char_p)),
        ("weights", C.POINTER(C.c_double)),
    ]


class InitialValue(C.Structure):
    _fields_ = [("name", C.c_char_p), ("value", C.c_double)]


class UnitFactor(C.Structure):
    _fields_ = [("component", C.c_int32), ("variable", C.c_char_p), ("factor", C.c_double)]


class GraphDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_components", C.c_int32),
        ("components", C.POINTER(ComponentDesc)),
        ("has_schema", C.c_int32),
        ("n_schema_variables", C.c_int32),
        ("schema_variables", C.POINTER(SchemaVariable)),
        ("n_aggregates", C.c_int32),
        ("aggregates", C.POINTER(AggregateDesc)),
        ("n_initial_values", C.c_int32),
        ("initial_values", C.POINTER(InitialValue)),
        ("n_unit_factors", C.c_int32),
        ("unit_factors", C.POINTER(UnitFactor)),
        ("four_box_weights", C.POINTER(C.c_double)),
        ("hemispheric_weights", C.POINTER(C.c_double)),
        ("n_times", C.c_int32),
        ("time_bounds", C.POINTER(C.c_double)),
        ("compute_dtype", C.c_int32),
        ("device", C.c_int32),
    ]


class Obs(C.Structure):
    _fields_ = [("variable", C.c_int32), ("time_index", C.c_int32), ("value", C.c_double), ("sigma", C.c_double)]


class Prior(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("reserved", C.c_int32),
        ("a", C.c_double),
        ("b", C.c_double),
        ("low", C.c_double),
        ("high", C.c_double),
    ]


class LogpostSummary(C.Structure):
    _fields_ = [
        ("max_logpost", C.c_double),
        ("argmax", C.c_int64),
        ("sum_finite", C.c_double),
        ("n_finite", C.c_int64),
        ("n_runs", C.c_int64),
    ]


class SamplerState(C.Structure):
    _fields_ = [
        ("positions", C.c_void_p), ("ld", C.c_int64), ("n_cols", C.c_int32), ("reserved", C.c_int32), ("n_walkers", C.c_int64),
        ("logpost", C.c_void_p), ("proposals", C.c_void_p), ("z", C.c_void_p), ("logpost_new", C.c_void_p * 2),
        ("n_accepted", C.c_void_p), ("a", C.c_double), ("seed", C.c_uint64), ("first_iteration", C.c_uint32), ("thin", C.c_uint32),
        ("chain_positions", C.c_void_p), ("chain_logpost", C.c_void_p), ("chain_capacity", C.c_int64),
    ]


# every symbol include/rscm_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_PD = C.c_void_p  # double* passed as integer addresses (host numpy or device data_ptr)
SYMBOLS = {
    "rscm_b200_ensemble_create": (C.c_int, [C.POINTER(GraphDesc), C.POINTER(_H)]),
    "rscm_b200_ensemble_destroy": (None, [_H]),
    "rscm_b200_last_error": (C.c_char_p, [_H]),
    "rscm_b200_last_global_error": (C.c_char_p, []),
    "rscm_b200_abi_version": (C.c_int, []),
    "rscm_b200_device_count": (C.c_int, []),
    "rscm_b200_n_variables": (C.c_int, [_H]),
    "rscm_b200_variable_name": (C.c_char_p, [_H, C.c_int]),
    "rscm_b200_variable_grid": (C.c_int, [_H, C.c_int]),
    "rscm_b200_variable_is_endogenous": (C.c_int, [_H, C.c_int]),
    "rscm_b200_variable_index": (C.c_int, [_H, C.c_char_p]),
    "rscm_b200_n_exogenous": (C.c_int, [_H]),
    "rscm_b200_exogenous_variable": (C.c_int, [_H, C.c_int]),
    "rscm_b200_n_nodes": (C.c_int, [_H]),
    "rscm_b200_execution_order": (C.c_int, [_H, C.POINTER(C.c_int), C.c_int]),
    "rscm_b200_variable_source": (C.c_int, [_H, C.c_int, C.c_char_p]),
    "rscm_b200_program_signature": (C.c_char_p, [_H]),
    "rscm_b200_program_is_jit": (C.c_int, [_H]),
    "rscm_b200_time_index": (C.c_int, [_H, C.c_double]),
    "rscm_b200_bind_parameters": (C.c_int, [_H, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_int]),
    "rscm_b200_select_outputs": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32]),
    "rscm_b200_output_rows": (C.c_int64, [_H]),
    "rscm_b200_run_device": (C.c_int, [_H, _PD, C.c_int64, C.c_int, _PD, C.c_int64, _PD, C.c_void_p, C.c_void_p]),
    "rscm_b200_run_host": (C.c_int, [_H, _PD, C.c_int64, C.c_int, _PD, C.c_int64, _PD, C.c_void_p]),
    "rscm_b200_set_target": (C.c_int, [_H, C.POINTER(Obs), C.c_int64, C.c_int]),
    "rscm_b200_set_priors": (C.c_int, [_H, C.POINTER(Prior), C.c_int]),
    "rscm_b200_logpost_device": (C.c_int, [_H, _PD, C.c_int64, C.c_int, _PD, C.c_int64, _PD, C.c_void_p, C.c_void_p]),
    "rscm_b200_logpost_host": (C.c_int, [_H, _PD, C.c_int64, C.c_int, _PD, C.c_int64, _PD, C.POINTER(LogpostSummary)]),
    "rscm_b200_launch_count": (C.c_int64, [_H]),
    "rscm_b200_shared_bytes": (C.c_int64, [_H, C.c_int]),
    "rscm_b200_kernel_ms": (C.c_double, [_H, C.c_int]),
    "rscm_b200_measure_fma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "rscm_b200_interpolate_device": (C.c_int, [_PD, C.c_int64, _PD, C.c_int64, C.c_int, _PD, C.c_int64, C.c_int, _PD, C.c_void_p]),
    "rscm_b200_device_math": (C.c_int, [C.c_int, _PD, C.c_int64, _PD, C.c_void_p]),
    "rscm_b200_member_quantiles": (C.c_int, [_PD, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.c_int, _PD, C.c_void_p]),
    "rscm_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rscm_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_H)]),
    "rscm_b200_comm_destroy": (None, [_H]),
    "rscm_b200_comm_last_error": (C.c_char_p, [_H]),
    "rscm_b200_comm_rank": (C.c_int, [_H]),
    "rscm_b200_comm_world": (C.c_int, [_H]),
    "rscm_b200_comm_peer_access": (C.c_int, [_H]),
    "rscm_b200_comm_shard": (C.c_int, [_H, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rscm_b200_comm_symmetric_alloc": (C.c_int, [_H, C.c_size_t, C.POINTER(C.c_void_p)]),
    "rscm_b200_allgather_f64": (C.c_int, [_H, _PD, C.c_int64, _PD, C.c_void_p]),
    "rscm_b200_logpost_sharded_device": (C.c_int, [_H, _H, _PD, C.c_int64, C.c_int, _PD, C.c_int64, _PD, C.c_void_p]),
    "rscm_b200_comm_check": (C.c_int, [_H]),
    "rscm_b200_sampler_iterate": (C.c_int, [_H, _H, C.POINTER(SamplerState), _PD, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "rscm_b200_stretch_propose": (C.c_int, [_PD, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_uint64,
                                            C.c_uint32, _PD, C.c_int64, _PD, C.c_void_p]),
    "rscm_b200_stretch_accept": (C.c_int, [_PD, C.c_int64, C.c_int, C.c_int64, C.c_int64, _PD, C.c_int64, _PD, _PD, _PD, C.c_uint64,
                                           C.c_uint32, _PD, C.c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"rscm_b200: native library {LIB_PATH} not found. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.rscm_b200_abi_version() != ABI_VERSION:
        raise ImportError("rscm_b200: ABI version mismatch between _ffi.py and librscm_b200.so")
    return lib


lib = _load()


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rscm_b200 error {code}: {message}")
        self.code = code


def check(code: int, handle=None) -> None:
    if code == OK:
        return
    msg = lib.rscm_b200_last_error(handle) if handle else lib.rscm_b200_last_global_error()
    raise EngineError(code, (msg or b"").decode("utf-8", "replace"))


def check_comm(code: int, comm=None) -> None:
    if code == OK:
        return
    msg = lib.rscm_b200_comm_last_error(comm) if comm else lib.rscm_b200_last_global_error()
    raise EngineError(code, (msg or b"").decode("utf-8", "replace"))
