"""ctypes wrapper of the CPU ORACLE (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  See ``rscm_oracle.h`` for what
the oracle restates and for the parity-pinning statement.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

TWO_LAYER, CARBON_CYCLE, CO2_ERF, GHG_FORCING = 1, 2, 3, 5
OZONE_FORCING, AEROSOL_DIRECT, AEROSOL_INDIRECT, CLIMATE_UDEB = 6, 7, 8, 9
FOUR_BOX_OHU, OCEAN_SURFACE_PP, CO2_BUDGET, TERRESTRIAL_CARBON, CH4_CHEMISTRY, N2O_CHEMISTRY = 10, 11, 12, 13, 14, 15
OCEAN_CARBON = 16
HALOCARBON_CHEMISTRY = 17
SCALAR, FOUR_BOX, HEMISPHERIC = 0, 1, 2
AGG_SUM, AGG_MEAN, AGG_WEIGHTED = 0, 1, 2


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB


class Obs(C.Structure):
    _fields_ = [("variable", C.c_int32), ("time_index", C.c_int32), ("value", C.c_double), ("sigma", C.c_double)]


class Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("low", C.c_double), ("high", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_model_new.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_variable_name.restype = C.c_char_p
        L.orc_variable_offset.restype = C.c_int64
        L.orc_output_size.restype = C.c_int64
        L.orc_ln_pdf.restype = C.c_double
        L.orc_log_prior.restype = C.c_double
        L.orc_ln_likelihood.restype = C.c_double
        L.orc_compute_aggregate.restype = C.c_double
        L.orc_co2_erf.restype = C.c_double
        L.orc_co2_erf.argtypes = [C.c_double] * 3
        L.orc_rk4_steps.argtypes = [C.c_double] * 3
        L.orc_ghg_forcings.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleModel:
    """Builder + runner over the C oracle (mirrors the reference's ModelBuilder semantics)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_model_new())
        self._built = False

    def __del__(self):
        try:
            self.L.orc_model_free(self.h)
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())
        return rc

    def add_component(self, kind, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        return self._ck(self.L.orc_add_component(self.h, kind, _dp(p), p.size))

    def add_schema_variable(self, name, grid=SCALAR):
        self._ck(self.L.orc_add_schema_variable(self.h, name.encode(), grid))

    def add_aggregate(self, name, op, contributors, weights=None, grid=SCALAR):
        names = (C.c_char_p * len(contributors))(*[c.encode() for c in contributors])
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.L.orc_add_aggregate(self.h, name.encode(), op, grid, len(contributors), names, _dp(w)))

    def set_initial_value(self, name, v):
        self._ck(self.L.orc_set_initial_value(self.h, name.encode(), C.c_double(v)))

    def set_time_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        self.T = b.size - 1
        self._ck(self.L.orc_set_time_bounds(self.h, _dp(b), self.T))

    def set_exogenous(self, name, values, grid=SCALAR):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.L.orc_set_exogenous(self.h, name.encode(), grid, _dp(v)))

    def set_unit_factor(self, component, variable, factor):
        self._ck(self.L.orc_set_unit_factor(self.h, component, variable.encode(), C.c_double(factor)))

    def set_grid_weights(self, grid, weights):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.L.orc_set_grid_weights(self.h, grid, _dp(w)))

    def build(self):
        self._ck(self.L.orc_build(self.h))
        self._built = True
        self.names = [self.L.orc_variable_name(self.h, i).decode() for i in range(self.L.orc_n_variables(self.h))]
        self.grids = [self.L.orc_variable_grid(self.h, i) for i in range(len(self.names))]
        return self

    def regions(self, name):
        return {SCALAR: 1, FOUR_BOX: 4, HEMISPHERIC: 2}[self.grids[self.names.index(name)]]

    def is_endogenous(self, name):
        return bool(self.L.orc_variable_is_endogenous(self.h, self.names.index(name)))

    def execution_order(self):
        buf = (C.c_int * 64)()
        n = self.L.orc_execution_order(self.h, buf)
        return [buf[i] for i in range(n)]

    def variable_source(self, component, variable):
        return self.L.orc_variable_source(self.h, component, variable.encode())

    def time_index(self, t):
        return self.L.orc_time_index(self.h, C.c_double(t))

    def run(self):
        out = np.empty(self.L.orc_output_size(self.h))
        self._ck(self.L.orc_run(self.h, _dp(out)))
        res = {}
        for i, n in enumerate(self.names):
            off = self.L.orc_variable_offset(self.h, i)
            r = self.regions(n)
            blk = out[off:off + self.T * r].reshape(self.T, r)
            res[n] = blk[:, 0].copy() if r == 1 else blk.copy()
        return res

    def _bind(self, bindings):
        """bindings: list of lists of (component_index, param_index) or (-1, variable name) per column."""
        bc, bi, cols = [], [], []
        for j, targets in enumerate(bindings):
            for (c, p) in targets:
                bc.append(c)
                bi.append(self.names.index(p) if c == -1 else p)
                cols.append(j)
        return np.array(bc, dtype=np.int32), np.array(bi, dtype=np.int32), np.array(cols, dtype=np.int32)

    def _expand(self, bindings, params):
        """Duplicate columns so that each (target) has its own column (C API: one target per column)."""
        bc, bi, cols = self._bind(bindings)
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        pe = np.ascontiguousarray(p[:, cols]) if cols.size else np.zeros((p.shape[0], 0))
        return bc, bi, pe

    def run_batch(self, bindings, params, exo_names, scenarios, out_names, n_threads=0, want_status=False):
        bc, bi, pe = self._expand(bindings, params)
        M = pe.shape[0]
        exo = np.array([self.names.index(n) for n in exo_names], dtype=np.int32)
        S = 0 if scenarios is None else scenarios.shape[0]
        sc = None if scenarios is None else np.ascontiguousarray(scenarios, dtype=np.float64)
        ov = np.array([self.names.index(n) for n in out_names], dtype=np.int32)
        rows = sum(self.T * self.regions(n) for n in out_names)
        runs = max(S, 1) * M
        out = np.empty((rows, runs))
        status = np.zeros(runs, dtype=np.uint8)
        self._ck(self.L.orc_run_batch(self.h, pe.shape[1], _dp(bc), _dp(bi), _dp(pe), C.c_int64(M), exo.size, _dp(exo), _dp(sc),
                                      C.c_int64(S), ov.size, _dp(ov), _dp(out), _dp(status), n_threads))
        return (out, status) if want_status else out

    def split(self, out, out_names):
        res, row = {}, 0
        for n in out_names:
            r = self.regions(n)
            blk = out[row:row + self.T * r].reshape(self.T, r, out.shape[1])
            res[n] = blk[:, 0, :] if r == 1 else blk
            row += self.T * r
        return res

    def make_obs(self, observations):
        arr = (Obs * max(1, len(observations)))()
        for i, (name, time, value, sigma) in enumerate(observations):
            arr[i] = Obs(self.names.index(name), self.time_index(time), value, sigma)
        return arr

    @staticmethod
    def make_priors(priors):
        arr = (Prior * max(1, len(priors)))()
        for i, p in enumerate(priors):
            p = tuple(p) + (0.0,) * (5 - len(p))
            arr[i] = Prior(int(p[0]), 0, p[1], p[2], p[3], p[4])
        return arr

    def log_posterior_batch(self, bindings, params, exo_names, scenarios, priors, observations, normalize=False, n_threads=0):
        """NOTE: priors are per *user* column; columns feeding several slots are expanded
        for the C call, so the prior of a duplicated column is applied once (others NONE)."""
        bc, bi, cols = self._bind(bindings)
        p = np.ascontiguousarray(params, dtype=np.float64)
        pe = np.ascontiguousarray(p[:, cols])
        pr = []
        seen = set()
        for c in cols:
            if priors is None or c in seen:
                pr.append((0, 0.0, 0.0))
            else:
                pr.append(priors[c])
                seen.add(c)
        M = pe.shape[0]
        exo = np.array([self.names.index(n) for n in exo_names], dtype=np.int32)
        S = 0 if scenarios is None else scenarios.shape[0]
        sc = None if scenarios is None else np.ascontiguousarray(scenarios, dtype=np.float64)
        obs = self.make_obs(observations)
        lp = np.empty(max(S, 1) * M)
        self._ck(self.L.orc_log_posterior_batch(self.h, pe.shape[1], _dp(bc), _dp(bi), _dp(pe), C.c_int64(M), exo.size, _dp(exo), _dp(sc),
                                                C.c_int64(S), self.make_priors(pr), obs, C.c_int64(len(observations)),
                                                1 if normalize else 0, _dp(lp), n_threads))
        return lp


def max_threads() -> int:
    return lib().orc_max_threads()


def ghg_forcings(params, co2, ch4, n2o):
    p = np.ascontiguousarray(params, dtype=np.float64)
    out = np.empty(3)
    lib().orc_ghg_forcings(_dp(p), co2, ch4, n2o, _dp(out))
    return out
