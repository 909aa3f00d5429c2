"""Mirror of ``rscm._lib.components`` (python/rscm/_lib/components.pyi) for the kinds on the hot path."""

from . import _ffi
from ._builders import ComponentBuilder

__all__ = ["CarbonCycleBuilder", "CO2ERFBuilder", "GTC_PER_PPM"]

GTC_PER_PPM = 2.13  # crates/rscm-components/src/constants.rs:37


class CarbonCycleBuilder(ComponentBuilder):
    """CarbonCycleParameters — crates/rscm-components/src/components/carbon_cycle.rs:24-34;
    ``step_size`` is SolverOptions (:37-40, default 0.1 at :83)."""

    KIND = _ffi.CARBON_CYCLE
    TYPE_NAME = "CarbonCycle"
    FIELDS = (("tau", None), ("conc_pi", None), ("alpha_temperature", None), ("step_size", 0.1))

    def with_solver_options(self, step_size: float) -> "CarbonCycleBuilder":
        self._parameters["step_size"] = float(step_size)
        return self


class CO2ERFBuilder(ComponentBuilder):
    """CO2ERFParameters — crates/rscm-components/src/components/co2_erf.rs:18-25."""

    KIND = _ffi.CO2_ERF
    TYPE_NAME = "CO2ERF"
    FIELDS = (("erf_2xco2", None), ("conc_pi", None))
