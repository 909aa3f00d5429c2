"""Mirror of ``rscm._lib.components`` (python/rscm/_lib/components.pyi) for the kinds on the hot path."""

from . import _ffi
from ._builders import ComponentBuilder

__all__ = ["CarbonCycleBuilder", "CO2ERFBuilder", "FourBoxOceanHeatUptakeBuilder", "OceanSurfacePartialPressureBuilder", "GTC_PER_PPM"]

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


class FourBoxOceanHeatUptakeBuilder(ComponentBuilder):
    """FourBoxOceanHeatUptakeParameters — crates/rscm-components/src/components/four_box_ocean_heat_uptake.rs
    (ratios must average to 1.0 +- 0.01 with equal weights, as `from_parameters` asserts)."""

    KIND = _ffi.FOUR_BOX_OHU
    TYPE_NAME = "FourBoxOceanHeatUptake"
    FIELDS = (("northern_ocean_ratio", 1.2), ("northern_land_ratio", 0.6), ("southern_ocean_ratio", 1.6), ("southern_land_ratio", 0.6))

    def build(self):
        c = super().build()
        avg = sum(c.params) / 4.0
        if abs(avg - 1.0) >= 0.01:
            raise ValueError(f"Regional ratios must average to 1.0 with equal weights (got {avg})")
        return c


class OceanSurfacePartialPressureBuilder(ComponentBuilder):
    """OceanSurfacePartialPressureParameters — .../ocean_carbon_cycle/ocean_surface_partial_pressure.rs (all required)."""

    KIND = _ffi.OCEAN_SURFACE_PP
    TYPE_NAME = "OceanSurfacePartialPressure"
    FIELDS = (
        ("ospp_preindustrial", None), ("sensitivity_ospp_to_temperature", None), ("sea_surface_temperature_preindustrial", None),
        *[(f"delta_ospp_offsets_{i}", None) for i in range(5)], *[(f"delta_ospp_coefficients_{i}", None) for i in range(5)],
    )

    @classmethod
    def from_parameters(cls, parameters: dict):
        flat = {}
        for k, v in parameters.items():
            if k in ("delta_ospp_offsets", "delta_ospp_coefficients"):
                if len(v) != 5:
                    raise ValueError(f"{k} needs 5 values")
                flat.update({f"{k}_{i}": x for i, x in enumerate(v)})
            else:
                flat[k] = v
        return super().from_parameters(flat)
