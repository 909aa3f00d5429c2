"""Mirror of ``rscm._lib.magicc`` (python/rscm/_lib/magicc.pyi) for the kinds with device code."""

from . import _ffi
from ._builders import ComponentBuilder

__all__ = ["GhgForcingBuilder"]


class GhgForcingBuilder(ComponentBuilder):
    """GhgForcingParameters — crates/rscm-magicc/src/parameters/ghg_forcing.rs (serde defaults)."""

    KIND = _ffi.GHG_FORCING
    TYPE_NAME = "GhgForcing"
    FIELDS = (
        ("method", "Olbl"),
        ("co2_pi", 278.0),
        ("ch4_pi", 722.0),
        ("n2o_pi", 270.0),
        ("delq2xco2", 3.71),
        ("ch4_radeff", 0.036),
        ("n2o_radeff", 0.12),
        ("olbl_co2_a1", -2.4785e-7),
        ("olbl_co2_b1", 7.5906e-4),
        ("olbl_co2_c1", -2.1492e-3),
        ("olbl_co2_d1", 5.2),
        ("olbl_ch4_a3", -8.9603e-5),
        ("olbl_ch4_b3", -1.2462e-4),
        ("olbl_ch4_d3", 0.045),
        ("olbl_n2o_a2", -3.4197e-4),
        ("olbl_n2o_b2", 2.5455e-4),
        ("olbl_n2o_c2", -2.4357e-4),
        ("olbl_n2o_d2", 0.14),
        ("adjust_co2", 1.05),
        ("adjust_ch4", 0.86),
        ("adjust_n2o", 1.0),
    )

    def _value(self, name, default):
        v = self._parameters.get(name, default)
        if name == "method":
            if isinstance(v, str):
                try:
                    return {"ipcctar": 0.0, "olbl": 1.0}[v.lower()]
                except KeyError:
                    raise ValueError(f"GhgForcing: unknown method {v!r}") from None
            return float(v)
        return float(v)
