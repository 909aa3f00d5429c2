"""Mirror of ``rscm._lib.magicc`` (python/rscm/_lib/magicc.pyi) for the kinds with device code.

Every builder accepts the reference's parameter dictionary (serde ``default`` semantics: missing keys take the
defaults of ``crates/rscm-magicc/src/parameters/*.rs``) and flattens it into the C ABI's parameter block.
"""

from . import _ffi
from ._builders import ComponentBuilder
from .core import Component

__all__ = ["GhgForcingBuilder", "OzoneForcingBuilder", "AerosolDirectBuilder", "AerosolIndirectBuilder", "ClimateUDEBBuilder",
           "CO2BudgetBuilder", "TerrestrialCarbonBuilder", "CH4ChemistryBuilder", "N2OChemistryBuilder", "OceanCarbonBuilder"]


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


class OzoneForcingBuilder(ComponentBuilder):
    """OzoneForcingParameters — crates/rscm-magicc/src/parameters/ozone_forcing.rs."""

    KIND = _ffi.OZONE_FORCING
    TYPE_NAME = "OzoneForcing"
    FIELDS = (
        ("eesc_reference", 1420.0), ("strat_o3_scale", -0.0043), ("strat_cl_exponent", 1.7), ("trop_radeff", 0.032),
        ("trop_oz_ch4", 5.7), ("trop_oz_nox", 0.168), ("trop_oz_co", 0.00396), ("trop_oz_voc", 0.01008), ("ch4_pi", 700.0),
        ("nox_pi", 0.0), ("co_pi", 0.0), ("nmvoc_pi", 0.0), ("temp_feedback_scale", -0.037),
    )


class _ArrayFieldsBuilder(ComponentBuilder):
    """Builders whose reference parameter struct has `[f64; 4]` fields: they are flattened to name_0..name_3."""

    ARRAYS: dict = {}

    @classmethod
    def from_parameters(cls, parameters: dict):
        flat = {}
        for k, v in parameters.items():
            if k in cls.ARRAYS:
                if len(v) != 4:
                    raise ValueError(f"{cls.TYPE_NAME}: {k} needs 4 values")
                for i, x in enumerate(v):
                    flat[f"{k}_{i}"] = x
            else:
                flat[k] = v
        return super().from_parameters(flat)


class AerosolDirectBuilder(_ArrayFieldsBuilder):
    """AerosolDirectParameters — crates/rscm-magicc/src/parameters/aerosol.rs."""

    KIND = _ffi.AEROSOL_DIRECT
    TYPE_NAME = "AerosolDirect"
    ARRAYS = {"sox_regional": None, "bc_regional": None, "oc_regional": None, "nitrate_regional": None}
    FIELDS = (
        ("sox_coefficient", -0.0035), ("bc_coefficient", 0.0077), ("oc_coefficient", -0.002), ("nitrate_coefficient", -0.001),
        *[(f"sox_regional_{i}", v) for i, v in enumerate([0.15, 0.55, 0.10, 0.20])],
        *[(f"bc_regional_{i}", v) for i, v in enumerate([0.15, 0.50, 0.15, 0.20])],
        *[(f"oc_regional_{i}", v) for i, v in enumerate([0.15, 0.45, 0.15, 0.25])],
        *[(f"nitrate_regional_{i}", v) for i, v in enumerate([0.15, 0.50, 0.15, 0.20])],
        ("sox_pi", 1.0), ("bc_pi", 2.5), ("oc_pi", 10.0), ("nox_pi", 10.0),
        ("harmonize", 0.0), ("harmonize_year", 2019.0), ("harmonize_target", -0.22),
    )


class AerosolIndirectBuilder(ComponentBuilder):
    """AerosolIndirectParameters — crates/rscm-magicc/src/parameters/aerosol.rs."""

    KIND = _ffi.AEROSOL_INDIRECT
    TYPE_NAME = "AerosolIndirect"
    FIELDS = (
        ("cloud_albedo_coefficient", -1.0), ("reference_burden", 50.0), ("sox_weight", 1.0), ("oc_weight", 0.3),
        ("sox_pi", 1.0), ("oc_pi", 10.0), ("harmonize", 0.0), ("harmonize_year", 2019.0), ("harmonize_target", -0.89),
    )


class CO2BudgetBuilder(ComponentBuilder):
    """CO2BudgetParameters — crates/rscm-magicc/src/parameters/co2_budget.rs:35-42."""

    KIND = _ffi.CO2_BUDGET
    TYPE_NAME = "CO2Budget"
    FIELDS = (("gtc_per_ppm", 2.123), ("co2_pi", 278.0))


class TerrestrialCarbonBuilder(ComponentBuilder):
    """TerrestrialCarbonParameters — crates/rscm-magicc/src/parameters/terrestrial_carbon.rs:153-190."""

    KIND = _ffi.TERRESTRIAL_CARBON
    TYPE_NAME = "TerrestrialCarbon"
    FIELDS = (
        ("npp_pi", 66.27), ("co2_pi", 278.0), ("beta", 0.6486), ("npp_temp_sensitivity", 0.0107), ("resp_temp_sensitivity", 0.0685),
        ("detritus_temp_sensitivity", 0.1358), ("soil_temp_sensitivity", 0.1541), ("humus_temp_sensitivity", 0.05),
        ("plant_pool_pi", 884.86), ("detritus_pool_pi", 92.77), ("soil_pool_pi", 1681.53), ("humus_pool_pi", 836.0),
        ("respiration_pi", 12.26), ("frac_npp_to_plant", 0.4483), ("frac_npp_to_detritus", 0.3998), ("frac_plant_to_detritus", 0.9989),
        ("frac_detritus_to_soil", 0.3), ("frac_soil_to_humus", 0.1), ("enable_fertilization", 1.0), ("enable_temp_feedback", 1.0),
    )


class CH4ChemistryBuilder(ComponentBuilder):
    """CH4ChemistryParameters — crates/rscm-magicc/src/parameters/ch4_chemistry.rs:131-152."""

    KIND = _ffi.CH4_CHEMISTRY
    TYPE_NAME = "CH4Chemistry"
    FIELDS = (
        ("ch4_pi", 722.0), ("natural_emissions", 209.0), ("tau_oh", 9.3), ("tau_soil", 150.0), ("tau_strat", 120.0), ("tau_trop_cl", 200.0),
        ("ch4_self_feedback", -0.32), ("oh_sensitivity_scale", 0.72), ("oh_nox_sensitivity", 0.0042), ("oh_co_sensitivity", -0.000105),
        ("oh_nmvoc_sensitivity", -0.000315), ("temp_sensitivity", 0.0316), ("include_temp_feedback", 1.0),
        ("include_emissions_feedback", 1.0), ("ppb_to_tg", 2.75), ("nox_reference", 0.0), ("co_reference", 0.0), ("nmvoc_reference", 0.0),
    )


class N2OChemistryBuilder(ComponentBuilder):
    """N2OChemistryParameters — crates/rscm-magicc/src/parameters/n2o_chemistry.rs:78-88."""

    KIND = _ffi.N2O_CHEMISTRY
    TYPE_NAME = "N2OChemistry"
    FIELDS = (("n2o_pi", 270.0), ("natural_emissions", 11.0), ("tau_n2o", 139.275), ("lifetime_feedback", -0.04), ("strat_delay", 1),
              ("ppb_to_tg", 4.79))


class OceanCarbonBuilder:
    """OceanCarbonParameters — crates/rscm-magicc/src/parameters/ocean_carbon.rs (presets gfdl_3d / bern_2d / hilda, :150-260).

    ``from_parameters({"model": "3D-GFDL" | "2D-BERN" | "HILDA", ...overrides})``; ``irf_early`` / ``irf_late`` may be given as the
    reference's tagged dicts ``{"type": "Polynomial", "coefficients": [...]}`` / ``{"type": "ExponentialSum", "coefficients": [...],
    "timescales": [...]}`` (at most 8 terms).  The flattened block layout is documented in include/rscm_b200.h."""

    TYPE_NAME = "OceanCarbon"
    _OSPP_OFF = [1.5568, 7.4706, 1.2748, 2.4491, 1.5468]
    _OSPP_COEF = [-0.013993, -0.20207, -0.12015, -0.12639, -0.15326]
    _COMMON = dict(co2_pi=278.0, pco2_pi=278.0, gas_exchange_scale=1.833492, temp_sensitivity=0.03717879, irf_scale=0.9492864,
                   steps_per_year=12, max_history_months=6000, enable_temp_feedback=True)
    PRESETS = {
        "3D-GFDL": dict(gas_exchange_tau=7.66, irf_switch_time=1.0, mixed_layer_depth=50.9, ocean_surface_area=3.55e14, sst_pi=17.7,
                        irf_early={"type": "Polynomial", "coefficients": [1.0, -2.2617, 14.002, -48.770, 82.986, -67.527, 21.037]},
                        irf_late={"type": "ExponentialSum", "coefficients": [0.01481, 0.019439, 0.038344, 0.066485, 0.24966, 0.70367],
                                  "timescales": [1.0e10, 347.55, 65.359, 15.281, 2.3488, 0.70177]}),
        "2D-BERN": dict(gas_exchange_tau=7.46, irf_switch_time=9.9, mixed_layer_depth=50.0, ocean_surface_area=3.5375e14, sst_pi=18.2997,
                        irf_early={"type": "ExponentialSum", "coefficients": [0.058648, 0.07515, 0.079338, 0.41413, 0.24845, 0.12429],
                                   "timescales": [1.0e10, 9.6218, 9.2364, 0.7603, 0.16294, 0.0032825]},
                        irf_late={"type": "ExponentialSum", "coefficients": [0.01369, 0.012456, 0.026933, 0.026994, 0.036608, 0.06738],
                                  "timescales": [1.0e10, 331.54, 107.57, 38.946, 11.677, 10.515]}),
        "HILDA": dict(gas_exchange_tau=9.06, irf_switch_time=2.0, mixed_layer_depth=75.0, ocean_surface_area=3.62e14, sst_pi=18.1716,
                      irf_early={"type": "ExponentialSum", "coefficients": [0.12935, 0.24093, 0.24071, 0.17003, 0.21898],
                                 "timescales": [1.0e10, 4.9792, 0.96083, 0.26936, 0.034569]},
                      irf_late={"type": "ExponentialSum", "coefficients": [0.022936, 0.035549, 0.037820, 0.089318, 0.13963, 0.24278],
                                "timescales": [1.0e10, 232.30, 68.736, 18.601, 5.2528, 1.2679]}),
    }
    _MODEL_ID = {"3D-GFDL": 0.0, "2D-BERN": 1.0, "HILDA": 2.0}

    def __init__(self, parameters: dict):
        self._parameters = dict(parameters)

    @classmethod
    def from_parameters(cls, parameters: dict):
        model = parameters.get("model", "3D-GFDL")
        if model not in cls.PRESETS:
            raise ValueError(f"OceanCarbon: unknown model {model!r}")
        known = set(cls._COMMON) | set(cls.PRESETS[model]) | {"model", "delta_ospp_offsets", "delta_ospp_coefficients"}
        unknown = set(parameters) - known
        if unknown:
            raise ValueError(f"OceanCarbon: unknown parameter(s) {sorted(unknown)}")
        return cls(parameters)

    @staticmethod
    def _flatten_irf(form: dict) -> list:
        coef = list(form["coefficients"])
        tau = list(form.get("timescales", []))
        kind = {"Polynomial": 0.0, "ExponentialSum": 1.0}[form["type"]]
        if len(coef) > 8 or (kind == 1.0 and len(tau) != len(coef)):
            raise ValueError("OceanCarbon: IRF forms take at most 8 terms with matching timescales")
        return [kind, float(len(coef))] + coef + [0.0] * (8 - len(coef)) + (tau + [1.0] * (8 - len(tau)))

    def build(self) -> Component:
        model = self._parameters.get("model", "3D-GFDL")
        p = {**self._COMMON, **self.PRESETS[model], "delta_ospp_offsets": self._OSPP_OFF, "delta_ospp_coefficients": self._OSPP_COEF}
        p.update({k: v for k, v in self._parameters.items() if k != "model"})
        vals = [self._MODEL_ID[model], p["co2_pi"], p["pco2_pi"], p["gas_exchange_scale"], p["gas_exchange_tau"], p["temp_sensitivity"],
                p["irf_scale"], p["mixed_layer_depth"], p["ocean_surface_area"], p["sst_pi"], p["steps_per_year"], p["max_history_months"],
                p["irf_switch_time"]]
        vals += self._flatten_irf(p["irf_early"]) + self._flatten_irf(p["irf_late"])
        vals += list(p["delta_ospp_offsets"]) + list(p["delta_ospp_coefficients"]) + [1.0 if p["enable_temp_feedback"] else 0.0]
        names = (["model", "co2_pi", "pco2_pi", "gas_exchange_scale", "gas_exchange_tau", "temp_sensitivity", "irf_scale", "mixed_layer_depth",
                  "ocean_surface_area", "sst_pi", "steps_per_year", "max_history_months", "irf_switch_time"]
                 + [f"irf_{w}_{f}" for w in ("early", "late") for f in (["kind", "n"] + [f"c{i}" for i in range(8)] + [f"t{i}" for i in range(8)])]
                 + [f"delta_ospp_offsets_{i}" for i in range(5)] + [f"delta_ospp_coefficients_{i}" for i in range(5)] + ["enable_temp_feedback"])
        # the emitter's name order interleaves early/late per form; rebuild it exactly as graph.cpp lists it
        names = (names[:13]
                 + ["irf_early_kind", "irf_early_n"] + [f"irf_early_c{i}" for i in range(8)] + [f"irf_early_t{i}" for i in range(8)]
                 + ["irf_late_kind", "irf_late_n"] + [f"irf_late_c{i}" for i in range(8)] + [f"irf_late_t{i}" for i in range(8)]
                 + names[-11:])
        assert len(vals) == 60 and len(names) == 60
        return Component(_ffi.OCEAN_CARBON, self.TYPE_NAME, names, vals)


class ClimateUDEBBuilder(_ArrayFieldsBuilder):
    """ClimateUDEBParameters — crates/rscm-magicc/src/parameters/climate_udeb.rs:235-300 (defaults)."""

    KIND = _ffi.CLIMATE_UDEB
    TYPE_NAME = "ClimateUDEB"
    ARRAYS = {"rf_regions_co2": None}
    FIELDS = (
        ("n_layers", 50), ("mixed_layer_depth", 60.0), ("layer_thickness", 100.0),
        ("kappa", 0.75), ("kappa_min", 0.1), ("kappa_dkdt", -0.191),
        ("w_initial", 3.5), ("w_variable_fraction", 0.7), ("w_threshold_temp_nh", 8.0), ("w_threshold_temp_sh", 8.0),
        ("ecs", 3.0), ("rf_2xco2", 3.71), ("rlo", 1.317),
        ("feedback_q_sensitivity", 7.84e-9), ("feedback_cumt_sensitivity", 0.08), ("feedback_cumt_period", 300.0),
        ("k_lo", 1.44), ("k_ns", 0.31), ("amplify_ocean_to_land", 1.02),
        ("nh_land_fraction", 0.42), ("sh_land_fraction", 0.21), ("depth_dependent_area", 1.0),
        ("temp_adjust_alpha", 1.04), ("temp_adjust_gamma", -0.002), ("polar_sinking_ratio", 0.2),
        ("land_heat_capacity_enabled", 1.0), ("k_lg", 0.1), ("land_hc_eff_thickness", 300.0),
        ("rf_regions_co2_0", 1.4089), ("rf_regions_co2_1", 1.37045), ("rf_regions_co2_2", 1.43333), ("rf_regions_co2_3", 1.33257),
        ("efficacy_apply", 0), ("prescribed_efficacy_co2", 1.0),
        ("ocean_temp_profile", 2), ("steps_per_year", 12), ("max_temperature", 25.0),
    )

    def _value(self, name, default):
        v = self._parameters.get(name, default)
        if name == "ocean_temp_profile" and isinstance(v, str):
            return {"analytical": 1.0, "cmip5": 2.0, "1": 1.0, "2": 2.0}[v.lower()]
        return float(v)


class HalocarbonChemistryBuilder:
    """HalocarbonParameters — crates/rscm-magicc/src/parameters/halocarbon.rs:104-200, default species tables :203-262.

    ``from_parameters({...})`` takes the reference's scalar fields and, optionally, ``fgases`` / ``montreal_gases`` as lists of the
    reference's species dicts (``name, lifetime, radiative_efficiency, concentration_pi, molecular_weight, n_cl, n_br,
    fractional_release``).  The device program is built for the reference's default species *list* (the variable names are part of the
    compiled graph); the species' *numbers* may be overridden, a different list is refused."""

    TYPE_NAME = "HalocarbonChemistry"
    SCALARS = (("br_multiplier", 60.0), ("cfc11_release_normalisation", 0.47), ("eesc_delay", 3.0), ("air_molar_mass", 28.97),
               ("atmospheric_mass_tg", 5.133e9), ("mixing_box_fraction", 0.949))
    SPECIES_FIELDS = ("lifetime", "radiative_efficiency", "concentration_pi", "molecular_weight", "n_cl", "n_br", "fractional_release")
    # name, lifetime (yr), radiative efficiency (W/m^2/ppb), pre-industrial (ppt), molecular weight, n_cl, n_br, fractional release
    FGASES = (
        ("CF4", 50000.0, 0.09, 0.0, 88.0, 0, 0, 0.0), ("C2F6", 10000.0, 0.25, 0.0, 138.0, 0, 0, 0.0), ("C3F8", 2600.0, 0.28, 0.0, 188.0, 0, 0, 0.0),
        ("C4F10", 2600.0, 0.36, 0.0, 238.0, 0, 0, 0.0), ("C5F12", 4100.0, 0.41, 0.0, 288.0, 0, 0, 0.0), ("C6F14", 3100.0, 0.44, 0.0, 338.0, 0, 0, 0.0),
        ("C7F16", 3000.0, 0.50, 0.0, 388.0, 0, 0, 0.0), ("C8F18", 3000.0, 0.55, 0.0, 438.0, 0, 0, 0.0), ("c-C4F8", 3200.0, 0.32, 0.0, 200.0, 0, 0, 0.0),
        ("HFC-23", 228.0, 0.18, 0.0, 70.0, 0, 0, 0.0), ("HFC-32", 5.4, 0.11, 0.0, 52.0, 0, 0, 0.0), ("HFC-43-10mee", 17.0, 0.359, 0.0, 252.0, 0, 0, 0.0),
        ("HFC-125", 31.0, 0.23, 0.0, 120.0, 0, 0, 0.0), ("HFC-134a", 14.0, 0.16, 0.0, 102.0, 0, 0, 0.0), ("HFC-143a", 51.0, 0.16, 0.0, 84.0, 0, 0, 0.0),
        ("HFC-152a", 1.6, 0.10, 0.0, 66.0, 0, 0, 0.0), ("HFC-227ea", 36.0, 0.26, 0.0, 170.0, 0, 0, 0.0), ("HFC-236fa", 213.0, 0.24, 0.0, 152.0, 0, 0, 0.0),
        ("HFC-245fa", 7.9, 0.24, 0.0, 134.0, 0, 0, 0.0), ("HFC-365mfc", 8.9, 0.22, 0.0, 148.0, 0, 0, 0.0), ("NF3", 569.0, 0.20, 0.0, 71.0, 0, 0, 0.0),
        ("SF6", 850.0, 0.57, 0.0, 146.0, 0, 0, 0.0), ("SO2F2", 36.0, 0.20, 0.0, 102.0, 0, 0, 0.0),
    )
    MONTREAL_GASES = (
        ("CFC-11", 52.0, 0.295, 0.0, 137.4, 3, 0, 0.47), ("CFC-12", 102.0, 0.364, 0.0, 120.9, 2, 0, 0.23), ("CFC-113", 93.0, 0.30, 0.0, 187.4, 3, 0, 0.29),
        ("CFC-114", 189.0, 0.31, 0.0, 170.9, 2, 0, 0.12), ("CFC-115", 540.0, 0.20, 0.0, 154.5, 1, 0, 0.04), ("HCFC-22", 11.9, 0.21, 0.0, 86.5, 1, 0, 0.13),
        ("HCFC-141b", 9.4, 0.16, 0.0, 116.9, 2, 0, 0.34), ("HCFC-142b", 18.0, 0.19, 0.0, 100.5, 1, 0, 0.17), ("CH3CCl3", 5.0, 0.07, 0.0, 133.4, 3, 0, 0.67),
        ("CCl4", 32.0, 0.174, 0.0, 153.8, 4, 0, 0.56), ("CH3Cl", 0.9, 0.004, 500.0, 50.5, 1, 0, 0.44), ("CH2Cl2", 0.5, 0.028, 0.0, 84.9, 2, 0, 0.0),
        ("CHCl3", 0.5, 0.07, 0.0, 119.4, 3, 0, 0.0), ("CH3Br", 0.8, 0.004, 5.0, 94.9, 0, 1, 0.60), ("Halon-1211", 16.0, 0.29, 0.0, 165.4, 1, 1, 0.62),
        ("Halon-1301", 72.0, 0.30, 0.0, 148.9, 0, 1, 0.28), ("Halon-2402", 28.0, 0.31, 0.0, 259.8, 0, 2, 0.65), ("Halon-1202", 2.5, 0.27, 0.0, 209.8, 0, 2, 0.62),
    )

    def __init__(self, parameters: dict):
        self._parameters = dict(parameters)

    @classmethod
    def species_names(cls):
        return [s[0] for s in cls.FGASES + cls.MONTREAL_GASES]

    @classmethod
    def from_parameters(cls, parameters: dict):
        known = {n for n, _ in cls.SCALARS} | {"fgases", "montreal_gases"}
        unknown = set(parameters) - known
        if unknown:
            raise ValueError(f"HalocarbonChemistry: unknown parameter(s) {sorted(unknown)}")
        return cls(parameters)

    def _species(self, key, defaults):
        given = self._parameters.get(key)
        if given is None:
            return [list(s[1:]) for s in defaults]
        if [g["name"] for g in given] != [s[0] for s in defaults]:
            raise ValueError(f"HalocarbonChemistry: the device program is built for the reference's default {key} list "
                             f"({', '.join(s[0] for s in defaults)}); only the species' numbers may change")
        return [[float(g[f]) for f in self.SPECIES_FIELDS] for g in given]

    def build(self) -> Component:
        names = [n for n, _ in self.SCALARS]
        vals = [float(self._parameters.get(n, d)) for n, d in self.SCALARS]
        rows = self._species("fgases", self.FGASES) + self._species("montreal_gases", self.MONTREAL_GASES)
        for sname, row in zip(self.species_names(), rows):
            names += [f"{f}[{sname}]" for f in self.SPECIES_FIELDS]
            vals += [float(v) for v in row]
        assert len(vals) == 293
        return Component(_ffi.HALOCARBON_CHEMISTRY, self.TYPE_NAME, names, vals)
