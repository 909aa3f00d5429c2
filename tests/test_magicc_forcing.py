"""MAGICC forcing box components (OzoneForcing, AerosolDirect, AerosolIndirect, with GhgForcing):
oracle known answers from the reference's unit tests (CPU) and GPU parity on a full forcing graph with
write-side FourBox -> Scalar aggregation and an 8-contributor Sum aggregate
(tests/regression/test_ghg_forcing.py:395-464 in the reference)."""

import math

import numpy as np
import pytest

from rscm_b200 import synthetic as syn
from rscm_b200.core import GridType, ModelBuilder, VariableSchema
from rscm_b200.magicc import AerosolDirectBuilder, AerosolIndirectBuilder, GhgForcingBuilder, OzoneForcingBuilder
from rscm_b200.two_layer import TwoLayerBuilder

from .helpers import oracle_bindings, oracle_from_builder, rel_err

ERF_PARTS = [
    "Effective Radiative Forcing|CO2", "Effective Radiative Forcing|CH4", "Effective Radiative Forcing|N2O",
    "Effective Radiative Forcing|O3|Stratospheric", "Effective Radiative Forcing|O3|Tropospheric",
    "Effective Radiative Forcing|O3|Temperature Feedback", "Effective Radiative Forcing|Aerosol|Direct",
    "Effective Radiative Forcing|Aerosol|Indirect",
]


def forcing_builder(aerosol_grid=GridType.Scalar, grid_weights=None):
    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "ppm")
    for n in ("NOx", "CO", "NMVOC", "SOx", "BC", "OC"):
        schema.add_variable(f"Emissions|{n}", "Mt/yr")
    schema.add_variable("EESC", "ppt")
    for n in ERF_PARTS:
        schema.add_variable(n, "W/m^2", aerosol_grid if n.endswith("Aerosol|Direct") else GridType.Scalar)
    schema.add_variable("Surface Temperature", "K")
    schema.add_variable("Deep Ocean Temperature", "K")
    contributors = [p for p in ERF_PARTS if aerosol_grid == GridType.Scalar or not p.endswith("Aerosol|Direct")]
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", contributors)
    b = (ModelBuilder().with_time_axis(syn.time_axis()).with_schema(schema)
         .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
         .with_rust_component(OzoneForcingBuilder.from_parameters({}).build())
         .with_rust_component(AerosolDirectBuilder.from_parameters({"sox_regional": [0.2, 0.5, 0.1, 0.2]}).build())
         .with_rust_component(AerosolIndirectBuilder.from_parameters({}).build())
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}))
    if grid_weights is not None:
        b.with_grid_weights(GridType.FourBox, grid_weights)
    return b


def forcing_scenarios(n=2):
    years = syn.time_axis().values()
    ramp = np.maximum(0.0, years - 1850.0) / 250.0
    out = []
    for s in range(n):
        f = 1.0 + 0.2 * s
        out.append({
            "Atmospheric Concentration|CO2": 278.0 * np.exp(0.0045 * f * np.maximum(0.0, years - 1850.0)),
            "Atmospheric Concentration|CH4": 722.0 + 1200.0 * ramp * f,
            "Atmospheric Concentration|N2O": 270.0 + 60.0 * ramp,
            "EESC": 1000.0 + 1500.0 * np.exp(-((years - 2000.0) / 40.0) ** 2),
            "Emissions|NOx": 10.0 + 30.0 * ramp, "Emissions|CO": 300.0 * ramp, "Emissions|NMVOC": 100.0 * ramp,
            "Emissions|SOx": 1.0 + 60.0 * ramp * np.exp(-np.maximum(0.0, years - 1990.0) / 60.0) * f,
            "Emissions|BC": 2.5 + 5.0 * ramp, "Emissions|OC": 10.0 + 20.0 * ramp,
        })
    return out


# ---- oracle known answers: crates/rscm-magicc/src/forcing/{ozone,aerosol_direct,aerosol_indirect}.rs unit tests ----
def _one_step(builder, scenario):
    return {k: v for k, v in oracle_from_builder(builder, scenario).run().items()}


def _single(component, inputs, T=3):
    axis = syn.time_axis(2000, 2000 + T - 1)
    b = ModelBuilder().with_time_axis(axis).with_rust_component(component)
    sc = {k: np.full(T, v) for k, v in inputs.items()}
    return oracle_from_builder(b, sc).run()


def test_ozone_known_answers():
    comp = OzoneForcingBuilder.from_parameters({}).build()
    base = {"EESC": 1420.0, "Atmospheric Concentration|CH4": 700.0, "Emissions|NOx": 0.0, "Emissions|CO": 0.0,
            "Emissions|NMVOC": 0.0, "Surface Temperature": 0.0}
    r = _single(comp, base)
    assert abs(r["Effective Radiative Forcing|O3|Stratospheric"][1]) < 1e-10  # zero at the reference EESC
    assert abs(r["Effective Radiative Forcing|O3|Tropospheric"][1]) < 1e-10   # zero at pre-industrial
    assert _single(comp, {**base, "EESC": 920.0})["Effective Radiative Forcing|O3|Stratospheric"][1] == 0.0
    lo = _single(comp, {**base, "EESC": 1620.0})["Effective Radiative Forcing|O3|Stratospheric"][1]
    hi = _single(comp, {**base, "EESC": 1820.0})["Effective Radiative Forcing|O3|Stratospheric"][1]
    assert hi < lo < 0.0 and abs(hi / lo - 2.0 ** 1.7) < 0.01          # power-law scaling test
    peak = _single(comp, {**base, "EESC": 2000.0})["Effective Radiative Forcing|O3|Stratospheric"][1]
    assert -0.15 < peak < 0.0
    fb = _single(comp, {**base, "Surface Temperature": 2.0})["Effective Radiative Forcing|O3|Temperature Feedback"][1]
    assert fb == pytest.approx(-0.074)


def test_aerosol_known_answers():
    d = AerosolDirectBuilder.from_parameters({}).build()
    pi = {"Emissions|SOx": 1.0, "Emissions|BC": 2.5, "Emissions|OC": 10.0, "Emissions|NOx": 10.0}
    r = _single(d, pi)["Effective Radiative Forcing|Aerosol|Direct"]
    assert np.all(r[1] == 0.0) and np.isnan(r[0]).all()                    # zero forcing -> uniform zeros
    r = _single(d, {**pi, "Emissions|SOx": 51.0})["Effective Radiative Forcing|Aerosol|Direct"][1]
    assert r.sum() == pytest.approx(-0.0035 * 50.0) and np.allclose(r / r.sum(), [0.15, 0.55, 0.10, 0.20])
    i = AerosolIndirectBuilder.from_parameters({}).build()
    assert _single(i, {"Emissions|SOx": 1.0, "Emissions|OC": 10.0})["Effective Radiative Forcing|Aerosol|Indirect"][1] == 0.0
    assert _single(i, {"Emissions|SOx": 0.5, "Emissions|OC": 5.0})["Effective Radiative Forcing|Aerosol|Indirect"][1] == 0.0
    v = _single(i, {"Emissions|SOx": 51.0, "Emissions|OC": 10.0})["Effective Radiative Forcing|Aerosol|Indirect"][1]
    assert v == pytest.approx(-math.log(2.0))


def test_write_side_aggregation_in_oracle():
    b = forcing_builder()
    m = oracle_from_builder(b, forcing_scenarios(1)[0])
    r = m.run()
    assert m.grids[m.names.index("Effective Radiative Forcing|Aerosol|Direct")] == 0  # stored as Scalar (schema wins)
    total = sum(r[p][1:] for p in ERF_PARTS)
    np.testing.assert_allclose(r["Effective Radiative Forcing"][1:], total, rtol=1e-13)


# ---- GPU parity -----------------------------------------------------------------------------------------
BINDS = {**syn.TWO_LAYER_BINDINGS, "adjust_ch4": "GhgForcing.adjust_ch4", "sox_coefficient": "AerosolDirect.sox_coefficient",
         "cloud": "AerosolIndirect.cloud_albedo_coefficient", "trop_radeff": "OzoneForcing.trop_radeff"}


def _params(M):
    rng = np.random.default_rng(17)
    tl = syn.uniform_params(syn.TWO_LAYER_RANGES, M, 3)
    tl[:, 1] *= 0.2  # keep a*Ts^2 from running away at this forcing level (overflow steps are not comparable)
    return np.column_stack([tl, rng.uniform(0.7, 1.0, M), rng.uniform(-0.005, -0.002, M),
                            rng.uniform(-1.5, -0.5, M), rng.uniform(0.02, 0.05, M)])


@pytest.mark.gpu
@pytest.mark.parametrize("aerosol_grid,weights", [(GridType.Scalar, None), (GridType.Scalar, [0.3, 0.2, 0.4, 0.1]), (GridType.FourBox, None)])
def test_full_forcing_graph_parity(aerosol_grid, weights, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = forcing_builder(aerosol_grid, weights)
    ens = b.build_ensemble().bind_parameters(BINDS)
    scen = forcing_scenarios(2)
    sc = ens.pack_scenarios(scen)
    p = _params(150)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, BINDS), p, ens.exogenous_names, sc, names), names)
    assert ens.program_is_jit()
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    direct = got["Effective Radiative Forcing|Aerosol|Direct"]
    assert direct.ndim == (3 if aerosol_grid == GridType.FourBox else 2)
    assert np.isfinite(got["Surface Temperature"]).all()


# ---- Mean / Weighted aggregates and the Hemispheric grid (schema.rs:760-806, aggregating.rs:611-618, transformations.rs) ----
def mean_and_hemispheric_builder(hemi_weights=None, four_box_weights=None):
    """GhgForcing -> Mean / Weighted aggregates (one contributor is an exogenous series with NaN years, so the Mean divides by
    the non-NaN count); AerosolDirect (FourBox) stored on the Hemispheric grid (write-side FourBox -> Hemispheric) ->
    Hemispheric Mean aggregate with an exogenous Hemispheric contributor -> read Hemispheric -> Scalar by TwoLayer."""
    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "ppm")
        schema.add_variable(f"Effective Radiative Forcing|{n}", "W/m^2")
    for n in ("SOx", "BC", "OC", "NOx"):
        schema.add_variable(f"Emissions|{n}", "Mt/yr")
    schema.add_variable("Effective Radiative Forcing|Aerosol|Direct", "W/m^2", GridType.Hemispheric)
    schema.add_variable("Effective Radiative Forcing|Volcanic", "W/m^2", GridType.Hemispheric)
    schema.add_variable("Effective Radiative Forcing|Other", "W/m^2")
    schema.add_variable("Surface Temperature", "K")
    schema.add_variable("Deep Ocean Temperature", "K")
    ghg = [f"Effective Radiative Forcing|{n}" for n in ("CO2", "CH4", "N2O")]
    schema.add_aggregate("Effective Radiative Forcing|GHG Mean", "W/m^2", "Mean", ghg + ["Effective Radiative Forcing|Other"])
    schema.add_aggregate("Effective Radiative Forcing|GHG Weighted", "W/m^2", "Weighted", ghg + ["Effective Radiative Forcing|Other"],
                         weights=[1.0, 0.5, 2.0, 0.25])
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Mean",
                         ["Effective Radiative Forcing|Aerosol|Direct", "Effective Radiative Forcing|Volcanic"], grid_type=GridType.Hemispheric)
    b = (ModelBuilder().with_time_axis(syn.time_axis(1900, 2000)).with_schema(schema)
         .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
         .with_rust_component(AerosolDirectBuilder.from_parameters({"sox_regional": [0.2, 0.5, 0.1, 0.2]}).build())
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}))
    if hemi_weights is not None:
        b.with_grid_weights(GridType.Hemispheric, hemi_weights)
    if four_box_weights is not None:
        b.with_grid_weights(GridType.FourBox, four_box_weights)
    return b


def mean_and_hemispheric_scenarios():
    years = syn.time_axis(1900, 2000).values()
    ramp = (years - 1900.0) / 100.0
    other = 0.3 * np.sin(years / 7.0)
    other[10:20] = np.nan          # a contributor that is missing for a decade: dropped from Mean / Weighted
    volc = np.column_stack([-0.5 * np.exp(-((years - 1963.0) / 2.0) ** 2), -0.2 * np.exp(-((years - 1991.0) / 2.0) ** 2)])
    volc[40:45, 1] = np.nan        # southern hemisphere missing: the Hemispheric Mean falls back to the other contributor there
    out = []
    for f in (1.0, 1.4):
        out.append({"Atmospheric Concentration|CO2": 296.0 + 75.0 * ramp * f, "Atmospheric Concentration|CH4": 900.0 + 850.0 * ramp,
                    "Atmospheric Concentration|N2O": 280.0 + 36.0 * ramp, "Emissions|SOx": 5.0 + 60.0 * ramp * f, "Emissions|BC": 2.5 + 5.0 * ramp,
                    "Emissions|OC": 10.0 + 20.0 * ramp, "Emissions|NOx": 10.0 + 30.0 * ramp,
                    "Effective Radiative Forcing|Other": other, "Effective Radiative Forcing|Volcanic": volc})
    return out


@pytest.mark.parametrize("hemi_weights,four_box_weights", [(None, None), ([0.6, 0.4], [0.3, 0.2, 0.4, 0.1])])
def test_mean_weighted_and_hemispheric_semantics_on_the_oracle(hemi_weights, four_box_weights):
    b = mean_and_hemispheric_builder(hemi_weights, four_box_weights)
    sc = mean_and_hemispheric_scenarios()[0]
    r = oracle_from_builder(b, sc).run()
    ghg = np.stack([r[f"Effective Radiative Forcing|{n}"] for n in ("CO2", "CH4", "N2O")])
    other = sc["Effective Radiative Forcing|Other"]
    # aggregates read their contributors at_end: index N+1 holds the mean of the values at N+1
    for t in (5, 15, 60):
        vals = list(ghg[:, t]) + ([] if np.isnan(other[t]) else [other[t]])
        assert r["Effective Radiative Forcing|GHG Mean"][t] == pytest.approx(sum(vals) / len(vals), rel=1e-14)
        w = [1.0, 0.5, 2.0] + ([] if np.isnan(other[t]) else [0.25])
        assert r["Effective Radiative Forcing|GHG Weighted"][t] == pytest.approx(sum(v * x for v, x in zip(vals, w)), rel=1e-14)
    direct = r["Effective Radiative Forcing|Aerosol|Direct"]
    assert direct.shape == (101, 2)                               # FourBox written, Hemispheric stored
    erf = r["Effective Radiative Forcing"]
    volc = sc["Effective Radiative Forcing|Volcanic"]
    assert erf.shape == (101, 2)
    assert erf[63, 0] == pytest.approx(0.5 * (direct[63, 0] + volc[63, 0]), rel=1e-14)
    assert erf[42, 1] == pytest.approx(direct[42, 1], rel=1e-14)   # NaN contributor dropped: mean of one
    assert np.isfinite(r["Surface Temperature"][1:]).all()


def test_grid_redefinition_without_a_schema_is_rejected():
    """verify_definition (model/validation.rs:30-90): without a schema two definitions of one variable on different grids
    are a GridTypeMismatch; the engine must not silently insert an aggregation."""
    from rscm_b200 import _ffi
    from rscm_b200.magicc import ClimateUDEBBuilder
    b = (ModelBuilder().with_time_axis(syn.time_axis(1900, 1910))
         .with_rust_component(ClimateUDEBBuilder.from_parameters({}).build())      # Surface Temperature: FourBox state
         .with_rust_component(OzoneForcingBuilder.from_parameters({}).build())     # Surface Temperature: Scalar input
         .with_initial_values({"Surface Temperature": 0.0}))
    with pytest.raises(_ffi.EngineError, match="grid type mismatch"):
        b.build_ensemble(device=-2)


@pytest.mark.gpu
@pytest.mark.parametrize("hemi_weights,four_box_weights", [(None, None), ([0.6, 0.4], [0.3, 0.2, 0.4, 0.1])])
def test_mean_weighted_and_hemispheric_gpu_parity(hemi_weights, four_box_weights, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = mean_and_hemispheric_builder(hemi_weights, four_box_weights)
    binds = {**syn.TWO_LAYER_BINDINGS, "adjust_ch4": "GhgForcing.adjust_ch4", "sox_coefficient": "AerosolDirect.sox_coefficient"}
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios(mean_and_hemispheric_scenarios())
    rng = np.random.default_rng(23)
    p = np.column_stack([syn.uniform_params(syn.TWO_LAYER_RANGES, 130, 4), rng.uniform(0.7, 1.0, 130), rng.uniform(-0.005, -0.002, 130)])
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, binds), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    assert got["Effective Radiative Forcing"].shape == (101, 2, 260) and got["Effective Radiative Forcing|Aerosol|Direct"].shape == (101, 2, 260)
