"""HalocarbonChemistry (41 species, exponential decay + forcing + EESC): the reference's own unit tests
(crates/rscm-magicc/src/chemistry/halocarbon.rs:358-741, parameters/halocarbon.rs:266-391) restated against the oracle through
full model runs, the host graph compiler's view of the 86-variable component, and GPU parity at 1e-9 — alone and inside the
emissions-driven MAGICC graph where its EESC drives OzoneForcing and its forcing joins the ERF aggregate."""

import numpy as np
import pytest

from rscm_b200 import _ffi, synthetic as syn
from rscm_b200.core import ModelBuilder
from rscm_b200.magicc import HalocarbonChemistryBuilder as HB

from .helpers import oracle_bindings, oracle_from_builder, rel_err

SPECIES = HB.species_names()
TABLE = {s[0]: s[1:] for s in HB.FGASES + HB.MONTREAL_GASES}
CONV = lambda mw: (28.97 / mw) * (1e9 / (5.133e9 * 1e12)) * 1e12 / 0.949  # noqa: E731  parameters/halocarbon.rs:162-172


def conc(s):
    return f"Atmospheric Concentration|{s}"


def halo_builder(start=1850, end=1860, initial=None, params=None):
    init = {conc(s): 0.0 for s in SPECIES}
    init.update({conc(k): v for k, v in (initial or {}).items()})
    return (ModelBuilder().with_time_axis(syn.time_axis(start, end))
            .with_rust_component(HB.from_parameters(params or {}).build()).with_initial_values(init))


def halo_scenario(n, emissions=None, default=0.0):
    sc = {f"Emissions|{s}": np.full(n, default) for s in SPECIES}
    for k, v in (emissions or {}).items():
        sc[f"Emissions|{k}"] = np.broadcast_to(np.asarray(v, dtype=float), (n,)).copy()
    return sc


def run(initial=None, emissions=None, start=1850, end=1860, default=0.0):
    b = halo_builder(start, end, initial)
    return oracle_from_builder(b, halo_scenario(end - start + 1, emissions, default)).run()


# ---- the reference's unit tests --------------------------------------------------------------------------------
def test_definitions_count_and_names():
    m = oracle_from_builder(halo_builder(), halo_scenario(11))
    assert len(m.names) == 41 * 2 + 4                                           # test_definitions_count
    for n in ("Emissions|CFC-11", conc("CFC-11"), "Emissions|HFC-134a", conc("HFC-134a"), "Forcing|Halocarbons", "EESC"):
        assert n in m.names                                                      # test_definitions_contain_expected_names
    assert len(HB.FGASES) == 23 and len(HB.MONTREAL_GASES) == 18                 # parameters tests :272-281
    assert all(s[7] == 0.0 for s in HB.FGASES)                                   # test_fgases_have_zero_release
    assert 0.01 < CONV(137.4) < 0.1                                              # test_emission_conversion_factor


def test_exponential_decay_without_emissions():
    r = run({"CF4": 100.0, "HFC-152a": 100.0})
    assert abs(r[conc("CF4")][1] - 100.0 * np.exp(-1.0 / 50000.0)) < 1e-10      # test_exponential_decay_no_emissions
    assert abs(r[conc("CF4")][1] - 100.0) / 100.0 < 0.001
    assert abs(r[conc("HFC-152a")][1] - 100.0 * np.exp(-1.0 / 1.6)) < 1e-8      # test_exponential_decay_short_lived
    assert r[conc("HFC-152a")][1] < 60.0


def test_decay_with_emissions_reaches_equilibrium():
    r = run(emissions={"HFC-134a": 100.0}, end=1950)
    eq = 100.0 * CONV(102.0) * 14.0
    assert abs(r[conc("HFC-134a")][-1] - eq) / eq < 0.01                         # test_decay_with_emissions_equilibrium


def test_zero_emissions_decays_to_zero():
    r = run({"CFC-11": 250.0}, end=2250)
    assert r[conc("CFC-11")][-1] < 0.25                                          # test_zero_emissions_decays_to_zero
    assert np.all(np.diff(r[conc("CFC-11")]) < 0.0)


def test_forcing_is_linear_in_concentration_and_splits_by_family():
    pi = {s: TABLE[s][2] for s in SPECIES}
    r0 = run(pi, end=1851)
    # a step of decay moves CH3Cl / CH3Br off their pre-industrial level: compare with the formula on the new concentrations
    for r in (r0, run({**pi, "CFC-11": 100.0, "HFC-134a": 80.0, "SF6": 10.0}, end=1851)):
        f = {s: (r[conc(s)][1] - TABLE[s][2]) * TABLE[s][1] / 1000.0 for s in SPECIES}   # species_forcing :137-145
        fg, mt = sum(f[s[0]] for s in HB.FGASES), sum(f[s[0]] for s in HB.MONTREAL_GASES)
        assert r["Forcing|F-gases"][1] == pytest.approx(fg, rel=1e-12, abs=1e-15)
        assert r["Forcing|Montreal Gases"][1] == pytest.approx(mt, rel=1e-12, abs=1e-15)
        assert r["Forcing|Halocarbons"][1] == pytest.approx(fg + mt, rel=1e-10, abs=1e-15)  # test_total_forcing_calculation
    assert np.isnan(r0["Forcing|Halocarbons"][0]) and np.isnan(r0["EESC"][0])    # pure outputs: NaN at index 0


def test_eesc_weights():
    r = run({"CFC-11": 200.0}, end=1851)
    assert r["EESC"][1] == pytest.approx(r[conc("CFC-11")][1] * 3.0 * 1.0, rel=1e-12)            # test_eesc_calculation_basic
    r = run({"Halon-1301": 3.0}, end=1851)
    assert r["EESC"][1] == pytest.approx(r[conc("Halon-1301")][1] * 60.0 * (0.28 / 0.47), rel=1e-12)  # test_eesc_bromine_contribution
    r = run({"CF4": 80.0, "SF6": 10.0, "HFC-134a": 100.0}, end=1851)
    assert r["EESC"][1] == 0.0                                                   # test_fgases_no_eesc_contribution
    r = run({"CH2Cl2": 50.0}, end=1851)
    assert r["EESC"][1] == 0.0                                                   # fractional_release == 0 contributes nothing (:211)


def test_step_all_species_and_realistic_levels():
    r = run({s: TABLE[s][2] for s in SPECIES}, end=1851, default=0.1)
    assert all(r[conc(s)][1] > 0.0 for s in SPECIES if TABLE[s][2] == 0.0)       # test_step_concentrations_all_species
    r = run({"CFC-11": 230.0, "CFC-12": 510.0, "CFC-113": 70.0, "HCFC-22": 245.0, "HFC-134a": 100.0, "SF6": 10.0, "CH3Cl": 540.0,
             "CH3Br": 5.0}, end=1851)
    assert 0.1 < r["Forcing|Halocarbons"][1] < 1.0 and 500.0 < r["EESC"][1] < 5000.0   # test_realistic_scenario


def test_a_nan_emission_does_not_stick():
    """solve reads through InputState::get_global (state/mod.rs:231-254): an endogenous series yields its latest non-NaN value,
    so the NaN written by a step with NaN emissions is skipped by the next step (code reading; the reference has no test for it)."""
    e = np.full(11, 10.0)
    e[4] = np.nan
    r = run({"CFC-11": 100.0}, emissions={"CFC-11": e})
    c = r[conc("CFC-11")]
    assert np.isnan(c[5]) and np.all(np.isfinite(np.delete(c, 5)))
    d, k = np.exp(-1.0 / 52.0), 10.0 * CONV(137.4) * 52.0
    assert c[6] == pytest.approx(c[4] * d + k * (1.0 - d), rel=1e-12)            # step 5 restarts from the value at index 4


# ---- host graph compiler -------------------------------------------------------------------------------------------
def test_engine_sees_the_same_graph_as_the_oracle():
    b = halo_builder()
    ens = b.build_ensemble(device=-2)
    m = oracle_from_builder(b, halo_scenario(11))
    assert ens.variable_names == m.names and ens.execution_order() == m.execution_order() and ens.program_is_jit()
    assert ens.exogenous_names == [f"Emissions|{s}" for s in SPECIES]
    with pytest.raises(_ffi.EngineError, match="per-graph"):
        ens.bind_parameters({"x": "HalocarbonChemistry.br_multiplier"})
    with pytest.raises(ValueError, match="default fgases list"):
        HB.from_parameters({"fgases": [dict(name="X", lifetime=1, radiative_efficiency=1, concentration_pi=0, molecular_weight=1,
                                            n_cl=0, n_br=0, fractional_release=0)]}).build()


# ---- GPU parity ------------------------------------------------------------------------------------------------------
def ramp_scenario(n, f=1.0):
    t = np.arange(n) / 100.0
    return {f"Emissions|{s}": f * (1.0 + i % 7) * (0.5 + 4.0 * t) for i, s in enumerate(SPECIES)}


@pytest.mark.gpu
def test_halocarbon_gpu_parity(tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = halo_builder(end=1950, initial={"CH3Cl": 500.0, "CH3Br": 5.0})
    binds = {"cfc11": "initial:" + conc("CFC-11"), "sf6": "initial:" + conc("SF6")}
    ens = b.build_ensemble().bind_parameters(binds)
    s1, s2 = ramp_scenario(101), ramp_scenario(101, 1.7)
    s2["Emissions|CFC-12"][30] = np.nan   # latest_value path
    sc = ens.pack_scenarios([s1, s2])
    p = syn.uniform_params({"cfc11": (0.0, 300.0), "sf6": (0.0, 12.0)}, 80, 3)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, binds), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    c12 = got[conc("CFC-12")]
    assert np.isnan(c12[31, 80:]).all() and np.isfinite(c12[32:, 80:]).all() and np.isfinite(c12[:, :80]).all()


@pytest.mark.gpu
def test_full_magicc_with_halocarbons_gpu_parity(tmp_path, monkeypatch):
    """The emissions-driven chain of the reference's regression test (tests/regression/test_ghg_forcing.py:470-620) with the
    halocarbon box wired in: HalocarbonChemistry -> EESC -> OzoneForcing, Forcing|Halocarbons -> ERF aggregate -> ClimateUDEB."""
    from .test_ocean_carbon import FULL_BINDS, full_magicc_builder, full_magicc_scenario
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = full_magicc_builder(end=1920, halocarbons=True)
    ens = b.build_ensemble().bind_parameters(FULL_BINDS)
    scen = []
    for f in (1.0, 1.3):
        s = full_magicc_scenario(end=1920, f=f)
        s.pop("EESC")
        s.update(ramp_scenario(71, f))
        scen.append(s)
    sc = ens.pack_scenarios(scen)
    p = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, 48, 43)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, FULL_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    assert np.isfinite(got["Surface Temperature"]).all() and np.all(got["EESC"][1:] > 0.0)
