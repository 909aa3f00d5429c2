"""TOML front-end: the reference's layered-config behaviour (python/rscm/config/loader.py doctests,
tests/test_config*.py of the reference) and the ensemble one-liner built on it."""

import numpy as np
import pytest

from rscm_b200 import config as cfgmod
from rscm_b200.config import (build_ensemble, build_model, component_registry, deep_merge, load_config, load_config_layers,
                              prior_parameter_set)

DEFAULTS = """
[model]
name = "two-layer-default"
type = "two-layer"
version = "1.0.0"

[time]
start = 1750
end = 2100

[components.climate]
type = "TwoLayer"

[components.climate.parameters]
lambda0 = 1.0
a = 0.0
efficacy = 1.0
eta = 0.7
heat_capacity_surface = 8.0
heat_capacity_deep = 100.0
"""
HIGH_ECS = """
[model]
name = "two-layer-high-ecs"

[components.climate.parameters]
lambda0 = 0.7
efficacy = 1.3
"""


@pytest.fixture
def layered(tmp_path):
    (tmp_path / "defaults.toml").write_text(DEFAULTS)
    (tmp_path / "high-ecs.toml").write_text(HIGH_ECS)
    return tmp_path


def test_deep_merge_and_layers(layered, caplog):
    assert deep_merge({"a": 1, "nested": {"x": 1, "y": 2}}, {"b": 2, "nested": {"y": 3}}) == {"a": 1, "b": 2, "nested": {"x": 1, "y": 3}}
    assert deep_merge({"l": [1, 2]}, {"l": [3]}) == {"l": [3]}          # lists are replaced, not concatenated
    cfg = load_config_layers(layered / "defaults.toml", layered / "high-ecs.toml")
    p = cfg["components"]["climate"]["parameters"]
    assert cfg["model"]["name"] == "two-layer-high-ecs" and cfg["model"]["type"] == "two-layer"
    assert p["lambda0"] == 0.7 and p["efficacy"] == 1.3 and p["eta"] == 0.7 and cfg["time"] == {"start": 1750, "end": 2100}
    assert load_config_layers() == {}
    (layered / "odd.toml").write_text("[surprise]\nx = 1\n")
    with caplog.at_level("WARNING"):
        load_config(layered / "odd.toml")
    assert "Unknown configuration keys" in caplog.text and "surprise" in caplog.text


def test_registry_and_errors(layered):
    assert "TwoLayer" in component_registry and "ClimateUDEB" in component_registry and "CarbonCycle" in component_registry
    with pytest.raises(KeyError, match="Unknown component"):
        component_registry.get("NoSuchThing")
    with pytest.raises(ValueError, match="already registered"):
        component_registry.register("TwoLayer", object)
    with pytest.raises(ValueError, match="Unknown model type"):
        build_model({"model": {"type": "three-layer"}})
    cfg = load_config(layered / "defaults.toml")
    cfg["components"]["climate"]["parameters"]["lambda0"] = 9.0            # outside the metadata range (0.1, 5.0)
    with pytest.raises(ValueError, match="Invalid parameters"):
        build_model(cfg)


def test_ensemble_of_a_configured_model_host_side(layered):
    cfg = load_config_layers(layered / "defaults.toml", layered / "high-ecs.toml")
    ens, names = build_ensemble(cfg, ["lambda0", "efficacy"], device=-2)
    assert names == ["lambda0", "efficacy"] and ens.param_names == names and not ens.program_is_jit()
    assert ens.variable_names == ["Effective Radiative Forcing", "Surface Temperature", "Deep Ocean Temperature"] and ens.n_times == 351
    ens2, names2 = build_ensemble(cfg, {"l": "TwoLayer.lambda0", "T0": "initial:Surface Temperature"}, device=-2)
    assert names2 == ["l", "T0"]
    ps = prior_parameter_set("TwoLayer", names)
    assert ps.param_names == names and ps.bounds() == ([0.8, 1.0], [1.5, 1.8])
    assert prior_parameter_set("TwoLayer", ["a"], "range").bounds() == ([0.0], [1.0])
    assert cfgmod.PARAMETER_METADATA["TwoLayer"]["eta"]["default"] == 0.7


@pytest.mark.gpu
def test_configured_model_and_its_ensemble_agree(layered):
    from rscm_b200 import synthetic as syn
    cfg = load_config_layers(layered / "defaults.toml", layered / "high-ecs.toml")
    forcing = syn.ssp_like_forcing(syn.time_axis().values())
    ens, names = build_ensemble(cfg, ["lambda0", "efficacy"])
    sc = ens.pack_scenarios([{"Effective Radiative Forcing": forcing}])
    ens.select_outputs(["Surface Temperature"])
    out = ens.run(np.array([[0.7, 1.3], [1.2, 1.0]]), sc)
    # the single configured member through the reference-shaped Model API
    from rscm_b200.core import Timeseries, InterpolationStrategy
    from rscm_b200.config import _two_layer_builder
    b = _two_layer_builder(cfg).with_exogenous_variable(
        "Effective Radiative Forcing", Timeseries(forcing, syn.time_axis(), "W/m^2", InterpolationStrategy.Previous))
    model = b.build()
    model.run()
    t_model = model.timeseries().get_timeseries_by_name("Surface Temperature").values()
    assert np.array_equal(np.asarray(t_model).reshape(-1), out[:, 0]) and not np.array_equal(out[:, 0], out[:, 1])
