"""Extracts the golden vectors the reference's own regression tests hold for the hot path
into small fixtures that travel to the GPU box (/root/reference does not exist there).

Source: /root/reference/tests/regression/data/ghg_forcing/{01_concentration_driven,
02_ghg_forcing_olbl}.csv (+ *_config.json) — MAGICC7 (Fortran) concentrations -> ERF|CO2/CH4/N2O,
compared by tests/regression/test_ghg_forcing.py:237-331 at rtol 1e-5 / atol 1e-6 with
PI = first-year concentrations and actual[1:] vs expected[:-1]; 03_emissions_driven (every row: SSP245 emissions,
concentrations, temperature; :623-735, xfail at 5 % upstream), 04_ecs_sweep_{1.5,2.0,3.0,4.0,4.5} and 05_co2_only_forcing
(CO2 concentration, ERF, global-mean temperature; :738-830, phased 3-5 %) -> ghg_forcing_03_04_05.npz.
data/ocean_udeb/01..12 -> ocean_udeb.npz (all twelve scenarios; test_ocean_udeb.py:222-562).

Run once in the build container:  python tests/golden/make_golden.py
"""
import csv
import json
import os

import numpy as np

REF = "/root/reference/tests/regression/data/ghg_forcing"
HERE = os.path.dirname(os.path.abspath(__file__))

VARS = {
    "co2": "Atmospheric Concentrations|CO2",
    "ch4": "Atmospheric Concentrations|CH4",
    "n2o": "Atmospheric Concentrations|N2O",
    "erf_co2": "Effective Radiative Forcing|CO2",
    "erf_ch4": "Effective Radiative Forcing|CH4",
    "erf_n2o": "Effective Radiative Forcing|N2O",
}


def load(name):
    with open(os.path.join(REF, name + ".csv")) as f:
        rows = list(csv.reader(f))
    header = rows[0]
    ivar = header.index("variable")
    ireg = header.index("region")
    tcols = [i for i, h in enumerate(header) if h[:4].isdigit()]
    years = np.array([float(header[i][:4]) for i in tcols])
    out = {"years": years}
    for key, var in VARS.items():
        for r in rows[1:]:
            if r[ivar] == var and r[ireg] == "World":
                out[key] = np.array([float(r[i]) for i in tcols])
                break
        else:
            raise KeyError(var)
    with open(os.path.join(REF, name + "_config.json")) as f:
        cfg = json.load(f)
    return out, cfg


if __name__ == "__main__":
    for name in ("01_concentration_driven", "02_ghg_forcing_olbl"):
        data, cfg = load(name)
        keep = {k: cfg[k] for k in cfg if k.startswith("core_")}
        np.savez_compressed(os.path.join(HERE, f"ghg_forcing_{name[:2]}.npz"), config=json.dumps(keep), **data)
        print(name, {k: v.shape for k, v in data.items()}, keep)


# ---- ClimateUDEB: tests/regression/data/ocean_udeb/*.csv (MAGICC7 global-mean Surface Temperature under
# ABRUPT-2XCO2 step forcing; compared by tests/regression/test_ocean_udeb.py with phased 1-5 % tolerances) ----
UDEB = "/root/reference/tests/regression/data/ocean_udeb"
UDEB_SCENARIOS = ["01_diffusion_only", "02_constant_upwelling", "03_depth_dependent_area", "04_variable_upwelling",
                  "05_temp_dependent_diffusivity", "06_ground_heat", "07_interhemispheric_exchange", "08_sst_to_sat",
                  "09_time_varying_ecs", "10_full_default", "11_efficacy_ar6", "12_efficacy_ar6_1pctco2"]


def load_udeb(name):
    with open(os.path.join(UDEB, name + ".csv")) as f:
        rows = list(csv.reader(f))
    header = rows[0]
    ivar, ireg = header.index("variable"), header.index("region")
    tcols = [i for i, h in enumerate(header) if h[:4].isdigit()]
    years = np.array([float(header[i][:4]) for i in tcols])
    temp = None
    for r in rows[1:]:
        if r[ivar] == "Surface Temperature" and r[ireg] == "World":
            temp = np.array([float(r[i]) for i in tcols])
    with open(os.path.join(UDEB, name + "_config.json")) as f:
        cfg = json.load(f)
    return years, temp, {k: v for k, v in cfg.items() if k.startswith(("core_", "rf_efficacy", "startyear", "endyear"))}


if __name__ == "__main__":
    out = {}
    for name in UDEB_SCENARIOS:
        years, temp, cfg = load_udeb(name)
        out[name + "/years"] = years
        out[name + "/temp"] = temp
        out[name + "/config"] = json.dumps(cfg)
        print(name, years[0], years[-1], temp[-1], cfg)
    np.savez_compressed(os.path.join(HERE, "ocean_udeb.npz"), **out)


# ---- ghg_forcing 03 / 04 / 05: every World row of the CSV, keyed by variable name ----
def load_all_rows(name):
    with open(os.path.join(REF, name + ".csv")) as f:
        rows = list(csv.reader(f))
    header = rows[0]
    ivar, ireg, iunit = header.index("variable"), header.index("region"), header.index("unit")
    tcols = [i for i, h in enumerate(header) if h[:4].isdigit()]
    out = {"years": np.array([float(header[i][:4]) for i in tcols])}
    units = {}
    for r in rows[1:]:
        if r[ireg] == "World":
            out[r[ivar]] = np.array([float(r[i]) for i in tcols])
            units[r[ivar]] = r[iunit]
    with open(os.path.join(REF, name + "_config.json")) as f:
        cfg = json.load(f)
    return out, units, cfg


if __name__ == "__main__":
    out = {}
    for name in ["03_emissions_driven", "05_co2_only_forcing"] + [f"04_ecs_sweep_{e}" for e in ("1.5", "2.0", "3.0", "4.0", "4.5")]:
        data, units, cfg = load_all_rows(name)
        for k, v in data.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/config"] = json.dumps(cfg)
        out[f"{name}/units"] = json.dumps(units)
        print(name, len(data) - 1, "series", data["years"][0], data["years"][-1])
    np.savez_compressed(os.path.join(HERE, "ghg_forcing_03_04_05.npz"), **out)
