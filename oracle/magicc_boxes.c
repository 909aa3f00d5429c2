/*
 * magicc_boxes.c — CPU ORACLE (test infrastructure, NOT product code).
 * Scalar box components restated from the reference (v0.5.0):
 *   crates/rscm-components/src/components/four_box_ocean_heat_uptake.rs
 *   crates/rscm-components/src/components/ocean_carbon_cycle/ocean_surface_partial_pressure.rs
 *   crates/rscm-magicc/src/carbon/budget.rs        (CO2Budget)
 *   crates/rscm-magicc/src/carbon/terrestrial.rs   (TerrestrialCarbon) + parameters/terrestrial_carbon.rs
 *   crates/rscm-magicc/src/chemistry/ch4.rs        (CH4Chemistry, Prather iterations) + parameters/ch4_chemistry.rs
 *   crates/rscm-magicc/src/chemistry/n2o.rs        (N2OChemistry) + parameters/n2o_chemistry.rs
 * Pinned by the reference's unit-test known answers (OSPP rstest cases 339.089 / 381.003, budget mass
 * conservation, steady states) — see tests/test_magicc_boxes.py.
 */
#include "orc_internal.h"

/* ---- FourBoxOceanHeatUptake: params = 4 regional ratios ---------------------------------------------- */
static int fbohu_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double erf = orc_in_get(c, 0, 0);
    for (int i = 0; i < 4; ++i) out[i] = erf * p[i];
    return 0;
}
static const orc_def fbohu_defs[] = {
    {"Effective Radiative Forcing|Aggregated", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Heat Uptake|Ocean", ORC_REQ_OUTPUT, ORC_GRID_FOUR_BOX},
};
const orc_kind_info orc_kind_fbohu = {ORC_FOUR_BOX_OHU, "FourBoxOceanHeatUptake", 2, fbohu_defs, 4, fbohu_solve, 0, NULL};

/* ---- OceanSurfacePartialPressure: params = ospp_preindustrial, sensitivity_ospp_to_temperature,
 *      sea_surface_temperature_preindustrial, delta_ospp_offsets[5], delta_ospp_coefficients[5] ------------- */
static int ospp_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double dsst = orc_in_get(c, 0, 0), d = orc_in_get(c, 1, 0);
    /* calculate_ospp — ocean_surface_partial_pressure.rs (powi by repeated multiplication) */
    const double d2 = d * d, d3 = d2 * d, d4 = d2 * d2;
    const double bits[5] = {d, d2 * 10e-3, -d3 * 10e-5, d4 * 10e-7, -d4 * 10e-10};
    double dot = 0.0;
    for (int i = 0; i < 5; ++i) dot += (p[3 + i] + p[8 + i] * p[2]) * bits[i];
    out[0] = (p[0] + dot) * exp(p[1] * dsst);
    return 0;
}
static const orc_def ospp_defs[] = {
    {"Sea Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Dissolved Inorganic Carbon", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Ocean Surface Partial Pressure|CO2", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_ospp = {ORC_OCEAN_SURFACE_PP, "OceanSurfacePartialPressure", 3, ospp_defs, 13, ospp_solve, 0, NULL};

/* ---- CO2Budget: params = gtc_per_ppm, co2_pi — carbon/budget.rs:121-205 ---------------------------------- */
static int budget_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    const double fossil = orc_in_get(c, 0, 0), landuse = orc_in_get(c, 1, 0);
    const double terrestrial = orc_in_get(c, 2, 0), ocean = orc_in_get(c, 3, 0);
    const double co2 = orc_in_start(c, 4, 0);
    const double dt = t1 - t0;
    const double total_emissions = fossil + landuse;
    const double total_uptake = terrestrial + ocean;
    const double net = total_emissions - total_uptake;
    out[0] = net;
    out[1] = (total_emissions > 0.0) ? net / total_emissions : 0.0;
    out[2] = co2 + (net * dt) / p[0];
    return 0;
}
static const orc_def budget_defs[] = {
    {"Emissions|CO2|Fossil", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|CO2|Land Use", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Carbon Flux|Terrestrial", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Carbon Flux|Ocean", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|CO2|Net", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Airborne Fraction|CO2", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|CO2", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_co2_budget = {ORC_CO2_BUDGET, "CO2Budget", 7, budget_defs, 2, budget_solve, 0, NULL};

/* ---- TerrestrialCarbon — carbon/terrestrial.rs:213-343; parameters/terrestrial_carbon.rs ---------------- */
enum { TC_NPP_PI, TC_CO2_PI, TC_BETA, TC_NPP_TS, TC_RESP_TS, TC_DET_TS, TC_SOIL_TS, TC_HUM_TS, TC_PLANT_PI, TC_DET_PI, TC_SOIL_PI,
       TC_HUM_PI, TC_RESP_PI, TC_F_NPP_PLANT, TC_F_NPP_DET, TC_F_PLANT_DET, TC_F_DET_SOIL, TC_F_SOIL_HUM, TC_FERT_ON, TC_TEMP_ON, TC_N };

static void tc_pool_step(double pool, double tau, double flux_in, double temp_factor, double dt, double *new_pool, double *turnover)
{ /* implicit_pool_step */
    const double k_eff = temp_factor / tau;
    const double half_k = 0.5 * k_eff * dt;
    double np = ((1.0 - half_k) * pool + flux_in * dt) / (1.0 + half_k);
    np = fmax(np, 0.0);
    *new_pool = np;
    *turnover = 0.5 * k_eff * (pool + np);
}

static int terrestrial_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    const double co2 = orc_in_get(c, 0, 0), temperature = orc_in_get(c, 1, 0), landuse = orc_in_get(c, 2, 0);
    const double plant = orc_in_start(c, 3, 0), detritus = orc_in_start(c, 4, 0), soil = orc_in_start(c, 5, 0), humus = orc_in_start(c, 6, 0);
    const double dt = t1 - t0;
    double fert = 1.0;
    if (p[TC_FERT_ON] != 0.0 && !(co2 <= 0.0)) fert = fmax(1.0 + p[TC_BETA] * log(co2 / p[TC_CO2_PI]), 0.1);
    const int tf = p[TC_TEMP_ON] != 0.0;
#define TFAC(s) (tf ? exp((s) * temperature) : 1.0)
    const double npp = p[TC_NPP_PI] * fert * TFAC(p[TC_NPP_TS]);
    const double respiration = p[TC_RESP_PI] * fert * TFAC(p[TC_RESP_TS]);
    const double tf_det = TFAC(p[TC_DET_TS]), tf_soil = TFAC(p[TC_SOIL_TS]), tf_hum = TFAC(p[TC_HUM_TS]);
#undef TFAC
    /* turnover times — parameters/terrestrial_carbon.rs */
    const double frac_npp_to_soil = fmax(1.0 - p[TC_F_NPP_PLANT] - p[TC_F_NPP_DET], 0.0);
    const double net_flux_plant = p[TC_F_NPP_PLANT] * p[TC_NPP_PI] - p[TC_RESP_PI];
    const double tau_plant = (net_flux_plant > 1e-10) ? p[TC_PLANT_PI] / net_flux_plant : 100.0;
    const double flux_into_det = p[TC_F_NPP_DET] * p[TC_NPP_PI] + p[TC_F_PLANT_DET] * net_flux_plant;
    const double tau_det = (flux_into_det > 1e-10) ? p[TC_DET_PI] / flux_into_det : 3.0;
    const double flux_det_out = p[TC_DET_PI] / tau_det;
    const double flux_into_soil = frac_npp_to_soil * p[TC_NPP_PI] + (1.0 - p[TC_F_PLANT_DET]) * net_flux_plant + p[TC_F_DET_SOIL] * flux_det_out;
    const double tau_soil = (flux_into_soil > 1e-10) ? p[TC_SOIL_PI] / flux_into_soil : 50.0;
    const double flux_soil_out = p[TC_SOIL_PI] / tau_soil;
    const double flux_into_hum = p[TC_F_SOIL_HUM] * flux_soil_out;
    const double tau_hum = (flux_into_hum > 1e-10) ? p[TC_HUM_PI] / flux_into_hum : 1000.0;
    /* solve_pools */
    double new_plant, to_plant, new_det, to_det, new_soil, to_soil, new_hum, to_hum;
    tc_pool_step(plant, tau_plant, npp * p[TC_F_NPP_PLANT] - respiration - landuse, 1.0, dt, &new_plant, &to_plant);
    tc_pool_step(detritus, tau_det, npp * p[TC_F_NPP_DET] + p[TC_F_PLANT_DET] * to_plant, tf_det, dt, &new_det, &to_det);
    const double flux_in_soil = npp * frac_npp_to_soil + (1.0 - p[TC_F_PLANT_DET]) * to_plant + p[TC_F_DET_SOIL] * to_det;
    tc_pool_step(soil, tau_soil, flux_in_soil, tf_soil, dt, &new_soil, &to_soil);
    tc_pool_step(humus, tau_hum, p[TC_F_SOIL_HUM] * to_soil, tf_hum, dt, &new_hum, &to_hum);
    const double det_to_atm = (1.0 - p[TC_F_DET_SOIL]) * to_det;
    const double soil_to_atm = (1.0 - p[TC_F_SOIL_HUM]) * to_soil;
    const double total_resp = respiration + det_to_atm + soil_to_atm + to_hum;
    out[0] = npp - total_resp - landuse;
    out[1] = new_plant; out[2] = new_det; out[3] = new_soil; out[4] = new_hum;
    return 0;
}
static const orc_def terrestrial_defs[] = {
    {"Atmospheric Concentration|CO2", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|CO2|Land Use", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Carbon Flux|Terrestrial", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Carbon Pool|Plant", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Carbon Pool|Detritus", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Carbon Pool|Soil", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Carbon Pool|Humus", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_terrestrial = {ORC_TERRESTRIAL_CARBON, "TerrestrialCarbon", 8, terrestrial_defs, TC_N, terrestrial_solve, 0, NULL};

/* ---- CH4Chemistry — chemistry/ch4.rs:55-350 -------------------------------------------------------------- */
enum { CH_PI, CH_NAT, CH_TAU_OH, CH_TAU_SOIL, CH_TAU_STRAT, CH_TAU_CL, CH_SELF, CH_GAMMA, CH_S_NOX, CH_S_CO, CH_S_VOC, CH_TEMP_S,
       CH_TEMP_ON, CH_EMIS_ON, CH_PPB_TG, CH_NOX_REF, CH_CO_REF, CH_VOC_REF, CH_N };

static int ch4_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double ch4_current = orc_in_start(c, 5, 0);
    int ok;
    double ch4_prev = orc_in_offset(c, 5, 0, -1, &ok); /* previous().unwrap_or(current) */
    if (!ok) ch4_prev = ch4_current;
    const double emissions = orc_in_get(c, 0, 0), temperature = orc_in_get(c, 1, 0);
    const double nox = orc_in_get(c, 2, 0), co = orc_in_get(c, 3, 0), nmvoc = orc_in_get(c, 4, 0);
    /* solve_concentration :245-300 */
    const double total_emissions = emissions + p[CH_NAT];
    const double burden_prev = ch4_prev * p[CH_PPB_TG];
    const double burden_ref = p[CH_PI] * p[CH_PPB_TG];
    const double tau_other = 1.0 / (1.0 / p[CH_TAU_SOIL] + 1.0 / p[CH_TAU_STRAT] + 1.0 / p[CH_TAU_CL]);
    double base = p[CH_TAU_OH];
    if (p[CH_EMIS_ON] != 0.0) {
        const double ex = -p[CH_GAMMA] * (p[CH_S_NOX] * (nox - p[CH_NOX_REF]) + p[CH_S_CO] * (co - p[CH_CO_REF]) + p[CH_S_VOC] * (nmvoc - p[CH_VOC_REF]));
        base = p[CH_TAU_OH] * exp(ex);
    }
    const double x = -p[CH_GAMMA] * p[CH_SELF];
    double burden = ch4_current * p[CH_PPB_TG];
    double delta_burden = 0.0, tau_oh = p[CH_TAU_OH];
    for (int i = 0; i < 4; ++i) { /* prather_iteration :207-243 */
        const double burden_mean = (burden + burden_prev) / 2.0;
        double tau = base * pow(fmax(burden_mean / burden_ref, 1.0), x);
        if (i > 0 && !(fabs(burden_prev) < 1e-10)) tau = tau * (1.0 - 0.5 * x * delta_burden / burden_prev);
        if (p[CH_TEMP_ON] != 0.0 && !(fabs(temperature) < 1e-10))
            tau = p[CH_TAU_OH] / (p[CH_TAU_OH] / tau + p[CH_TEMP_S] * fmax(temperature, 0.0));
        delta_burden = total_emissions - burden_mean / tau - burden_mean / tau_other;
        burden = burden_prev + delta_burden;
        tau_oh = tau;
    }
    out[0] = 1.0 / (1.0 / tau_oh + 1.0 / tau_other);
    out[1] = burden / p[CH_PPB_TG];
    return 0;
}
static const orc_def ch4_defs[] = {
    {"Emissions|CH4", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|NOx", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|CO", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|NMVOC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Lifetime|CH4", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|CH4", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_ch4 = {ORC_CH4_CHEMISTRY, "CH4Chemistry", 7, ch4_defs, CH_N, ch4_solve, 0, NULL};

/* ---- N2OChemistry — chemistry/n2o.rs:171-275: params n2o_pi, natural_emissions, tau_n2o, lifetime_feedback,
 *      strat_delay, ppb_to_tg ------------------------------------------------------------------------------ */
static int n2o_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    const double dt = t1 - t0;
    const double cur = orc_in_start(c, 1, 0);
    int ok;
    double prev = orc_in_offset(c, 1, 0, -1, &ok);
    if (!ok) prev = cur;
    int delay = (int)p[4];
    if (delay < 1) delay = 1;
    double t_delay = orc_in_offset(c, 1, 0, -delay, &ok);
    if (!ok) t_delay = prev;
    double t_delay_m1 = orc_in_offset(c, 1, 0, -(delay + 1), &ok);
    if (!ok) t_delay_m1 = t_delay;
    const double lagged = (t_delay + t_delay_m1) / 2.0;
    const double total_emissions = orc_in_get(c, 0, 0) + p[1];
    const double burden_prev = prev * p[5], burden_lagged = lagged * p[5], burden_ref = p[0] * p[5];
    double burden = cur * p[5], tau_eff = p[2];
    for (int i = 0; i < 4; ++i) {
        const double mid = (burden_prev + burden) / 2.0;
        tau_eff = p[2] * pow(fmax(mid / burden_ref, 1.0), p[3]);
        burden = burden_prev + (total_emissions - burden_lagged / tau_eff) * dt;
    }
    out[0] = tau_eff;
    out[1] = burden / p[5];
    return 0;
}
static const orc_def n2o_defs[] = {
    {"Emissions|N2O", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Lifetime|N2O", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|N2O", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_n2o = {ORC_N2O_CHEMISTRY, "N2OChemistry", 3, n2o_defs, 6, n2o_solve, 0, NULL};
