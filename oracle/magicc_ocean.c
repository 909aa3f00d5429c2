/*
 * magicc_ocean.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restates OceanCarbon of the reference (v0.5.0): crates/rscm-magicc/src/carbon/ocean.rs:63-260 and
 * crates/rscm-magicc/src/parameters/ocean_carbon.rs (IRF forms :99-130, presets, irf/scale_irf :378-397,
 * delta_pco2_from_dic, ocean_pco2, dic_conversion_factor).
 *
 * Parameter block (60 values): 0 model (informational), 1 co2_pi, 2 pco2_pi, 3 gas_exchange_scale,
 * 4 gas_exchange_tau, 5 temp_sensitivity, 6 irf_scale, 7 mixed_layer_depth, 8 ocean_surface_area, 9 sst_pi,
 * 10 steps_per_year, 11 max_history_months, 12 irf_switch_time,
 * 13 early kind (0 polynomial, 1 exponential sum), 14 early n, 15..22 early coefficients, 23..30 early timescales,
 * 31 late kind, 32 late n, 33..40 late coefficients, 41..48 late timescales,
 * 49..53 delta_ospp_offsets, 54..58 delta_ospp_coefficients, 59 enable_temp_feedback
 */
#include "orc_internal.h"

#include <stdlib.h>
#include <string.h>

#define OC_PPM_TO_GTC 2.124 /* carbon/ocean.rs:63 */
#define OC_MAXH 8192

typedef struct {
    int n;
    double flux[OC_MAXH];
} ocean_state;

static double irf_form(const double *f, double t)
{ /* IrfForm::evaluate — f = {kind, n, coef[8], tau[8]} */
    const int n = (int)f[1];
    if (f[0] == 0.0) {
        double r = 0.0;
        for (int i = n - 1; i >= 0; --i) r = r * t + f[2 + i];
        return r;
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += f[2 + i] * exp(-t / f[10 + i]);
    return s;
}

double orc_ocean_irf(const double *p, double t)
{ /* OceanCarbonParameters::irf + scale_irf */
    const double raw = (t < p[12]) ? irf_form(&p[13], t) : irf_form(&p[31], t);
    const double f = p[6];
    return (raw * f) / (raw * f + 1.0 - raw);
}

static void ocean_init(const double *p, void *vs)
{
    (void)p;
    ((ocean_state *)vs)->n = 0;
}

/* solve_impl / solve_ocean — carbon/ocean.rs:167-260.  inputs: [CO2 (get), SST (get), pCO2 (state), cumulative (state)];
 * outputs: [Carbon Flux|Ocean, Ocean Surface pCO2, Cumulative Ocean Uptake] */
static int ocean_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *vs)
{
    ocean_state *s = (ocean_state *)vs;
    const double co2 = orc_in_get(c, 0, 0), dsst = orc_in_get(c, 1, 0);
    double pco2 = orc_in_start(c, 2, 0), cumulative = orc_in_start(c, 3, 0);
    const double dt = t1 - t0;
    const int steps = (int)p[10], max_hist = (int)p[11];
    const double dt_month = dt / (double)steps;
    const double k_gas = p[3] / (p[4] * 12.0);                 /* gas_exchange_rate */
    const double dic_conv = 1.72e17 / (p[7] * p[8]);           /* dic_conversion_factor */
    double total_flux = 0.0;
    for (int m = 0; m < steps; ++m) {
        const double flux_ppm = k_gas * (co2 - pco2);
        if (s->n < OC_MAXH) s->flux[s->n++] = flux_ppm;
        if (s->n > max_hist) { memmove(s->flux, s->flux + 1, sizeof(double) * (size_t)(s->n - 1)); s->n--; }
        const double flux_gtc_yr = flux_ppm * 12.0 * OC_PPM_TO_GTC;
        total_flux += flux_gtc_yr / (double)steps;
        cumulative += flux_gtc_yr * dt_month;
        /* calculate_delta_dic */
        double integral = 0.0;
        const int n = s->n;
        for (int i = 0; i < n; ++i) {
            const double t_since = (double)(n - 1 - i) * (1.0 / 12.0);
            integral += s->flux[i] * orc_ocean_irf(p, t_since) * 1.0;
        }
        const double ddic = integral * dic_conv;
        /* delta_pco2_from_dic */
        const double d2 = ddic * ddic, d3 = d2 * ddic, d4 = d2 * d2, d5 = d4 * ddic;
        const double pw[5] = {ddic, d2 * 1e-3, -d3 * 1e-5, d4 * 1e-7, -d5 * 1e-10};
        double dp = 0.0;
        for (int i = 0; i < 5; ++i) dp += (p[49 + i] + p[54 + i] * p[9]) * pw[i];
        const double tfac = (p[59] != 0.0) ? exp(p[5] * dsst) : 1.0;
        pco2 = (p[2] + dp) * tfac;
    }
    out[0] = total_flux;
    out[1] = pco2;
    out[2] = cumulative;
    return 0;
}

static const orc_def ocean_defs[] = {
    {"Atmospheric Concentration|CO2", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Sea Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Carbon Flux|Ocean", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Ocean Surface pCO2", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Cumulative Ocean Uptake", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_ocean_carbon = {ORC_OCEAN_CARBON, "OceanCarbon", 5, ocean_defs, 60, ocean_solve, sizeof(ocean_state), ocean_init};
