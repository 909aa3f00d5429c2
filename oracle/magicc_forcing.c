/*
 * magicc_forcing.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restates crates/rscm-magicc/src/forcing/ghg.rs (reference v0.5.0).
 * Pinned by the MAGICC7 golden CSVs the reference's own regression tests use
 * (tests/regression/data/ghg_forcing/01_concentration_driven.csv and
 * 02_ghg_forcing_olbl.csv at rtol 1e-5 / atol 1e-6,
 * tests/regression/test_ghg_forcing.py:55-56,237-331).
 */
#include "orc_internal.h"

/* GhgForcing parameter block order (parameters/ghg_forcing.rs field order):
 *  0 method (0 = Ipcctar, 1 = Olbl), 1 co2_pi, 2 ch4_pi, 3 n2o_pi, 4 delq2xco2,
 *  5 ch4_radeff, 6 n2o_radeff, 7 olbl_co2_a1, 8 olbl_co2_b1, 9 olbl_co2_c1,
 *  10 olbl_co2_d1, 11 olbl_ch4_a3, 12 olbl_ch4_b3, 13 olbl_ch4_d3,
 *  14 olbl_n2o_a2, 15 olbl_n2o_b2, 16 olbl_n2o_c2, 17 olbl_n2o_d2,
 *  18 adjust_co2, 19 adjust_ch4, 20 adjust_n2o */
enum { G_METHOD, G_CO2_PI, G_CH4_PI, G_N2O_PI, G_DELQ2X, G_CH4_RADEFF, G_N2O_RADEFF,
       G_CO2_A1, G_CO2_B1, G_CO2_C1, G_CO2_D1, G_CH4_A3, G_CH4_B3, G_CH4_D3,
       G_N2O_A2, G_N2O_B2, G_N2O_C2, G_N2O_D2, G_ADJ_CO2, G_ADJ_CH4, G_ADJ_N2O, G_NPARAM };

/* overlap_f — ghg.rs:122-125 */
static double overlap_f(double ch4, double n2o)
{
    const double mn = ch4 * n2o;
    return 0.47 * log(1.0 + 2.01e-5 * pow(mn, 0.75) + 5.31e-15 * ch4 * pow(mn, 1.52));
}

/* ghg.rs:164-200 (IPCCTAR) and :205-259 (OLBL); calculate_forcings :264-279 */
void orc_ghg_forcings(const double *p, double co2, double ch4, double n2o, double *out)
{
    double co2_raw, ch4_raw, n2o_raw;
    if (p[G_METHOD] == 0.0) {
        const double alpha = p[G_DELQ2X] / log(2.0);
        co2_raw = alpha * log(co2 / p[G_CO2_PI]);
        {
            const double direct = p[G_CH4_RADEFF] * (sqrt(ch4) - sqrt(p[G_CH4_PI]));
            const double overlap = overlap_f(ch4, p[G_N2O_PI]) - overlap_f(p[G_CH4_PI], p[G_N2O_PI]);
            ch4_raw = direct - overlap;
        }
        {
            const double direct = p[G_N2O_RADEFF] * (sqrt(n2o) - sqrt(p[G_N2O_PI]));
            const double overlap = overlap_f(p[G_CH4_PI], n2o) - overlap_f(p[G_CH4_PI], p[G_N2O_PI]);
            n2o_raw = direct - overlap;
        }
    } else {
        const double co2_pi = p[G_CO2_PI];
        const double delta = co2 - co2_pi;
        const double n2o_overlap = p[G_CO2_C1] * sqrt(n2o);
        const double c_max = co2_pi - p[G_CO2_B1] / (2.0 * p[G_CO2_A1]);
        double alpha;
        if (co2 >= c_max)
            alpha = -p[G_CO2_B1] * p[G_CO2_B1] / (4.0 * p[G_CO2_A1]) + p[G_CO2_D1] + n2o_overlap;
        else if (co2 <= co2_pi)
            alpha = p[G_CO2_D1] + n2o_overlap;
        else
            alpha = p[G_CO2_A1] * delta * delta + p[G_CO2_B1] * delta + p[G_CO2_D1] + n2o_overlap;
        co2_raw = alpha * log(co2 / co2_pi);
        {
            const double coeff = p[G_CH4_A3] * sqrt(ch4) + p[G_CH4_B3] * sqrt(n2o) + p[G_CH4_D3];
            ch4_raw = coeff * (sqrt(ch4) - sqrt(p[G_CH4_PI]));
        }
        {
            const double coeff = p[G_N2O_A2] * sqrt(co2) + p[G_N2O_B2] * sqrt(n2o)
                               + p[G_N2O_C2] * sqrt(ch4) + p[G_N2O_D2];
            n2o_raw = coeff * (sqrt(n2o) - sqrt(p[G_N2O_PI]));
        }
    }
    out[0] = co2_raw * p[G_ADJ_CO2];
    out[1] = ch4_raw * p[G_ADJ_CH4];
    out[2] = n2o_raw * p[G_ADJ_N2O];
}

/* Component::solve — ghg.rs:296-318 */
static int ghg_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    orc_ghg_forcings(p, orc_in_get(c, 0, 0), orc_in_get(c, 1, 0), orc_in_get(c, 2, 0), out);
    return 0;
}

static const orc_def ghg_defs[] = {
    {"Atmospheric Concentration|CO2", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|CH4", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|N2O", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|CO2", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|CH4", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|N2O", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_ghg_forcing = {ORC_GHG_FORCING, "GhgForcing", 6, ghg_defs, G_NPARAM,
                                            ghg_solve, 0, NULL};
