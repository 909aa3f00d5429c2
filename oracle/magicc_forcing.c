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

/* ------------------------------------------------------------------------- */
/* OzoneForcing — crates/rscm-magicc/src/forcing/ozone.rs                     */
/* params (parameters/ozone_forcing.rs field order): eesc_reference,          */
/*  strat_o3_scale, strat_cl_exponent, trop_radeff, trop_oz_ch4, trop_oz_nox, */
/*  trop_oz_co, trop_oz_voc, ch4_pi, nox_pi, co_pi, nmvoc_pi,                 */
/*  temp_feedback_scale                                                       */
/* ------------------------------------------------------------------------- */
static int ozone_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double eesc = orc_in_get(c, 0, 0), ch4 = orc_in_get(c, 1, 0), nox = orc_in_get(c, 2, 0);
    const double co = orc_in_get(c, 3, 0), nmvoc = orc_in_get(c, 4, 0), temperature = orc_in_get(c, 5, 0);
    /* calculate_strat_forcing — ozone.rs:187-203 */
    const double delta_eesc = eesc - p[0];
    out[0] = (delta_eesc <= 0.0) ? 0.0 : p[1] * pow(delta_eesc / 100.0, p[2]);
    /* calculate_trop_forcing — ozone.rs:234-262 */
    const double ch4_term = (ch4 > 0.0 && p[8] > 0.0) ? p[4] * log(ch4 / p[8]) : 0.0;
    const double precursor = p[5] * (nox - p[9]) + p[6] * (co - p[10]) + p[7] * (nmvoc - p[11]);
    out[1] = p[3] * (ch4_term + precursor);
    out[2] = p[12] * temperature;
    return 0;
}
static const orc_def ozone_defs[] = {
    {"EESC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|CH4", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|NOx", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|CO", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|NMVOC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|O3|Stratospheric", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|O3|Tropospheric", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|O3|Temperature Feedback", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_ozone_forcing = {ORC_OZONE_FORCING, "OzoneForcing", 9, ozone_defs, 13, ozone_solve, 0, NULL};

/* ------------------------------------------------------------------------- */
/* AerosolDirect — crates/rscm-magicc/src/forcing/aerosol_direct.rs           */
/* params: sox_coefficient, bc_coefficient, oc_coefficient,                   */
/*  nitrate_coefficient, sox_regional[4], bc_regional[4], oc_regional[4],     */
/*  nitrate_regional[4], sox_pi, bc_pi, oc_pi, nox_pi, harmonize,             */
/*  harmonize_year, harmonize_target (the last three are carried, unused by   */
/*  the reference's solve)                                                    */
/* ------------------------------------------------------------------------- */
static int aerosol_direct_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double sox = orc_in_get(c, 0, 0), bc = orc_in_get(c, 1, 0), oc = orc_in_get(c, 2, 0), nox = orc_in_get(c, 3, 0);
    /* calculate_species_forcing — :150-170 */
    const double f_sox = p[0] * (sox - p[20]), f_bc = p[1] * (bc - p[21]);
    const double f_oc = p[2] * (oc - p[22]), f_nit = p[3] * (nox - p[23]);
    const double total = f_sox + f_bc + f_oc + f_nit;
    /* distribute_regional — :172-192 */
    if (fabs(total) < 1e-15) { for (int i = 0; i < 4; ++i) out[i] = 0.0; return 0; }
    const double total_abs = fabs(f_sox) + fabs(f_bc) + fabs(f_oc) + fabs(f_nit);
    if (total_abs < 1e-15) { for (int i = 0; i < 4; ++i) out[i] = total / 4.0; return 0; }
    for (int i = 0; i < 4; ++i) {
        const double pattern = (fabs(f_sox) * p[4 + i] + fabs(f_bc) * p[8 + i] + fabs(f_oc) * p[12 + i] + fabs(f_nit) * p[16 + i]) / total_abs;
        out[i] = total * pattern;
    }
    return 0;
}
static const orc_def aerosol_direct_defs[] = {
    {"Emissions|SOx", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|BC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|OC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|NOx", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|Aerosol|Direct", ORC_REQ_OUTPUT, ORC_GRID_FOUR_BOX},
};
const orc_kind_info orc_kind_aerosol_direct = {ORC_AEROSOL_DIRECT, "AerosolDirect", 5, aerosol_direct_defs, 27, aerosol_direct_solve, 0, NULL};

/* ------------------------------------------------------------------------- */
/* AerosolIndirect — crates/rscm-magicc/src/forcing/aerosol_indirect.rs       */
/* params: cloud_albedo_coefficient, reference_burden, sox_weight, oc_weight, */
/*  sox_pi, oc_pi, harmonize, harmonize_year, harmonize_target                */
/* ------------------------------------------------------------------------- */
static int aerosol_indirect_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)t0; (void)t1; (void)st;
    const double sox = orc_in_get(c, 0, 0), oc = orc_in_get(c, 1, 0);
    /* calculate_cloud_albedo — :152-188 */
    const double burden = p[2] * sox + p[3] * oc;
    const double burden_pi = p[2] * p[4] + p[3] * p[5];
    const double delta = burden - burden_pi;
    out[0] = (delta <= 0.0) ? 0.0 : p[0] * log(1.0 + delta / p[1]);
    return 0;
}
static const orc_def aerosol_indirect_defs[] = {
    {"Emissions|SOx", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Emissions|OC", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|Aerosol|Indirect", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
};
const orc_kind_info orc_kind_aerosol_indirect = {ORC_AEROSOL_INDIRECT, "AerosolIndirect", 3, aerosol_indirect_defs, 9, aerosol_indirect_solve, 0, NULL};
