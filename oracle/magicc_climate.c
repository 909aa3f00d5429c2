/*
 * magicc_climate.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restates ClimateUDEB of the reference (v0.5.0):
 *   crates/rscm-magicc/src/climate/udeb/mod.rs        (from_parameters, solve_impl, adjusted_ecs, ...)
 *   crates/rscm-magicc/src/climate/udeb/ocean_column.rs (step_hemisphere, layer_diffusivities, ...)
 *   crates/rscm-magicc/src/climate/lamcalc.rs          (LAMCALC iteration)
 *   crates/rscm-magicc/src/climate/state.rs            (ClimateUDEBState)
 *   crates/rscm-magicc/src/parameters/climate_udeb.rs  (derived quantities, CMIP5 profiles)
 *   crates/rscm-core/src/utils/linear_algebra.rs       (thomas_solve, invert_4x4)
 * Pinned loosely (1-5 %) by the MAGICC7 golden CSVs of tests/regression/data/ocean_udeb and
 * ghg_forcing/04,05 (tests/regression/test_ocean_udeb.py, test_ghg_forcing.py:700-735).
 */
#include "orc_internal.h"

#include <string.h>

#define UDEB_MAXL 64
#define UDEB_MAXH 4096

/* parameter block order = ClimateUDEBParameters declaration order, rf_regions_co2 expanded */
enum {
    U_NLAYERS, U_MLD, U_DZ, U_KAPPA, U_KAPPA_MIN, U_KAPPA_DKDT, U_W0, U_WVAR, U_WT_NH, U_WT_SH, U_ECS, U_RF2X, U_RLO,
    U_FB_Q, U_FB_CUMT, U_FB_PERIOD, U_KLO, U_KNS, U_AMP, U_NH_LAND, U_SH_LAND, U_DDA, U_TA_ALPHA, U_TA_GAMMA, U_PI_RATIO,
    U_LHC_ON, U_KLG, U_LHC_THICK, U_RFR0, U_RFR1, U_RFR2, U_RFR3, U_EFF_APPLY, U_EFF_CO2, U_PROFILE, U_STEPS, U_TMAX, U_NPARAM
};

#define DIFFUSIVITY_CM2S_TO_M2YR 3155.76
#define RHO_SEAWATER 1026.0
#define CP_SEAWATER 3985.0
#define SECONDS_PER_YEAR 31557600.0

#include "cmip5_profiles.inc"

typedef struct {
    int lam_ok;
    double lambda_ocean, lambda_land, co2_eff, qfrac[4];
    double af_top[UDEB_MAXL], af_bot[UDEB_MAXL], af_diff[UDEB_MAXL];
    double T[2][UDEB_MAXL], init[2][UDEB_MAXL];
    double w[2], land[2], ground[2], alpha_eff[2], hx[2];
    double t_polar;
    int nhist;
    double hist[UDEB_MAXH], dth[UDEB_MAXH];
} udeb_state;

/* invert_4x4 — rscm-core/src/utils/linear_algebra.rs:102-166 */
static int invert4(const double m[4][4], double inv[4][4])
{
    double aug[4][8];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) { aug[i][j] = m[i][j]; aug[i][j + 4] = 0.0; }
        aug[i][i + 4] = 1.0;
    }
    for (int col = 0; col < 4; ++col) {
        int max_row = col;
        double max_val = fabs(aug[col][col]);
        for (int row = col + 1; row < 4; ++row) {
            double v = fabs(aug[row][col]);
            if (v > max_val) { max_val = v; max_row = row; }
        }
        if (max_val < 1e-15) return 0;
        if (max_row != col)
            for (int j = 0; j < 8; ++j) { double t = aug[col][j]; aug[col][j] = aug[max_row][j]; aug[max_row][j] = t; }
        const double pivot = aug[col][col];
        for (int j = 0; j < 8; ++j) aug[col][j] /= pivot;
        for (int row = 0; row < 4; ++row) {
            if (row == col) continue;
            const double f = aug[row][col];
            for (int j = 0; j < 8; ++j) aug[row][j] -= f * aug[col][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i][j] = aug[i][j + 4];
    return 1;
}

static void box_fractions(const double *p, double *a)
{ /* global_box_fractions — parameters/climate_udeb.rs:366-372 */
    const double fgnl = p[U_NH_LAND] / 2.0, fgsl = p[U_SH_LAND] / 2.0;
    a[0] = 0.5 - fgnl; a[1] = fgnl; a[2] = 0.5 - fgsl; a[3] = fgsl;
}

static void compute_qfrac(const double *rf, const double *area, double *q)
{ /* lamcalc.rs:115-130 */
    double s = 0.0;
    for (int i = 0; i < 4; ++i) s += rf[i] * area[i];
    for (int i = 0; i < 4; ++i) q[i] = (fabs(s) <= 1e-15) ? 1.0 : rf[i] / s;
}

/* lamcalc — climate/lamcalc.rs:178-290; returns 0 when it does not converge */
static int lamcalc(const double *p, double ecs, double *lam_o_out, double *lam_l_out, double *eff_out)
{
    enum { MAXIT = 40 };
    const double q2x = p[U_RF2X], k_lo = p[U_KLO], k_ns = p[U_KNS], rlo = p[U_RLO], alpha = p[U_AMP];
    double area[4], qfrac[4];
    box_fractions(p, area);
    const double fgno = area[0], fgnl = area[1], fgso = area[2], fgsl = area[3];
    const double lam = q2x / ecs;
    const double fratio = (fgno + fgso) / (fgnl + fgsl);
    compute_qfrac(&p[U_RFR0], area, qfrac);
    double lamo[MAXIT + 2] = {0}, diff[MAXIT + 2] = {0};
    lamo[1] = lam;
    lamo[2] = lam + 0.7;
    double dlamo = 0.7;
    int iflag = 0;
    for (int i = 2; i <= MAXIT; ++i) {
        const double lam_l = lam + fratio * (lam - lamo[i]) / rlo;
        const double lam_o = lamo[i];
        const double mtx[4][4] = {/* build_coupling_matrix — lamcalc.rs:85-107 */
            {fgno * lam_o + k_lo * alpha + k_ns, -k_lo, -k_ns, 0.0},
            {-k_lo * alpha, fgnl * lam_l + k_lo, 0.0, 0.0},
            {-k_ns, 0.0, fgso * lam_o + k_lo * alpha + k_ns, -k_lo},
            {0.0, 0.0, -k_lo * alpha, fgsl * lam_l + k_lo}};
        double inv[4][4];
        if (!invert4(mtx, inv)) return 0;
        double temps[4];
        for (int r = 0; r < 4; ++r) {
            double s = 0.0;
            for (int c = 0; c < 4; ++c) s += inv[r][c] * area[c] * qfrac[c];
            temps[r] = q2x * s;
        }
        const double ocean_mean = (fgno * temps[0] + fgso * temps[2]) / (fgno + fgso);
        const double land_mean = (fgnl * temps[1] + fgsl * temps[3]) / (fgnl + fgsl);
        const double rlo_est = land_mean / ocean_mean;
        diff[i] = rlo - rlo_est;
        if (fabs(diff[i]) < 0.001) {
            /* calc_internal_efficacy — lamcalc.rs:142-170 */
            double rf_sum = 0.0;
            for (int c = 0; c < 4; ++c) rf_sum += p[U_RFR0 + c] * area[c];
            double eff = 1.0;
            if (fabs(rf_sum) > 1e-15) {
                double tg = 0.0;
                for (int r = 0; r < 4; ++r) tg += area[r] * temps[r];
                eff = tg / ecs;
            }
            *lam_o_out = lam_o; *lam_l_out = lam_l; *eff_out = eff;
            return 1;
        }
        if (diff[i] * diff[i - 1] < 0.0) iflag = 1;
        if (iflag == 0) {
            if (fabs(diff[i]) > fabs(diff[i - 1])) dlamo = -dlamo;
            lamo[i + 1] = lamo[i] + dlamo;
        } else if (diff[i] * diff[i - 1] < 0.0) {
            const double den = diff[i] - diff[i - 1];
            lamo[i + 1] = (fabs(den) < 1e-30) ? lamo[i] + dlamo : lamo[i] - diff[i] * (lamo[i] - lamo[i - 1]) / den;
        } else {
            const int i2 = i - 2;
            const double den = diff[i] - diff[i2];
            lamo[i + 1] = (fabs(den) < 1e-30) ? lamo[i] + dlamo : lamo[i] - diff[i] * (lamo[i] - lamo[i2]) / den;
        }
    }
    return 0;
}

static double ocean_area_at_depth(const double *p, double depth)
{ /* parameters/climate_udeb.rs: ocean_area_at_depth */
    static const double D[12] = {0.0, 200.0, 500.0, 1000.0, 1500.0, 2000.0, 2500.0, 3000.0, 3500.0, 4000.0, 4500.0, 5000.0};
    static const double A[12] = {1.0, 0.975, 0.95, 0.92, 0.91, 0.87, 0.81, 0.72, 0.55, 0.38, 0.18, 0.05};
    double hydro;
    if (depth <= D[0]) hydro = A[0];
    else if (depth >= D[11]) hydro = A[11];
    else {
        hydro = A[0];
        for (int i = 1; i < 12; ++i)
            if (depth <= D[i]) { hydro = A[i - 1] + (depth - D[i - 1]) / (D[i] - D[i - 1]) * (A[i] - A[i - 1]); break; }
    }
    return 1.0 + p[U_DDA] * (hydro - 1.0);
}

static double heat_capacity(double depth) { return RHO_SEAWATER * CP_SEAWATER * depth / SECONDS_PER_YEAR; }

/* ClimateUDEB::from_parameters + create_initial_state — udeb/mod.rs:81-145, state.rs:52-90 */
static void udeb_init(const double *p, void *vs)
{
    udeb_state *s = (udeb_state *)vs;
    memset(s, 0, sizeof *s);
    const int n = (int)p[U_NLAYERS];
    double area[4];
    box_fractions(p, area);
    s->lam_ok = (n >= 2 && n <= UDEB_MAXL && isfinite(p[U_EFF_CO2]) && p[U_EFF_CO2] > 0.0)
                    ? lamcalc(p, p[U_ECS], &s->lambda_ocean, &s->lambda_land, &s->co2_eff) : 0;
    compute_qfrac(&p[U_RFR0], area, s->qfrac);
    for (int l = 0; l < n && l < UDEB_MAXL; ++l) { /* compute_area_factors */
        double zt, zb;
        if (l == 0) { zt = 0.0; zb = p[U_MLD]; }
        else { zt = p[U_MLD] + ((double)l - 1.0) * p[U_DZ]; zb = zt + p[U_DZ]; }
        const double at = ocean_area_at_depth(p, zt), ab = ocean_area_at_depth(p, zb), avg = (at + ab) / 2.0;
        s->af_top[l] = at / avg; s->af_bot[l] = ab / avg; s->af_diff[l] = (at - ab) / avg;
    }
    for (int h = 0; h < 2; ++h) { /* initial_ocean_profile */
        for (int l = 0; l < n && l < UDEB_MAXL; ++l) {
            if (p[U_PROFILE] == 2.0) {
                const double *c = h == 0 ? CMIP5_NH : CMIP5_SH;
                s->init[h][l] = c[l < 50 ? l : 49];
            } else {
                const double kap = p[U_KAPPA] * DIFFUSIVITY_CM2S_TO_M2YR;
                s->init[h][l] = (l == 0) ? 17.2 : 1.0 + (17.2 - 1.0) * exp(-p[U_W0] * (((double)l - 1.0) * p[U_DZ] + 0.5 * p[U_DZ]) / kap);
            }
        }
        s->w[h] = p[U_W0];
        s->alpha_eff[h] = p[U_TA_ALPHA];
    }
    s->t_polar = 1.0;
}

static double sst_to_air(const double *p, double sst)
{ /* udeb/mod.rs:377-397 */
    const double alpha = p[U_TA_ALPHA], gamma = p[U_TA_GAMMA];
    const double t_star = (fabs(gamma) > 1e-15) ? -(alpha - 1.0) / (2.0 * gamma) : INFINITY;
    if (sst < t_star) return alpha * sst + gamma * sst * sst;
    return sst + (alpha * t_star + gamma * t_star * t_star - t_star);
}

static double land_temperature(const double *p, double ocean_temp, double land_forcing, double f_l, double lambda_land)
{ /* udeb/mod.rs:352-375 */
    const double num = land_forcing * f_l + p[U_KLO] * p[U_AMP] * ocean_temp;
    const double den = lambda_land * f_l + p[U_KLO];
    return fmin(num / den, p[U_TMAX]);
}

static void apply_efficacy(const double *p, const udeb_state *s, double erf, double co2_eff, double *f)
{ /* udeb/mod.rs:253-270 */
    double e = erf;
    const int mode = (int)p[U_EFF_APPLY];
    if (mode == 1) e = erf * p[U_EFF_CO2];
    else if (mode == 2 && isfinite(co2_eff) && co2_eff > 0.0) e = erf * p[U_EFF_CO2] / co2_eff;
    for (int i = 0; i < 4; ++i) f[i] = e * s->qfrac[i];
}

static double adjusted_ecs(const double *p, const udeb_state *s, double forcing)
{ /* udeb/mod.rs:302-350 */
    const double period = p[U_FB_PERIOD], cumt_2x = p[U_ECS] * period;
    double cum_t = 0.0;
    if (s->nhist > 0) {
        double rem = period, sum = 0.0;
        for (int i = s->nhist - 1; i >= 0; --i) {
            if (rem <= 0.0) break;
            const double dt = s->dth[i];
            if (dt <= rem) { sum += s->hist[i]; rem -= dt; }
            else { sum += s->hist[i] * (rem / dt); rem = 0.0; }
        }
        cum_t = sum;
    }
    const double cumt_factor = (fabs(cumt_2x) > 1e-15) ? 1.0 + p[U_FB_CUMT] * (cum_t - cumt_2x) / cumt_2x : 1.0;
    const double q_factor = 1.0 + p[U_FB_Q] * (fmax(forcing, 0.0) - p[U_RF2X]);
    return p[U_ECS] * cumt_factor * q_factor;
}

/* step_hemisphere — udeb/ocean_column.rs:54-241 (+ layer_diffusivities :23-52, thomas_solve) */
static double step_hemisphere(const double *p, udeb_state *s, int hemi, double forcing, double dt, double lam_o, double lam_l,
                              double hx, double ground_temp, double alpha_eff)
{
    const int n = (int)p[U_NLAYERS];
    const double dz = p[U_DZ], dz_mix = p[U_MLD], pi_ratio = p[U_PI_RATIO], w = s->w[hemi];
    double kap[UDEB_MAXL], a[UDEB_MAXL], b[UDEB_MAXL], c[UDEB_MAXL], d[UDEB_MAXL], cp[UDEB_MAXL], dp[UDEB_MAXL];
    double *T = s->T[hemi];
    {
        const double total_depth = dz_mix + ((double)n - 1.0) * dz;
        const double t_top = T[0], t_bottom = T[n - 1], kmin = p[U_KAPPA_MIN] * DIFFUSIVITY_CM2S_TO_M2YR;
        for (int l = 0; l < n - 1; ++l) {
            const double depth = dz_mix + (double)l * dz;
            const double rel = depth / total_depth;
            const double k = ((1.0 - rel) * p[U_KAPPA_DKDT] * (t_top - t_bottom) + p[U_KAPPA]) * DIFFUSIVITY_CM2S_TO_M2YR;
            kap[l] = fmax(k, kmin);
        }
    }
    const double c_mix = heat_capacity(dz_mix);
    for (int i = 0; i < n; ++i) a[i] = b[i] = c[i] = d[i] = 0.0;
    const double f_l = (hemi == 0 ? p[U_NH_LAND] : p[U_SH_LAND]) / 2.0;
    const double f_o = 0.5 - f_l;
    const double denominator = f_o * (p[U_KLO] + f_l * lam_l);
    const double term_feedback = alpha_eff / c_mix * (lam_o + lam_l * p[U_KLO] * p[U_AMP] * f_l / denominator);
    const double dz1 = dz / 2.0;
    const double term_diff = kap[0] / (dz_mix * dz1) * dt;
    const double term_upwell = w / dz_mix * dt;
    const double forcing_amp = 1.0 + p[U_KLO] * f_l / denominator;
    const double *aft = s->af_top, *afb = s->af_bot, *afd = s->af_diff;
    b[0] = 1.0 + term_feedback * dt * aft[0] + term_diff * afb[0] + term_upwell * pi_ratio * afb[0];
    c[0] = -(term_diff + term_upwell) * afb[0];
    d[0] = T[0] + (forcing * forcing_amp + hx) / c_mix * dt * aft[0];
    if (p[U_LHC_ON] != 0.0) d[0] -= p[U_KLG] * (s->land[hemi] - ground_temp) / (c_mix * f_o) * dt * aft[0];
    for (int i = 1; i < n - 1; ++i) {
        const double dz_up = (i == 1) ? dz1 : dz;
        const double tdu = kap[i - 1] / (dz * dz_up) * dt;
        const double tdd = kap[i] / (dz * dz) * dt;
        const double tul = w / dz * dt;
        a[i] = -tdu * aft[i];
        b[i] = 1.0 + tdu * aft[i] + tdd * afb[i] + tul * aft[i];
        c[i] = -(tdd + tul) * afb[i];
        d[i] = T[i] + pi_ratio * tul * T[0] * afd[i];
    }
    {
        const double tdu = kap[n - 2] / (dz * dz) * dt;
        const double tub = w / dz * dt;
        a[n - 1] = -tdu * aft[n - 1];
        b[n - 1] = 1.0 + (tdu + tub) * aft[n - 1];
        d[n - 1] = T[n - 1] + pi_ratio * tub * T[0] * aft[n - 1];
    }
    const double delta_w = w - p[U_W0];
    if (fabs(delta_w) > 1e-15) {
        const double *init = s->init[hemi];
        const double tp = s->t_polar;
        d[0] += dt / dz_mix * delta_w * (init[1] - tp) * afb[0];
        const double dtdz = dt / dz;
        for (int i = 1; i < n - 1; ++i) {
            d[i] += dtdz * delta_w * (init[i + 1] * afb[i] - init[i] * aft[i]);
            d[i] += dtdz * delta_w * tp * afd[i];
        }
        d[n - 1] += dtdz * delta_w * (tp - init[n - 1]) * aft[n - 1];
    }
    /* thomas_solve — linear_algebra.rs:41-79 */
    cp[0] = c[0] / b[0];
    dp[0] = d[0] / b[0];
    for (int i = 1; i < n; ++i) {
        const double den = b[i] - a[i] * cp[i - 1];
        if (i < n - 1) cp[i] = c[i] / den;
        dp[i] = (d[i] - a[i] * dp[i - 1]) / den;
    }
    double x_next = dp[n - 1];
    T[n - 1] = fmin(x_next, p[U_TMAX]);
    for (int i = n - 2; i >= 0; --i) {
        const double x = dp[i] - cp[i] * x_next; /* back substitution uses the uncapped solution */
        T[i] = fmin(x, p[U_TMAX]);
        x_next = x;
    }
    return T[0];
}

/* solve_impl — udeb/mod.rs:399-660.  inputs: [ERF (Input), Surface Temperature (State, FourBox)];
 * outputs: [Heat Uptake, Ocean Heat Content, Sea Surface Temperature, Surface Temperature[4]] */
static int udeb_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *vs)
{
    udeb_state *s = (udeb_state *)vs;
    if (!s->lam_ok) return 1; /* from_parameters failed: the member's model cannot be constructed */
    const int n = (int)p[U_NLAYERS], steps_n = (int)p[U_STEPS];
    const double erf_start = orc_in_start(c, 0, 0);
    int ok;
    double erf_end = orc_in_end(c, 0, 0, &ok);
    if (!ok) erf_end = erf_start;
    const double steps = (double)steps_n;
    const double prev[4] = {orc_in_start(c, 1, 0), orc_in_start(c, 1, 1), orc_in_start(c, 1, 2), orc_in_start(c, 1, 3)};
    if (s->T[0][0] == 0.0 && prev[0] != 0.0) { /* warm start */
        s->T[0][0] = prev[0]; s->T[1][0] = prev[2];
        s->land[0] = prev[1]; s->land[1] = prev[3];
        s->ground[0] = s->land[0]; s->ground[1] = s->land[1];
    }
    const double dt_year = t1 - t0, dt_sub = dt_year / steps;
    const double erf_mid = (erf_start + erf_end) / 2.0;
    const double aecs = adjusted_ecs(p, s, erf_mid);
    double lam_o = s->lambda_ocean, lam_l = s->lambda_land, co2_eff = s->co2_eff;
    if (fabs(aecs - p[U_ECS]) > 1e-10) {
        double lo, ll, ef;
        if (lamcalc(p, aecs, &lo, &ll, &ef)) { lam_o = lo; lam_l = ll; co2_eff = ef; }
    }
    double area[4];
    box_fractions(p, area);
    const double fgno = area[0], fgnl = area[1], fgso = area[2], fgsl = area[3];
    const double c_ground = (p[U_LHC_ON] != 0.0) ? heat_capacity(p[U_LHC_THICK]) : 0.0;
    const double a_nh = s->alpha_eff[0], a_sh = s->alpha_eff[1];
    for (int step = 1; step <= steps_n; ++step) {
        const double frac = (double)step / steps;
        const double erf = erf_start + frac * (erf_end - erf_start);
        double forcing[4];
        apply_efficacy(p, s, erf, co2_eff, forcing);
        if (p[U_LHC_ON] != 0.0) {
            const double fl[2] = {fgnl, fgsl};
            for (int h = 0; h < 2; ++h) {
                if (fl[h] < 1e-15) continue;
                const double flux = p[U_KLG] * (s->land[h] - s->ground[h]);
                s->ground[h] += flux / (fl[h] * c_ground) * dt_sub;
            }
        }
        const double g_nh = s->ground[0], g_sh = s->ground[1];
        const double sst_nh = step_hemisphere(p, s, 0, forcing[0], dt_sub, lam_o, lam_l, s->hx[0], g_nh, a_nh);
        const double sst_sh = step_hemisphere(p, s, 1, forcing[2], dt_sub, lam_o, lam_l, s->hx[1], g_sh, a_sh);
        const double air_nho = sst_to_air(p, sst_nh), air_sho = sst_to_air(p, sst_sh);
        s->land[0] = land_temperature(p, air_nho, forcing[1], fgnl, lam_l);
        s->land[1] = land_temperature(p, air_sho, forcing[3], fgsl, lam_l);
        if (fgno > 1e-15) s->hx[0] = p[U_KNS] / fgno * (air_sho - air_nho);
        if (fgso > 1e-15) s->hx[1] = p[U_KNS] / fgso * (air_nho - air_sho);
        const double gt = air_nho * fgno + s->land[0] * fgnl + air_sho * fgso + s->land[1] * fgsl;
        { /* update_upwelling — ocean_column.rs:243-259 */
            const double w0 = p[U_W0], fv = p[U_WVAR], wmin = w0 * (1.0 - fv);
            s->w[0] = fmax(w0 * (1.0 - fv * fmin(gt / p[U_WT_NH], 1.0)), wmin);
            s->w[1] = fmax(w0 * (1.0 - fv * fmin(gt / p[U_WT_SH], 1.0)), wmin);
        }
    }
    const double sst_nh = s->T[0][0], sst_sh = s->T[1][0];
    s->alpha_eff[0] = (fabs(sst_nh) < 1e-15) ? p[U_TA_ALPHA] : sst_to_air(p, sst_nh) / sst_nh;
    s->alpha_eff[1] = (fabs(sst_sh) < 1e-15) ? p[U_TA_ALPHA] : sst_to_air(p, sst_sh) / sst_sh;
    const double st[4] = {sst_to_air(p, sst_nh), s->land[0], sst_to_air(p, sst_sh), s->land[1]};
    const double gt = st[0] * fgno + st[1] * fgnl + st[2] * fgso + st[3] * fgsl;
    if (s->nhist < UDEB_MAXH) { s->hist[s->nhist] = gt * dt_year; s->dth[s->nhist] = dt_year; s->nhist++; }
    double f_end[4];
    apply_efficacy(p, s, erf_end, co2_eff, f_end);
    { /* calculate_heat_uptake — ocean_column.rs:262-284 */
        const double lams[4] = {lam_o, lam_l, lam_o, lam_l};
        double q = 0.0, fb = 0.0;
        for (int i = 0; i < 4; ++i) { q += area[i] * f_end[i]; fb += area[i] * lams[i] * st[i]; }
        out[0] = q - fb;
    }
    { /* calculate_ocean_heat_content — ocean_column.rs:286-306 */
        const double rho_c = RHO_SEAWATER * CP_SEAWATER;
        double total = 0.0;
        for (int h = 0; h < 2; ++h) {
            total += rho_c * p[U_MLD] * s->T[h][0];
            for (int l = 1; l < n; ++l) total += rho_c * p[U_DZ] * s->T[h][l];
        }
        out[1] = total / 2.0;
    }
    out[2] = (sst_nh + sst_sh) / 2.0;
    for (int i = 0; i < 4; ++i) out[3 + i] = st[i];
    return 0;
}

static const orc_def udeb_defs[] = {
    {"Effective Radiative Forcing", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Heat Uptake", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Ocean Heat Content", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Sea Surface Temperature", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_STATE, ORC_GRID_FOUR_BOX},
};
const orc_kind_info orc_kind_climate_udeb = {ORC_CLIMATE_UDEB, "ClimateUDEB", 5, udeb_defs, U_NPARAM, udeb_solve,
                                             sizeof(udeb_state), udeb_init};
