// climate_udeb.cuh — MAGICC upwelling-diffusion energy-balance model (ClimateUDEB) on the device.
//
// Reference: crates/rscm-magicc/src/climate/udeb/mod.rs (solve_impl :399-660, adjusted_ecs :302-350,
// apply_efficacy_and_qfrac :253-270, sst_to_air_temperature :377-397, calculate_land_temperature :352-375),
// udeb/ocean_column.rs (step_hemisphere :54-241, layer_diffusivities :23-52, update_upwelling :243-259,
// heat uptake / ocean heat content :262-306), climate/lamcalc.rs (:85-290), climate/state.rs,
// rscm-core/src/utils/linear_algebra.rs (thomas_solve :41-79, invert_4x4 :102-166).
//
// One thread = one member.  Where the state lives:
//   * S[20]      (registers): LAMCALC result at the member's base ECS, upwelling rates, land / ground
//                temperatures, alpha_eff, inter-hemispheric exchange, history length;
//   * cx.sm      (shared memory, [150][BLOCK] per CTA, conflict-free): the two 50-layer ocean columns and
//                the Thomas sweep's c' array — the tridiagonal rows are built on the fly, d' overwrites T;
//   * cx.scratch (global, member-interleaved [T][runs]): the T*dt history of the cumulative-temperature
//                feedback, summed newest-to-oldest in the reference's order;
//   * cx.ctab    (shared memory, per graph): area factors af_top/af_bottom/af_diff and the initial ocean
//                profile of both hemispheres — they depend only on geometry parameters, which are
//                per-graph (not bindable per member), so the host computes them once.
#pragma once

namespace rscm_dev {

enum {
    U_NLAYERS, U_MLD, U_DZ, U_KAPPA, U_KAPPA_MIN, U_KAPPA_DKDT, U_W0, U_WVAR, U_WT_NH, U_WT_SH, U_ECS, U_RF2X, U_RLO,
    U_FB_Q, U_FB_CUMT, U_FB_PERIOD, U_KLO, U_KNS, U_AMP, U_NH_LAND, U_SH_LAND, U_DDA, U_TA_ALPHA, U_TA_GAMMA, U_PI_RATIO,
    U_LHC_ON, U_KLG, U_LHC_THICK, U_RFR0, U_RFR1, U_RFR2, U_RFR3, U_EFF_APPLY, U_EFF_CO2, U_PROFILE, U_STEPS, U_TMAX, U_NPARAM
};
enum { US_OK, US_LAMO, US_LAML, US_EFF, US_QF0, US_QF1, US_QF2, US_QF3, US_W0, US_W1, US_LAND0, US_LAND1, US_GR0, US_GR1,
       US_AE0, US_AE1, US_HX0, US_HX1, US_NHIST, US_N };

constexpr int UDEB_MAXL = 50;

template <class R> __device__ __forceinline__ R r_min(R a, R b) { return (b < a || a != a) ? b : a; } // f64::min: NaN-ignoring
template <class R> __device__ __forceinline__ R r_max(R a, R b) { return (b > a || a != a) ? b : a; }

// invert_4x4 — Gauss-Jordan with partial pivoting
template <class R> __device__ inline bool udeb_invert4(const R (&m)[4][4], R (&inv)[4][4])
{
    R aug[4][8];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) { aug[i][j] = m[i][j]; aug[i][j + 4] = R(0); }
        aug[i][i + 4] = R(1);
    }
    for (int col = 0; col < 4; ++col) {
        int max_row = col;
        R max_val = r_abs(aug[col][col]);
        for (int row = col + 1; row < 4; ++row) {
            const R v = r_abs(aug[row][col]);
            if (v > max_val) { max_val = v; max_row = row; }
        }
        if (max_val < R(1e-15)) return false;
        if (max_row != col)
            for (int j = 0; j < 8; ++j) { const R t = aug[col][j]; aug[col][j] = aug[max_row][j]; aug[max_row][j] = t; }
        const R pivot = aug[col][col];
        for (int j = 0; j < 8; ++j) aug[col][j] /= pivot;
        for (int row = 0; row < 4; ++row) {
            if (row == col) continue;
            const R f = aug[row][col];
            for (int j = 0; j < 8; ++j) aug[row][j] -= f * aug[col][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i][j] = aug[i][j + 4];
    return true;
}

template <class R> __device__ __forceinline__ void udeb_fractions(const R *P, R (&a)[4])
{
    const R fgnl = P[U_NH_LAND] / R(2), fgsl = P[U_SH_LAND] / R(2);
    a[0] = R(0.5) - fgnl; a[1] = fgnl; a[2] = R(0.5) - fgsl; a[3] = fgsl;
}

template <class R> __device__ __forceinline__ void udeb_qfrac(const R *P, const R (&area)[4], R (&q)[4])
{
    R s = R(0);
    for (int i = 0; i < 4; ++i) s += P[U_RFR0 + i] * area[i];
    for (int i = 0; i < 4; ++i) q[i] = (r_abs(s) <= R(1e-15)) ? R(1) : P[U_RFR0 + i] / s;
}

// lamcalc — hybrid step / secant iteration on lambda_ocean, tolerance 1e-3 on the land/ocean warming ratio
template <class R> __device__ inline bool udeb_lamcalc(const R *P, R ecs, R &lam_o_out, R &lam_l_out, R &eff_out)
{
    constexpr int MAXIT = 40;
    const R q2x = P[U_RF2X], k_lo = P[U_KLO], k_ns = P[U_KNS], rlo = P[U_RLO], alpha = P[U_AMP];
    R area[4], qfrac[4];
    udeb_fractions(P, area);
    udeb_qfrac(P, area, qfrac);
    const R fgno = area[0], fgnl = area[1], fgso = area[2], fgsl = area[3];
    const R lam = q2x / ecs;
    const R fratio = (fgno + fgso) / (fgnl + fgsl);
    R lamo[MAXIT + 2], diff[MAXIT + 2];
    for (int i = 0; i < MAXIT + 2; ++i) { lamo[i] = R(0); diff[i] = R(0); }
    lamo[1] = lam;
    lamo[2] = lam + R(0.7);
    R dlamo = R(0.7);
    int iflag = 0;
    for (int i = 2; i <= MAXIT; ++i) {
        const R lam_l = lam + fratio * (lam - lamo[i]) / rlo;
        const R lam_o = lamo[i];
        const R mtx[4][4] = {{fgno * lam_o + k_lo * alpha + k_ns, -k_lo, -k_ns, R(0)},
                             {-k_lo * alpha, fgnl * lam_l + k_lo, R(0), R(0)},
                             {-k_ns, R(0), fgso * lam_o + k_lo * alpha + k_ns, -k_lo},
                             {R(0), R(0), -k_lo * alpha, fgsl * lam_l + k_lo}};
        R inv[4][4];
        if (!udeb_invert4(mtx, inv)) return false;
        R temps[4];
        for (int r = 0; r < 4; ++r) {
            R s = R(0);
            for (int c = 0; c < 4; ++c) s += inv[r][c] * area[c] * qfrac[c];
            temps[r] = q2x * s;
        }
        const R ocean_mean = (fgno * temps[0] + fgso * temps[2]) / (fgno + fgso);
        const R land_mean = (fgnl * temps[1] + fgsl * temps[3]) / (fgnl + fgsl);
        diff[i] = rlo - land_mean / ocean_mean;
        if (r_abs(diff[i]) < R(0.001)) {
            R rf_sum = R(0);
            for (int c = 0; c < 4; ++c) rf_sum += P[U_RFR0 + c] * area[c];
            R eff = R(1);
            if (r_abs(rf_sum) > R(1e-15)) {
                R tg = R(0);
                for (int r = 0; r < 4; ++r) tg += area[r] * temps[r];
                eff = tg / ecs;
            }
            lam_o_out = lam_o; lam_l_out = lam_l; eff_out = eff;
            return true;
        }
        if (diff[i] * diff[i - 1] < R(0)) iflag = 1;
        if (iflag == 0) {
            if (r_abs(diff[i]) > r_abs(diff[i - 1])) dlamo = -dlamo;
            lamo[i + 1] = lamo[i] + dlamo;
        } else if (diff[i] * diff[i - 1] < R(0)) {
            const R den = diff[i] - diff[i - 1];
            lamo[i + 1] = (r_abs(den) < R(1e-30)) ? lamo[i] + dlamo : lamo[i] - diff[i] * (lamo[i] - lamo[i - 1]) / den;
        } else {
            const R den = diff[i] - diff[i - 2];
            lamo[i + 1] = (r_abs(den) < R(1e-30)) ? lamo[i] + dlamo : lamo[i] - diff[i] * (lamo[i] - lamo[i - 2]) / den;
        }
    }
    return false;
}

template <class R> __device__ __forceinline__ R udeb_heat_capacity(R depth)
{
    return R(1026.0) * R(3985.0) * depth / R(31557600.0);
}

template <class R> __device__ __forceinline__ R udeb_sst_to_air(const R *P, R sst)
{
    const R alpha = P[U_TA_ALPHA], gamma = P[U_TA_GAMMA];
    if (r_abs(gamma) > R(1e-15)) {
        const R t_star = -(alpha - R(1)) / (R(2) * gamma);
        if (!(sst < t_star)) return sst + (alpha * t_star + gamma * t_star * t_star - t_star);
    }
    return alpha * sst + gamma * sst * sst;
}

template <class R> __device__ __forceinline__ R udeb_land_temperature(const R *P, R ocean_temp, R land_forcing, R f_l, R lambda_land)
{
    const R num = land_forcing * f_l + P[U_KLO] * P[U_AMP] * ocean_temp;
    const R den = lambda_land * f_l + P[U_KLO];
    return r_min(num / den, P[U_TMAX]);
}

template <class R> __device__ __forceinline__ void udeb_apply_efficacy(const R *P, const R *S, R erf, R co2_eff, R (&f)[4])
{
    R e = erf;
    const int mode = static_cast<int>(P[U_EFF_APPLY]);
    if (mode == 1) e = erf * P[U_EFF_CO2];
    else if (mode == 2 && (co2_eff - co2_eff) == R(0) && co2_eff > R(0)) e = erf * P[U_EFF_CO2] / co2_eff; // finite and > 0
    for (int i = 0; i < 4; ++i) f[i] = e * S[US_QF0 + i];
}

constexpr int CLIMATE_UDEB_NP = U_NPARAM;
constexpr int CLIMATE_UDEB_ND = 1;

template <class R> __device__ __forceinline__ void climate_udeb_prepare(const R *, R *D) { D[0] = R(0); }

// from_parameters + create_initial_state
template <class R> __device__ inline void climate_udeb_init_state(const R *P, const R *, R *S, const StepCtx<R> &cx, NodeRef nr)
{
    const int n = static_cast<int>(P[U_NLAYERS]);
    R area[4], q[4];
    udeb_fractions(P, area);
    udeb_qfrac(P, area, q);
    R lo = R(0), ll = R(0), ef = R(1);
    const bool ok = (P[U_EFF_CO2] > R(0)) && (P[U_EFF_CO2] == P[U_EFF_CO2]) && udeb_lamcalc(P, P[U_ECS], lo, ll, ef);
    S[US_OK] = ok ? R(1) : R(0);
    S[US_LAMO] = lo; S[US_LAML] = ll; S[US_EFF] = ef;
    for (int i = 0; i < 4; ++i) S[US_QF0 + i] = q[i];
    S[US_W0] = S[US_W1] = P[U_W0];
    S[US_LAND0] = S[US_LAND1] = S[US_GR0] = S[US_GR1] = R(0);
    S[US_AE0] = S[US_AE1] = P[U_TA_ALPHA];
    S[US_HX0] = S[US_HX1] = R(0);
    S[US_NHIST] = R(0);
    R *T = cx.sm + nr.sm * BLOCK;
    for (int i = 0; i < 2 * n; ++i) T[i * BLOCK] = R(0);
}

// step_hemisphere: build the tridiagonal rows on the fly, Thomas forward sweep (c' to shared memory, d' over T),
// back substitution, temperature cap.  T = this thread's column (stride BLOCK), cp = this thread's c' array.
template <class R>
__device__ inline R udeb_step_hemisphere(const R *P, const R *S, const double *ctab, R *T, R *cp, int hemi, R forcing, R dt, R lam_o,
                                         R lam_l, R hx, R ground_temp, R alpha_eff)
{
    const int n = static_cast<int>(P[U_NLAYERS]);
    const double *aft = ctab, *afb = ctab + n, *afd = ctab + 2 * n, *init = ctab + (3 + hemi) * n;
    const R dz = P[U_DZ], dz_mix = P[U_MLD], pi_ratio = P[U_PI_RATIO], w = S[US_W0 + hemi];
    const R conv = R(3155.76); // DIFFUSIVITY_CM2S_TO_M2YR
    const R total_depth = dz_mix + (R(n) - R(1)) * dz;
    const R t_top = T[0], t_bottom = T[(n - 1) * BLOCK];
    const R kmin = P[U_KAPPA_MIN] * conv;
    const R dk = P[U_KAPPA_DKDT] * (t_top - t_bottom);
    // kappa at the boundary below layer l
    auto kappa = [&](int l) -> R {
        const R rel = (dz_mix + R(l) * dz) / total_depth;
        return r_max(((R(1) - rel) * dk + P[U_KAPPA]) * conv, kmin);
    };
    const R c_mix = udeb_heat_capacity(dz_mix);
    const R f_l = (hemi == 0 ? P[U_NH_LAND] : P[U_SH_LAND]) / R(2);
    const R f_o = R(0.5) - f_l;
    const R denominator = f_o * (P[U_KLO] + f_l * lam_l);
    const R term_feedback = alpha_eff / c_mix * (lam_o + lam_l * P[U_KLO] * P[U_AMP] * f_l / denominator);
    const R dz1 = dz / R(2);
    const R forcing_amp = R(1) + P[U_KLO] * f_l / denominator;
    const R delta_w = w - P[U_W0];
    const bool dw = r_abs(delta_w) > R(1e-15);
    const R tp = R(1); // polar_sinking_temp (state.rs default)
    const R dtdz = dt / dz;
    const R t0 = t_top; // mixed-layer temperature before the solve (entrainment terms)

    // row 0
    R k_prev = kappa(0);
    R cprev, dprev;
    {
        const R term_diff = k_prev / (dz_mix * dz1) * dt;
        const R term_upwell = w / dz_mix * dt;
        const R b0 = R(1) + term_feedback * dt * R(aft[0]) + term_diff * R(afb[0]) + term_upwell * pi_ratio * R(afb[0]);
        const R c0 = -(term_diff + term_upwell) * R(afb[0]);
        R d0 = t0 + (forcing * forcing_amp + hx) / c_mix * dt * R(aft[0]);
        if (P[U_LHC_ON] != R(0)) d0 -= P[U_KLG] * (S[US_LAND0 + hemi] - ground_temp) / (c_mix * f_o) * dt * R(aft[0]);
        if (dw) d0 += dt / dz_mix * delta_w * (R(init[1]) - tp) * R(afb[0]);
        cprev = c0 / b0;
        dprev = d0 / b0;
        cp[0] = cprev;
        T[0] = dprev;
    }
    const R tul = w / dz * dt;
    for (int i = 1; i < n - 1; ++i) {
        const R dz_up = (i == 1) ? dz1 : dz;
        const R k_i = kappa(i);
        const R tdu = k_prev / (dz * dz_up) * dt;
        const R tdd = k_i / (dz * dz) * dt;
        const R at = R(aft[i]), ab = R(afb[i]), ad = R(afd[i]);
        const R ai = -tdu * at;
        const R bi = R(1) + tdu * at + tdd * ab + tul * at;
        const R ci = -(tdd + tul) * ab;
        R di = T[i * BLOCK] + pi_ratio * tul * t0 * ad;
        if (dw) {
            di += dtdz * delta_w * (R(init[i + 1]) * ab - R(init[i]) * at);
            di += dtdz * delta_w * tp * ad;
        }
        const R den = bi - ai * cprev;
        cprev = ci / den;
        dprev = (di - ai * dprev) / den;
        cp[i * BLOCK] = cprev;
        T[i * BLOCK] = dprev;
        k_prev = k_i;
    }
    {
        const int i = n - 1;
        const R tdu = k_prev / (dz * dz) * dt;
        const R at = R(aft[i]);
        const R ai = -tdu * at;
        const R bi = R(1) + (tdu + tul) * at;
        R di = T[i * BLOCK] + pi_ratio * tul * t0 * at;
        if (dw) di += dtdz * delta_w * (tp - R(init[i])) * at;
        const R den = bi - ai * cprev;
        dprev = (di - ai * dprev) / den;
    }
    // back substitution on the uncapped solution; the cap applies to what is stored
    const R tmax = P[U_TMAX];
    R x = dprev;
    T[(n - 1) * BLOCK] = r_min(x, tmax);
    for (int i = n - 2; i >= 0; --i) {
        x = T[i * BLOCK] - cp[i * BLOCK] * x;
        T[i * BLOCK] = r_min(x, tmax);
    }
    return T[0];
}

// in: [ERF at_start, ERF at_end, Surface Temperature[4] at_start]
// out: [Heat Uptake, Ocean Heat Content, Sea Surface Temperature, Surface Temperature[4]]
template <class R>
__device__ inline bool climate_udeb_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef nr)
{
    if (S[US_OK] == R(0)) return false; // from_parameters failed for this member (LAMCALC did not converge)
    const int n = static_cast<int>(P[U_NLAYERS]), steps_n = static_cast<int>(P[U_STEPS]);
    R *T0 = cx.sm + nr.sm * BLOCK, *T1 = T0 + n * BLOCK, *cp = T0 + 2 * n * BLOCK;
    const double *ctab = cx.ctab + nr.ctab;
    const R erf_start = in[0], erf_end = in[1];
    if (T0[0] == R(0) && in[2] != R(0)) { // warm start from non-zero initial surface temperatures
        T0[0] = in[2]; T1[0] = in[4];
        S[US_LAND0] = in[3]; S[US_LAND1] = in[5];
        S[US_GR0] = S[US_LAND0]; S[US_GR1] = S[US_LAND1];
    }
    const R dt_year = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    const R steps = R(steps_n);
    const R dt_sub = dt_year / steps;

    // adjusted_ecs: sum of T*dt over the last `period` years, newest to oldest
    const int nhist = static_cast<int>(S[US_NHIST]);
    R cum_t = R(0);
    {
        R rem = P[U_FB_PERIOD], sum = R(0);
        for (int i = nhist - 1; i >= 0; --i) {
            if (rem <= R(0)) break;
            const R dt = R(cx.bounds[i + 1] - cx.bounds[i]);
            const R h = R(cx.scratch[static_cast<long long>(nr.scr + i) * cx.runs]);
            if (dt <= rem) { sum += h; rem -= dt; }
            else { sum += h * (rem / dt); rem = R(0); }
        }
        cum_t = sum;
    }
    const R cumt_2x = P[U_ECS] * P[U_FB_PERIOD];
    const R cumt_factor = (r_abs(cumt_2x) > R(1e-15)) ? R(1) + P[U_FB_CUMT] * (cum_t - cumt_2x) / cumt_2x : R(1);
    const R erf_mid = (erf_start + erf_end) / R(2);
    const R q_factor = R(1) + P[U_FB_Q] * (r_max(erf_mid, R(0)) - P[U_RF2X]);
    const R aecs = P[U_ECS] * cumt_factor * q_factor;

    R lam_o = S[US_LAMO], lam_l = S[US_LAML], co2_eff = S[US_EFF];
    if (r_abs(aecs - P[U_ECS]) > R(1e-10)) {
        R lo, ll, ef;
        if (udeb_lamcalc(P, aecs, lo, ll, ef)) { lam_o = lo; lam_l = ll; co2_eff = ef; }
    }
    R area[4];
    udeb_fractions(P, area);
    const R fgno = area[0], fgnl = area[1], fgso = area[2], fgsl = area[3];
    const bool lhc = P[U_LHC_ON] != R(0);
    const R c_ground = lhc ? udeb_heat_capacity(P[U_LHC_THICK]) : R(0);
    const R a_nh = S[US_AE0], a_sh = S[US_AE1];
    for (int step = 1; step <= steps_n; ++step) {
        const R frac = R(step) / steps;
        const R erf = erf_start + frac * (erf_end - erf_start);
        R forcing[4];
        udeb_apply_efficacy(P, S, erf, co2_eff, forcing);
        if (lhc) {
            if (!(fgnl < R(1e-15))) S[US_GR0] += P[U_KLG] * (S[US_LAND0] - S[US_GR0]) / (fgnl * c_ground) * dt_sub;
            if (!(fgsl < R(1e-15))) S[US_GR1] += P[U_KLG] * (S[US_LAND1] - S[US_GR1]) / (fgsl * c_ground) * dt_sub;
        }
        const R sst_nh = udeb_step_hemisphere(P, S, ctab, T0, cp, 0, forcing[0], dt_sub, lam_o, lam_l, S[US_HX0], S[US_GR0], a_nh);
        const R sst_sh = udeb_step_hemisphere(P, S, ctab, T1, cp, 1, forcing[2], dt_sub, lam_o, lam_l, S[US_HX1], S[US_GR1], a_sh);
        const R air_nho = udeb_sst_to_air(P, sst_nh), air_sho = udeb_sst_to_air(P, sst_sh);
        S[US_LAND0] = udeb_land_temperature(P, air_nho, forcing[1], fgnl, lam_l);
        S[US_LAND1] = udeb_land_temperature(P, air_sho, forcing[3], fgsl, lam_l);
        if (fgno > R(1e-15)) S[US_HX0] = P[U_KNS] / fgno * (air_sho - air_nho);
        if (fgso > R(1e-15)) S[US_HX1] = P[U_KNS] / fgso * (air_nho - air_sho);
        const R gt = air_nho * fgno + S[US_LAND0] * fgnl + air_sho * fgso + S[US_LAND1] * fgsl;
        const R w0 = P[U_W0], fv = P[U_WVAR], wmin = w0 * (R(1) - fv);
        S[US_W0] = r_max(w0 * (R(1) - fv * r_min(gt / P[U_WT_NH], R(1))), wmin);
        S[US_W1] = r_max(w0 * (R(1) - fv * r_min(gt / P[U_WT_SH], R(1))), wmin);
    }
    const R sst_nh = T0[0], sst_sh = T1[0];
    S[US_AE0] = (r_abs(sst_nh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(P, sst_nh) / sst_nh;
    S[US_AE1] = (r_abs(sst_sh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(P, sst_sh) / sst_sh;
    const R st[4] = {udeb_sst_to_air(P, sst_nh), S[US_LAND0], udeb_sst_to_air(P, sst_sh), S[US_LAND1]};
    const R gt = st[0] * fgno + st[1] * fgnl + st[2] * fgso + st[3] * fgsl;
    cx.scratch[static_cast<long long>(nr.scr + nhist) * cx.runs] = static_cast<double>(gt * dt_year);
    S[US_NHIST] = R(nhist + 1);
    R f_end[4];
    udeb_apply_efficacy(P, S, erf_end, co2_eff, f_end);
    {
        const R lams[4] = {lam_o, lam_l, lam_o, lam_l};
        R q = R(0), fb = R(0);
        for (int i = 0; i < 4; ++i) { q += area[i] * f_end[i]; fb += area[i] * lams[i] * st[i]; }
        out[0] = q - fb;
    }
    {
        const R rho_c = R(1026.0) * R(3985.0);
        R total = R(0);
        for (int h = 0; h < 2; ++h) {
            const R *T = h == 0 ? T0 : T1;
            total += rho_c * P[U_MLD] * T[0];
            for (int l = 1; l < n; ++l) total += rho_c * P[U_DZ] * T[l * BLOCK];
        }
        out[1] = total / R(2);
    }
    out[2] = (sst_nh + sst_sh) / R(2);
    for (int i = 0; i < 4; ++i) out[3 + i] = st[i];
    return true;
}

} // namespace rscm_dev
