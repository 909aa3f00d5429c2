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
//   * cx.sm      (shared memory, [50][2][BLOCK] per CTA, conflict-free): the two 50-layer ocean columns — the
//                tridiagonal rows are built on the fly, d' overwrites T (fp32: also the c' columns);
//   * cx.scratch (global, member-interleaved [100 + T][runs]): the c' columns of the Thomas sweeps (fp64), then
//                the T*dt history of the cumulative-temperature feedback, summed newest-to-oldest in the
//                reference's order;
//   * cx.ctab    (shared memory, per graph): area factors af_top/af_bottom/af_diff, the entrainment
//                combinations of the initial ocean profile of both hemispheres and the relative-depth factor
//                of the diffusivity profile — they depend only on geometry parameters, which are
//                per-graph (not bindable per member), so the host computes them once (graph.cpp).
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
constexpr int UDEB_ROW = 2 * BLOCK; // per-thread shared-memory scratch: values per layer and CTA {T_nh, T_sh}
constexpr int UDEB_CT = 8;          // constant table: values per layer

// f64::min / f64::max: NaN-ignoring, like fmin / fmax
__device__ __forceinline__ double r_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float r_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }

// Reciprocal for the Thomas sweep: the pivots of the diagonally dominant ocean-column system are >= 1, never
// subnormal or infinite, so the special-case slow path of 1/x (a divergent call that also keeps the compiler from
// interleaving the two hemispheres' recurrences) is dropped: MUFU seed (~2^-20) + two Newton steps (<= 1 ulp; a NaN
// pivot stays NaN).
__device__ __forceinline__ double r_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float r_rcp(float x) { return __frcp_rn(x); }

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

// lamcalc — hybrid step / secant iteration on lambda_ocean, tolerance 1e-3 on the land/ocean warming ratio.
// The reference inverts the 4x4 box-coupling matrix by Gauss-Jordan (linear_algebra.rs:102-166); its sparsity
//   [ A0      -k_lo   -k_ns    0    ]      rows 1 and 3 couple each land box to its own ocean box only,
//   [ -k_lo*a  B1      0       0    ]      so eliminating them leaves a 2x2 system in the two ocean boxes,
//   [ -k_ns    0       A2     -k_lo ]      solved here by Cramer's rule (same solution up to rounding; the
//   [ 0        0      -k_lo*a  B3   ]      singular-pivot failure becomes a vanishing B1, B3 or determinant).
// Only the last three iterates of the reference's lamo[]/diff[] arrays are ever read: kept as scalars.
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
    const R f0 = q2x * area[0] * qfrac[0], f1 = q2x * area[1] * qfrac[1], f2 = q2x * area[2] * qfrac[2], f3 = q2x * area[3] * qfrac[3];
    const R kla = k_lo * alpha, inv_o = R(1) / (fgno + fgso), inv_l = R(1) / (fgnl + fgsl);
    R l0 = R(0), l1 = lam, l2 = lam + R(0.7); // lamo[i-2], lamo[i-1], lamo[i]
    R d0 = R(0), d1 = R(0);                    // diff[i-2], diff[i-1]
    R dlamo = R(0.7);
    bool bracketed = false;
    for (int i = 2; i <= MAXIT; ++i) {
        const R lam_o = l2;
        const R lam_l = lam + fratio * (lam - lam_o) / rlo;
        const R B1 = fgnl * lam_l + k_lo, B3 = fgsl * lam_l + k_lo;
        if (r_abs(B1) < R(1e-15) || r_abs(B3) < R(1e-15)) return false;
        const R r1 = R(1) / B1, r3 = R(1) / B3;
        const R a00 = fgno * lam_o + kla + k_ns - k_lo * kla * r1;
        const R a22 = fgso * lam_o + kla + k_ns - k_lo * kla * r3;
        const R g0 = f0 + k_lo * f1 * r1, g2 = f2 + k_lo * f3 * r3;
        const R det = a00 * a22 - k_ns * k_ns;
        if (r_abs(det) < R(1e-15)) return false;
        const R rd = R(1) / det;
        const R t0 = (g0 * a22 + k_ns * g2) * rd, t2 = (a00 * g2 + k_ns * g0) * rd;
        const R t1 = (f1 + kla * t0) * r1, t3 = (f3 + kla * t2) * r3;
        const R ocean_mean = (fgno * t0 + fgso * t2) * inv_o;
        const R land_mean = (fgnl * t1 + fgsl * t3) * inv_l;
        const R d2 = rlo - land_mean / ocean_mean;
        if (r_abs(d2) < R(0.001)) {
            R rf_sum = R(0);
            for (int c = 0; c < 4; ++c) rf_sum += P[U_RFR0 + c] * area[c];
            R eff = R(1);
            if (r_abs(rf_sum) > R(1e-15)) eff = (area[0] * t0 + area[1] * t1 + area[2] * t2 + area[3] * t3) / ecs;
            lam_o_out = lam_o; lam_l_out = lam_l; eff_out = eff;
            return true;
        }
        const bool flip = d2 * d1 < R(0);
        bracketed = bracketed || flip;
        R next;
        if (!bracketed) {
            if (r_abs(d2) > r_abs(d1)) dlamo = -dlamo;
            next = l2 + dlamo;
        } else {
            const R dref = flip ? d1 : d0, lref = flip ? l1 : l0;
            const R den = d2 - dref;
            next = (r_abs(den) < R(1e-30)) ? l2 + dlamo : l2 - d2 * (l2 - lref) / den;
        }
        l0 = l1; l1 = l2; l2 = next;
        d0 = d1; d1 = d2;
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
    R *col = cx.sm + nr.sm * BLOCK * (8 / static_cast<int>(sizeof(R)));
    for (int i = 0; i < n; ++i) col[i * UDEB_ROW] = col[i * UDEB_ROW + BLOCK] = R(0);
}

// Constants of one model year for step_hemisphere (everything that does not change between sub-steps).
template <class R> struct UdebYear {
    R cA, cA1, cM, dt_mix, dtdz, dt_cmix, kc, kmin, dkdt_c, pi_ratio, tmax;
    R tfb_dt[2], famp[2], lhc_c[2]; // per hemisphere
    bool lhc;
};

// step_hemisphere for both hemispheres at once, each solved by two-ended elimination.
//
// The reference runs thomas_solve top to bottom (linear_algebra.rs:41-79): one 50-long recurrence with a
// reciprocal in the loop-carried chain, i.e. pure latency for a thread.  The same tridiagonal system is solved here
// by eliminating from both ends towards the middle ("burn at both ends"): rows 0..k the usual way
// (x_i = d'_i - c'_i x_{i+1}), rows n-1..k+1 mirrored (x_i = d"_i - a'_i x_{i-1}), a 2x2 solve where they meet,
// and substitution outwards in both directions.  Same solution up to rounding (the system is strictly diagonally
// dominant), but with the two hemispheres that makes four independent recurrences per thread instead of one.
// Rows are built on the fly; -c' (or -a') goes to the per-thread c' column and d' over T; one reciprocal per row; the
// temperature cap applies to what is stored, the substitution carries the uncapped value (ocean_column.rs:226-238).
// col  = this thread's ocean columns, layer-major: layer i holds {T_nh, T_sh} at col[(2*i + h) * BLOCK];
// cp   = this thread's c' columns, layer i at cp[(2*i + h) * cs] (where they live: see climate_udeb_solve).
// ctab = per layer {af_top, af_bottom, af_diff, omr_i, g_nh, g_sh, omr_{i-1}, 0} (graph.cpp udeb_const_table).
template <class R> struct UdebRow { R at, ab, ad, om, g[2], omu; };

template <class R> __device__ __forceinline__ UdebRow<R> udeb_row(const double *ct)
{
    const double2 c01 = *reinterpret_cast<const double2 *>(ct), c23 = *reinterpret_cast<const double2 *>(ct + 2),
                  c45 = *reinterpret_cast<const double2 *>(ct + 4); // 64-byte rows: 16-byte aligned
    UdebRow<R> r;
    r.at = R(c01.x); r.ab = R(c01.y); r.ad = R(c23.x); r.om = R(c23.y); r.g[0] = R(c45.x); r.g[1] = R(c45.y); r.omu = R(ct[6]);
    return r;
}

template <class R>
__device__ __forceinline__ void udeb_step_both(const UdebYear<R> &y, const R *P, R *S, const double *ctab, int n, R *col, R *cp,
                                               long long cs, R forcing_nh, R forcing_sh)
{
    const R forcing[2] = {forcing_nh, forcing_sh};
    R dkc[2], tul[2], pt0[2], dwc[2];
    R tu[2], cn[2], dpt[2]; // top sweep:    k_{i-1}/(dz dz_up) dt, -c'_{i-1}, d'_{i-1}
    R tb[2], an[2], dpb[2]; // bottom sweep: k_i/(dz dz) dt,       -a'_{i+1}, d"_{i+1}
    R *rb = col + (n - 1) * UDEB_ROW;
    R *pb = cp + (n - 1) * 2 * cs, *pt = cp; // c' columns: layer i, hemisphere h at cp[(2*i + h) * cs]
    const double *cb = ctab + (n - 1) * UDEB_CT;
    {
        const R at0 = R(ctab[0]), ab0 = R(ctab[1]), om0 = R(ctab[3]);
        const UdebRow<R> c = udeb_row<R>(cb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const R w = S[US_W0 + h];
            const R t0 = col[h * BLOCK]; // mixed-layer temperature before the solve (entrainment terms)
            const R tbot = rb[h * BLOCK];
            dkc[h] = y.dkdt_c * (t0 - tbot);
            const R delta_w = w - P[U_W0];
            const R dwv = (r_abs(delta_w) > R(1e-15)) ? delta_w : R(0);
            dwc[h] = y.dtdz * dwv;
            tul[h] = w * y.dtdz;
            pt0[h] = y.pi_ratio * tul[h] * t0;
            { // row 0: mixed layer
                const R k0 = r_max(om0 * dkc[h] + y.kc, y.kmin);
                const R term_diff = k0 * y.cM, term_upwell = w * y.dt_mix;
                const R b0 = R(1) + y.tfb_dt[h] * at0 + term_diff * ab0 + term_upwell * y.pi_ratio * ab0;
                R d0 = t0 + (forcing[h] * y.famp[h] + S[US_HX0 + h]) * y.dt_cmix * at0;
                if (y.lhc) d0 -= y.lhc_c[h] * (S[US_LAND0 + h] - S[US_GR0 + h]) * at0;
                d0 += y.dt_mix * dwv * R(ctab[4 + h]);
                const R r = r_rcp(b0);
                cn[h] = (term_diff + term_upwell) * ab0 * r;
                dpt[h] = d0 * r;
                pt[h * cs] = cn[h];
                col[h * BLOCK] = dpt[h];
                tu[h] = k0 * y.cA1; // the layer below the mixed layer sees half a layer thickness upwards
            }
            { // row n-1: bottom layer (no diffusion below)
                const R ku = r_max(c.omu * dkc[h] + y.kc, y.kmin);
                tb[h] = ku * y.cA;
                const R m = tb[h] * c.at;
                const R bi = R(1) + m + tul[h] * c.at;
                const R di = tbot + pt0[h] * c.at + dwc[h] * c.g[h];
                const R r = r_rcp(bi);
                an[h] = m * r;
                dpb[h] = di * r;
                pb[h * cs] = an[h];
                rb[h * BLOCK] = dpb[h];
            }
        }
    }
    // interior rows 1..n-2: the top sweep takes 1..k, the bottom sweep n-2..k+1 (one more when n is odd)
    const int k = (n - 2) >> 1;
    R *rt = col;
    const double *ctp = ctab;
    auto bottom_row = [&](R cu) {
        rb -= UDEB_ROW; cb -= UDEB_CT; pb -= 2 * cs;
        const UdebRow<R> c = udeb_row<R>(cb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const R ku = r_max(c.omu * dkc[h] + y.kc, y.kmin);
            const R m = ku * cu * c.at; // -a_i
            const R tdd = tb[h];
            const R bi = R(1) + m + tdd * c.ab + tul[h] * c.at;
            const R cnum = (tdd + tul[h]) * c.ab; // -c_i
            const R di = rb[h * BLOCK] + pt0[h] * c.ad + dwc[h] * c.g[h];
            const R r = r_rcp(bi - cnum * an[h]);
            an[h] = m * r;
            dpb[h] = (di + cnum * dpb[h]) * r;
            pb[h * cs] = an[h];
            rb[h * BLOCK] = dpb[h];
            tb[h] = ku * y.cA;
        }
    };
    for (int j = 0; j < k; ++j) {
        rt += UDEB_ROW; ctp += UDEB_CT; pt += 2 * cs;
        const UdebRow<R> c = udeb_row<R>(ctp);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const R tdd = r_max(c.om * dkc[h] + y.kc, y.kmin) * y.cA;
            const R m = tu[h] * c.at; // -a_i
            const R bi = R(1) + m + tdd * c.ab + tul[h] * c.at;
            const R cnum = (tdd + tul[h]) * c.ab; // -c_i
            const R di = rt[h * BLOCK] + pt0[h] * c.ad + dwc[h] * c.g[h];
            const R r = r_rcp(bi - m * cn[h]);
            cn[h] = cnum * r;
            dpt[h] = (di + m * dpt[h]) * r;
            pt[h * cs] = cn[h];
            rt[h * BLOCK] = dpt[h];
            tu[h] = tdd;
        }
        bottom_row(y.cA);
    }
    if (n & 1) bottom_row(n == 3 ? y.cA1 : y.cA);
    // rt = row k (top sweep's last), rb = row k+1 (bottom sweep's last): 2x2 solve, then outwards
    R xu[2], xd[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        xu[h] = (dpt[h] + cn[h] * dpb[h]) * r_rcp(R(1) - cn[h] * an[h]);
        xd[h] = dpb[h] + an[h] * xu[h];
        rt[h * BLOCK] = r_min(xu[h], y.tmax);
        rb[h * BLOCK] = r_min(xd[h], y.tmax);
    }
    auto down_row = [&]() {
        rb += UDEB_ROW; pb += 2 * cs;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            xd[h] = rb[h * BLOCK] + pb[h * cs] * xd[h];
            rb[h * BLOCK] = r_min(xd[h], y.tmax);
        }
    };
    for (int j = 0; j < k; ++j) {
        rt -= UDEB_ROW; pt -= 2 * cs;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            xu[h] = rt[h * BLOCK] + pt[h * cs] * xu[h];
            rt[h * BLOCK] = r_min(xu[h], y.tmax);
        }
        down_row();
    }
    if (n & 1) down_row();
}

// in: [ERF at_start, ERF at_end, Surface Temperature[4] at_start]
// out: [Heat Uptake, Ocean Heat Content, Sea Surface Temperature, Surface Temperature[4]]
template <class R>
__device__ inline bool climate_udeb_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef nr)
{
    if (S[US_OK] == R(0)) return false; // from_parameters failed for this member (LAMCALC did not converge)
    const int n = static_cast<int>(P[U_NLAYERS]), steps_n = static_cast<int>(P[U_STEPS]);
    R *col = cx.sm + nr.sm * BLOCK * (8 / static_cast<int>(sizeof(R))); // layer-major {T_nh, T_sh}, see udeb_step_both
    // The c' columns of the Thomas sweeps.  The per-thread shared-memory scratch is 100 8-byte words: in fp64 the
    // ocean columns fill it (two CTAs per SM) and c' goes to this run's rows of the global scratch (coalesced, lives
    // in L2: written and read back within one sub-step); in fp32 it holds both.
    R *cp;
    long long cs;
    if (sizeof(R) == 8) {
        cp = reinterpret_cast<R *>(cx.scratch0 + static_cast<long long>(nr.scr) * cx.runs) + cx.run;
        cs = cx.runs;
    } else {
        cp = col + 2 * UDEB_MAXL * BLOCK;
        cs = BLOCK;
    }
    const double *ctab = cx.ctab + nr.ctab;
    const R erf_start = in[0], erf_end = in[1];
    if (col[0] == R(0) && in[2] != R(0)) { // warm start from non-zero initial surface temperatures
        col[0] = in[2]; col[BLOCK] = in[4];
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
        // Which entries the window covers follows from the time axis alone (shared memory, block-uniform): entries
        // [first, nhist) count fully, entry first-1 with weight `partial` if the window ends inside it.  The values are
        // then summed in the reference's order with independent global loads the compiler can batch (a loop that decides
        // and loads in one go serialises a load latency per year of history: a third of the kernel's time at 350 years).
        R rem = P[U_FB_PERIOD], partial = R(0);
        int first = nhist;
        for (int i = nhist - 1; i >= 0; --i) {
            if (rem <= R(0)) break;
            const R dt = R(cx.bounds[i + 1] - cx.bounds[i]);
            if (dt <= rem) { first = i; rem -= dt; }
            else { partial = rem / dt; rem = R(0); }
        }
        const double *hist = cx.scratch + static_cast<long long>(nr.scr + 2 * UDEB_MAXL) * cx.runs;
        R sum = R(0);
#pragma unroll 8
        for (int i = nhist - 1; i >= first; --i) sum += R(hist[static_cast<long long>(i) * cx.runs]);
        if (partial > R(0) && first > 0) sum += R(hist[static_cast<long long>(first - 1) * cx.runs]) * partial;
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
    UdebYear<R> y;
    {
        const R dz = P[U_DZ], dz_mix = P[U_MLD], dz1 = dz / R(2), conv = R(3155.76); // DIFFUSIVITY_CM2S_TO_M2YR
        const R c_mix = udeb_heat_capacity(dz_mix);
        y.cA = dt_sub / (dz * dz);
        y.cA1 = dt_sub / (dz * dz1);
        y.cM = dt_sub / (dz_mix * dz1);
        y.dt_mix = dt_sub / dz_mix;
        y.dtdz = dt_sub / dz;
        y.dt_cmix = dt_sub / c_mix;
        y.kc = P[U_KAPPA] * conv;
        y.kmin = P[U_KAPPA_MIN] * conv;
        y.dkdt_c = P[U_KAPPA_DKDT] * conv;
        y.pi_ratio = P[U_PI_RATIO];
        y.tmax = P[U_TMAX];
        y.lhc = lhc;
        for (int h = 0; h < 2; ++h) {
            const R f_l = (h == 0 ? fgnl : fgsl), f_o = R(0.5) - f_l;
            const R denominator = f_o * (P[U_KLO] + f_l * lam_l);
            y.tfb_dt[h] = S[US_AE0 + h] / c_mix * (lam_o + lam_l * P[U_KLO] * P[U_AMP] * f_l / denominator) * dt_sub;
            y.famp[h] = R(1) + P[U_KLO] * f_l / denominator;
            y.lhc_c[h] = lhc ? P[U_KLG] / (c_mix * f_o) * dt_sub : R(0);
        }
    }
    const R gr_c0 = (lhc && !(fgnl < R(1e-15))) ? P[U_KLG] / (fgnl * c_ground) * dt_sub : R(0);
    const R gr_c1 = (lhc && !(fgsl < R(1e-15))) ? P[U_KLG] / (fgsl * c_ground) * dt_sub : R(0);
    const R inv_steps = R(1) / steps;
    const R hx_c0 = (fgno > R(1e-15)) ? P[U_KNS] / fgno : R(0), hx_c1 = (fgso > R(1e-15)) ? P[U_KNS] / fgso : R(0);
    const R w0 = P[U_W0], fv = P[U_WVAR], wmin = w0 * (R(1) - fv);
    const R inv_wt_nh = R(1) / P[U_WT_NH], inv_wt_sh = R(1) / P[U_WT_SH];
    for (int step = 1; step <= steps_n; ++step) {
        const R frac = R(step) * inv_steps;
        const R erf = erf_start + frac * (erf_end - erf_start);
        R forcing[4];
        udeb_apply_efficacy(P, S, erf, co2_eff, forcing);
        if (lhc) {
            if (!(fgnl < R(1e-15))) S[US_GR0] += gr_c0 * (S[US_LAND0] - S[US_GR0]);
            if (!(fgsl < R(1e-15))) S[US_GR1] += gr_c1 * (S[US_LAND1] - S[US_GR1]);
        }
        udeb_step_both(y, P, S, ctab, n, col, cp, cs, forcing[0], forcing[2]);
        const R air_nho = udeb_sst_to_air(P, col[0]), air_sho = udeb_sst_to_air(P, col[BLOCK]);
        S[US_LAND0] = udeb_land_temperature(P, air_nho, forcing[1], fgnl, lam_l);
        S[US_LAND1] = udeb_land_temperature(P, air_sho, forcing[3], fgsl, lam_l);
        if (fgno > R(1e-15)) S[US_HX0] = hx_c0 * (air_sho - air_nho);
        if (fgso > R(1e-15)) S[US_HX1] = hx_c1 * (air_nho - air_sho);
        const R gt = air_nho * fgno + S[US_LAND0] * fgnl + air_sho * fgso + S[US_LAND1] * fgsl;
        S[US_W0] = r_max(w0 * (R(1) - fv * r_min(gt * inv_wt_nh, R(1))), wmin);
        S[US_W1] = r_max(w0 * (R(1) - fv * r_min(gt * inv_wt_sh, R(1))), wmin);
    }
    const R sst_nh = col[0], sst_sh = col[BLOCK];
    S[US_AE0] = (r_abs(sst_nh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(P, sst_nh) / sst_nh;
    S[US_AE1] = (r_abs(sst_sh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(P, sst_sh) / sst_sh;
    const R st[4] = {udeb_sst_to_air(P, sst_nh), S[US_LAND0], udeb_sst_to_air(P, sst_sh), S[US_LAND1]};
    const R gt = st[0] * fgno + st[1] * fgnl + st[2] * fgso + st[3] * fgsl;
    cx.scratch[static_cast<long long>(nr.scr + 2 * UDEB_MAXL + nhist) * cx.runs] = static_cast<double>(gt * dt_year);
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
            const R *T = col + h * BLOCK;
            total += rho_c * P[U_MLD] * T[0];
            for (int l = 1; l < n; ++l) total += rho_c * P[U_DZ] * T[l * UDEB_ROW];
        }
        out[1] = total / R(2);
    }
    out[2] = (sst_nh + sst_sh) / R(2);
    for (int i = 0; i < 4; ++i) out[3 + i] = st[i];
    return true;
}

} // namespace rscm_dev
