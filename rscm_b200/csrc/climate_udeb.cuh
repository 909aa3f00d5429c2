// climate_udeb.cuh — MAGICC upwelling-diffusion energy-balance model (ClimateUDEB) on the device.
//
// Reference: crates/rscm-magicc/src/climate/udeb/mod.rs (solve_impl :399-660, adjusted_ecs :302-350,
// apply_efficacy_and_qfrac :253-270, sst_to_air_temperature :377-397, calculate_land_temperature :352-375),
// udeb/ocean_column.rs (step_hemisphere :54-241, layer_diffusivities :23-52, update_upwelling :243-259,
// heat uptake / ocean heat content :262-306), climate/lamcalc.rs (:85-290), climate/state.rs,
// rscm-core/src/utils/linear_algebra.rs (thomas_solve :41-79, invert_4x4 :102-166).
//
// FOUR THREADS = ONE MEMBER (Prog::LANES = 4).  The reference solves, twelve times a year and for each hemisphere, a
// 50-row tridiagonal system top to bottom: one 50-long recurrence with a division in the loop-carried chain.  Here the
// system is eliminated from both ends towards the middle (rows 0..k the usual way, rows n-1..k+1 mirrored, a 2x2 solve
// where they meet, substitution outwards — same solution up to rounding: the matrix is strictly diagonally dominant),
// and each of the four half-sweeps of a member (2 hemispheres x 2 ends) belongs to one ROLE:
//     role & 1 = hemisphere (0 NH, 1 SH),   role & 2 = end (0: top sweep, rows 0..k;  2: bottom sweep, rows n-1..k+1).
// A CTA of four warps serves 32 members: warp w is role w, lane l is member l (StepCtx, components.cuh), so a warp runs
// ONE role — no selects or divergence between the top and the bottom sweep — and warp 0 alone runs everything that is
// not a row: the rest of the component graph, the history sum of the ECS feedback and LAMCALC, whose results it
// publishes; the other warps wait at a barrier meanwhile and cost no issue slots.  A thread keeps ITS <= 25 rows in
// registers for the whole run (the temperatures persist in S[], the eliminated off-diagonal lives only inside a sub-step):
// there is no shared-memory column for them and no address arithmetic.  Per sub-step the four roles meet twice through
// the member's exchange column in shared memory, each time followed by a __syncthreads(): the pivot pair where the
// sweeps meet, and after the substitution the new edge temperatures (the top roles' are the sea-surface temperatures
// every role needs for the land boxes; all four are what the next sub-step's sweeps start from).  The small land / ground /
// upwelling updates of a sub-step are computed by all four roles, so that they hold identical scalar state.
// Row coefficients that depend only on geometry (area factors, entrainment combinations of the initial profile,
// relative depth of the diffusivity profile) come from a host-computed table laid out per role and row
// (graph.cpp udeb_const_table), read with 128-bit shared-memory loads at immediate offsets.
//
// Where the state lives:
//   * S[0..18]   (registers, identical in the four roles): LAMCALC result at the member's base ECS, upwelling rates,
//                land / ground temperatures, alpha_eff, inter-hemispheric exchange, history length;
//   * S[19..43]  (registers): this role's rows of the ocean column, sweep order (row 0 = the role's end of the column);
//   * cx.sm      (shared memory, [25][BLOCK] per CTA, conflict-free): the eliminated off-diagonal of the role's rows,
//                written by the sweep and read back by the substitution of the same sub-step;
//   * cx.xch     (shared memory, 28 doubles per member): the exchange column;
//   * cx.scratch (global, member-interleaved [T][runs]): the T*dt history of the cumulative-temperature feedback, each
//                role sums a quarter of the window newest-to-oldest and role 0 adds the four partial sums;
//   * cx.ctab    (shared memory, per graph): the role tables.
#pragma once

namespace rscm_dev {

enum {
    U_NLAYERS, U_MLD, U_DZ, U_KAPPA, U_KAPPA_MIN, U_KAPPA_DKDT, U_W0, U_WVAR, U_WT_NH, U_WT_SH, U_ECS, U_RF2X, U_RLO,
    U_FB_Q, U_FB_CUMT, U_FB_PERIOD, U_KLO, U_KNS, U_AMP, U_NH_LAND, U_SH_LAND, U_DDA, U_TA_ALPHA, U_TA_GAMMA, U_PI_RATIO,
    U_LHC_ON, U_KLG, U_LHC_THICK, U_RFR0, U_RFR1, U_RFR2, U_RFR3, U_EFF_APPLY, U_EFF_CO2, U_PROFILE, U_STEPS, U_TMAX, U_NPARAM
};
enum { US_OK, US_LAMO, US_LAML, US_EFF, US_QF0, US_QF1, US_QF2, US_QF3, US_W0, US_W1, US_LAND0, US_LAND1, US_GR0, US_GR1,
       US_AE0, US_AE1, US_HX0, US_HX1, US_NHIST, US_T };

#define RSCM_INF_F (__int_as_float(0x7f800000))
constexpr int UDEB_LANES = 4;
constexpr int UDEB_MAXR = 25;       // rows per lane (n_layers <= 50)
constexpr int UDEB_CT = 6;          // role table: doubles per row {near area, far area, af_diff, g, omr, af_top}
constexpr int UDEB_NS = US_T + UDEB_MAXR;

// f64::min / f64::max: NaN-ignoring, like fmin / fmax
__device__ __forceinline__ double r_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float r_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }

// Reciprocal for the Thomas sweep: the pivots of the diagonally dominant ocean-column system are >= 1, never
// subnormal or infinite, so the special-case slow path of 1/x is dropped: MUFU seed (relative error e0 ~ 2^-20) and
// one cubically convergent step r (1 + e + e^2), e = 1 - x r: error e0^3 ~ 2^-60 plus rounding (<= 1 ulp; a NaN pivot
// stays NaN).  Three dependent FMAs in the loop-carried chain of the sweep instead of four.
__device__ __forceinline__ double r_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
__device__ __forceinline__ float r_rcp(float x) { return __frcp_rn(x); }

template <class R, class PT = R> __device__ __forceinline__ void udeb_fractions(const PT *P, R (&a)[4])
{
    const R fgnl = R(P[U_NH_LAND]) / R(2), fgsl = R(P[U_SH_LAND]) / R(2);
    a[0] = R(0.5) - fgnl; a[1] = fgnl; a[2] = R(0.5) - fgsl; a[3] = fgsl;
}

template <class R, class PT = R> __device__ __forceinline__ void udeb_qfrac(const PT *P, const R (&area)[4], R (&q)[4])
{
    R s = R(0);
    for (int i = 0; i < 4; ++i) s += R(P[U_RFR0 + i]) * area[i];
    for (int i = 0; i < 4; ++i) q[i] = (r_abs(s) <= R(1e-15)) ? R(1) : R(P[U_RFR0 + i]) / s;
}

// lamcalc — hybrid step / secant iteration on lambda_ocean, tolerance 1e-3 on the land/ocean warming ratio.
// The reference inverts the 4x4 box-coupling matrix by Gauss-Jordan (linear_algebra.rs:102-166); its sparsity
//   [ A0      -k_lo   -k_ns    0    ]      rows 1 and 3 couple each land box to its own ocean box only,
//   [ -k_lo*a  B1      0       0    ]      so eliminating them leaves a 2x2 system in the two ocean boxes,
//   [ -k_ns    0       A2     -k_lo ]      solved here by Cramer's rule (same solution up to rounding; the
//   [ 0        0      -k_lo*a  B3   ]      singular-pivot failure becomes a vanishing B1, B3 or determinant).
// Only the last three iterates of the reference's lamo[]/diff[] arrays are ever read: kept as scalars.
// (R = the arithmetic, PT = the type the parameters are stored in: the fp32 path runs LAMCALC in fp64 — its secant
// iteration stops at the first iterate with |diff| < 0.001, and which iterate that is must not depend on the precision:
// stopping one iterate earlier or later moves the feedback parameters by up to 1e-3)
template <class R, class PT = R> __device__ inline bool udeb_lamcalc(const PT *P, R ecs, R &lam_o_out, R &lam_l_out, R &eff_out)
{
    constexpr int MAXIT = 40;
    const R q2x = R(P[U_RF2X]), k_lo = R(P[U_KLO]), k_ns = R(P[U_KNS]), rlo = R(P[U_RLO]), alpha = R(P[U_AMP]);
    R area[4], qfrac[4];
    udeb_fractions<R, PT>(P, area);
    udeb_qfrac<R, PT>(P, area, qfrac);
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
        const R r1 = r_rcp(B1), r3 = r_rcp(B3); // |B| >= 1e-15 checked above, far from the subnormal range
        const R a00 = fgno * lam_o + kla + k_ns - k_lo * kla * r1;
        const R a22 = fgso * lam_o + kla + k_ns - k_lo * kla * r3;
        const R g0 = f0 + k_lo * f1 * r1, g2 = f2 + k_lo * f3 * r3;
        const R det = a00 * a22 - k_ns * k_ns;
        if (r_abs(det) < R(1e-15)) return false;
        const R rd = r_rcp(det);
        const R t0 = (g0 * a22 + k_ns * g2) * rd, t2 = (a00 * g2 + k_ns * g0) * rd;
        const R t1 = (f1 + kla * t0) * r1, t3 = (f3 + kla * t2) * r3;
        const R ocean_mean = (fgno * t0 + fgso * t2) * inv_o;
        const R land_mean = (fgnl * t1 + fgsl * t3) * inv_l;
        const R d2 = rlo - land_mean / ocean_mean;
        if (r_abs(d2) < R(0.001)) {
            R rf_sum = R(0);
            for (int c = 0; c < 4; ++c) rf_sum += R(P[U_RFR0 + c]) * area[c];
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

// sst_to_air_temperature with the member's constants hoisted: t_star = -(alpha - 1)/(2 gamma) and the offset of the
// linear continuation above it (mod.rs:377-397) do not change during a run
template <class R> struct UdebAir { R alpha, gamma, t_star, offs; bool quad; };
template <class R> __device__ __forceinline__ UdebAir<R> udeb_air_constants(const R *P)
{
    UdebAir<R> a;
    a.alpha = P[U_TA_ALPHA]; a.gamma = P[U_TA_GAMMA];
    a.quad = r_abs(a.gamma) > R(1e-15);
    a.t_star = a.quad ? -(a.alpha - R(1)) / (R(2) * a.gamma) : R(0);
    a.offs = a.alpha * a.t_star + a.gamma * a.t_star * a.t_star - a.t_star;
    return a;
}
template <class R> __device__ __forceinline__ R udeb_sst_to_air(const UdebAir<R> &a, R sst)
{
    const R lin = sst + a.offs, par = a.alpha * sst + a.gamma * sst * sst;
    return (a.quad && !(sst < a.t_star)) ? lin : par;
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
template <class R, int N> __device__ inline void climate_udeb_init_state(const R *P, const R *, R *S, const StepCtx<R> &, NodeRef)
{
    R area[4], q[4];
    udeb_fractions(P, area);
    udeb_qfrac(P, area, q);
    R lo = R(0), ll = R(0), ef = R(1);
    double lo_d = 0.0, ll_d = 0.0, ef_d = 1.0;
    const bool ok = (P[U_EFF_CO2] > R(0)) && (P[U_EFF_CO2] == P[U_EFF_CO2]) && udeb_lamcalc<double, R>(P, static_cast<double>(P[U_ECS]), lo_d, ll_d, ef_d);
    lo = R(lo_d); ll = R(ll_d); ef = R(ef_d);
    S[US_OK] = ok ? R(1) : R(0);
    S[US_LAMO] = lo; S[US_LAML] = ll; S[US_EFF] = ef;
    for (int i = 0; i < 4; ++i) S[US_QF0 + i] = q[i];
    S[US_W0] = S[US_W1] = P[U_W0];
    S[US_LAND0] = S[US_LAND1] = S[US_GR0] = S[US_GR1] = R(0);
    S[US_AE0] = S[US_AE1] = P[U_TA_ALPHA];
    S[US_HX0] = S[US_HX1] = R(0);
    S[US_NHIST] = R(0);
#pragma unroll
    for (int j = 0; j < UDEB_MAXR; ++j) S[US_T + j] = R(0);
}

// Constants of one model year for the sweeps (everything that does not change between sub-steps), for this lane's
// hemisphere.
template <class R> struct UdebYear {
    R cA, cM, dt_mix, dtdz, dt_cmix, kc, kmin, dkdt_c, pi_ratio, tmax;
    R kcA, kminA; // kc cA, kmin cA: interior rows compute k cA = max(omr (dkdt_c dT cA) + kc cA, kmin cA) directly
    R tfb_dt, famp, lhc_c;
    bool lhc;
};

template <class R> struct UdebRow { R an, af, ad, g, om, at; };

template <class R> __device__ __forceinline__ UdebRow<R> udeb_row(const double *ct)
{
    const double2 c01 = *reinterpret_cast<const double2 *>(ct), c23 = *reinterpret_cast<const double2 *>(ct + 2),
                  c45 = *reinterpret_cast<const double2 *>(ct + 4); // 48-byte rows: 16-byte aligned
    UdebRow<R> r;
    r.an = R(c01.x); r.af = R(c01.y); r.ad = R(c23.x); r.g = R(c23.y); r.om = R(c45.x); r.at = R(c45.y);
    return r;
}

// exchange slots of a member's column (relative to NodeRef::xch), see climate_udeb_solve
enum { UX_EDGE = 0, UX_MEET = 4, UX_HIST = 14, UX_LAM = 18, UX_OHC = 22, UX_T0 = 26 };

// f64::max(x, lo) / f64::min(x, hi) for a bound that is not NaN (the year set-up replaces a NaN bound by -inf / +inf,
// which is what the NaN-ignoring f64::max / min make of it): one compare and a select instead of the library fmax / fmin
// (compare, NaN quieting, selects: eight instructions).  A NaN x yields the bound, as f64::max / min do.
// (written as setp + selp: the compiler turns the C++ conditional back into the fmax / fmin pattern)
__device__ __forceinline__ double udeb_floor(double x, double lo)
{
    double r;
    asm("{\n.reg .pred p;\nsetp.gt.f64 p, %1, %2;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(x), "d"(lo));
    return r;
}
__device__ __forceinline__ double udeb_cap(double x, double hi)
{
    double r;
    asm("{\n.reg .pred p;\nsetp.lt.f64 p, %1, %2;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(x), "d"(hi));
    return r;
}
__device__ __forceinline__ float udeb_floor(float x, float lo) { return x > lo ? x : lo; }
__device__ __forceinline__ float udeb_cap(float x, float hi) { return x < hi ? x : hi; }

// One sub-step of step_hemisphere for this lane's half-sweep.
//
// Row i of the system (ocean_column.rs:118-160), with tup = k_{i-1} dt/(dz dz_up), tdn = k_i dt/dz^2, tul = w dt/dz:
//     -a_i = tup af_top_i,   -c_i = (tdn + tul) af_bottom_i,   b_i = 1 - a_i + tdn af_bottom_i + tul af_top_i,
//     d_i = T_i + pi_ratio tul T_0 af_diff_i + dt/dz dw g_i.
// A sweep meets the coupling towards the row it has already eliminated ("near") and the one towards the next row
// ("far"); for the top sweep near = -a_i, far = -c_i, for the bottom sweep the other way round.  With the diffusivity the
// previous row computed carried along, both are (carry + u_n) area_n and (k_new + u_f) area_f, where (u_n, u_f) =
// (0, tul) / (tul, 0) and the areas come from the role table — one instruction stream for all four lanes, no selects:
//     b_i = 1 + near + far + tul (af_top_i - af_bottom_i),        pivot = b_i - near f_prev,
//     f_i = far / pivot,   d'_i = (d_i + near d'_prev) / pivot,   back substitution  x_i = d'_i + f_i x_next.
// T holds the lane's rows (sweep order); on return it holds the new temperatures, capped (ocean_column.rs:226-238: the
// cap applies to what is stored, the substitution carries the uncapped value).
// N = n_layers is a compile-time constant of the program (the graph compiler passes it as a template argument), so the
// row loops are straight-line code over register rows: the top sweep owns rows 0..K (NT = K + 1 of them), the bottom
// sweep rows N-1..K+1 (NB = N - K - 1, one more than NT when N is odd — that row is the only predicated one).
template <class R, int N>
__device__ __forceinline__ void udeb_substep(const UdebYear<R> &y, const R *P, const R *S, const double *tab, int role, double *X, R *T,
                                             R *F, R forcing)
{
    const bool bottom = (role & 2) != 0; // warp-uniform
    const int h = role & 1;
    constexpr int K = (N - 2) >> 1, NT = K + 1, NB = N - K - 1;
    static_assert(NT >= 1 && NB >= NT && NB - NT <= 1 && NB <= UDEB_MAXR, "row split");
    // F: the eliminated off-diagonal of this lane's rows, row j at F[j * BLOCK] (this thread's shared-memory column:
    // consecutive threads, consecutive words); it lives only inside the sub-step
    // edge temperatures of this hemisphere before the solve: the mixed layer (top lane's row 0) and the bottom layer
    // (every role published its edge temperature T[0] at the end of the previous sub-step, or at the start of the year)
    const R tedge = T[0];
    const R tpart = R(X[(UX_EDGE + (role ^ 2)) * 32]);
    const R t0 = bottom ? tpart : tedge, tbot = bottom ? tedge : tpart;
    const R w = h ? S[US_W1] : S[US_W0];
    const R dkc = y.dkdt_c * (t0 - tbot), dkcA = dkc * y.cA;
    const R delta_w = w - P[U_W0];
    const R dwv = (r_abs(delta_w) > R(1e-15)) ? delta_w : R(0);
    const R dwc = y.dtdz * dwv;
    const R tul = w * y.dtdz;
    const R pt0 = y.pi_ratio * tul * t0;
    const R u_n = bottom ? tul : R(0), u_f = bottom ? R(0) : tul;
    // The loop-carried chain of the sweep is pivot (FMA) -> reciprocal (MUFU + 3 FMA) -> f (MUL): about 57 cycles per row
    // at the measured 8-cycle DFMA and 17-cycle MUFU.RCP64H latencies (tools/micro/fp64_latency.cu), which three warps per
    // scheduler cover.  UDEB_DIVISION_FREE selects an alternative with three LINEAR recurrences over q_j = product of the
    // pivots (g_j = far_j q_{j-1}, q_j = b_j q_{j-1} - near_j g_{j-1}, D_j = d_j q_{j-1} + near_j D_{j-1}; f_j = g_j / q_j,
    // d'_j = D_j / q_j): one FMA in the chain, independent reciprocals — but five more FP64 instructions per row, and the
    // kernel is bound by issue slots, not by latency: measured 1.90e8 against 1.98e8 member-years/s for the chained form.
    R carry, q1, g1, D1, fp, dp;
    {
        const UdebRow<R> c = udeb_row<R>(tab);
        const R k = udeb_floor(c.om * dkc + y.kc, y.kmin);
        if (!bottom) { // row 0 of the top sweep: mixed layer (ocean_column.rs:93-150)
            const R term_diff = k * y.cM, term_upwell = w * y.dt_mix;
            q1 = R(1) + y.tfb_dt * c.at + term_diff * c.af + term_upwell * y.pi_ratio * c.af;
            D1 = t0 + (forcing * y.famp + (h ? S[US_HX1] : S[US_HX0])) * y.dt_cmix * c.at;
            if (y.lhc) D1 -= y.lhc_c * ((h ? S[US_LAND1] : S[US_LAND0]) - (h ? S[US_GR1] : S[US_GR0])) * c.at;
            D1 += y.dt_mix * dwv * c.g;
            g1 = (term_diff + term_upwell) * c.af;
        } else { // row 0 of the bottom sweep: bottom layer, no diffusion below (:163-175)
            g1 = k * y.cA * c.at;
            q1 = R(1) + g1 + tul * c.at;
            D1 = tbot + pt0 * c.at + dwc * c.g;
        }
        const R r = r_rcp(q1);
        fp = g1 * r;
        dp = D1 * r;
        carry = k * y.cA;
        F[0] = fp;
        T[0] = dp;
    }
    auto row = [&](int j) {
        const UdebRow<R> c = udeb_row<R>(tab + j * UDEB_CT);
        const R knew = udeb_floor(c.om * dkcA + y.kcA, y.kminA);
        const R near = (carry + u_n) * c.an;
        const R far = (knew + u_f) * c.af;
        const R bi = (R(1) + near) + (far + tul * c.ad);
        const R di = T[j] + pt0 * c.ad + dwc * c.g;
#ifndef UDEB_DIVISION_FREE
        const R r = r_rcp(bi - near * fp);
        fp = far * r;
        dp = (di + near * dp) * r;
        const R g = fp, q = R(1), D = dp;
#else
        const R g = far * q1;
        const R q = bi * q1 - near * g1;
        const R D = di * q1 + near * D1;
        const R r = r_rcp(q);
        fp = g * r;
        dp = D * r;
#endif
        F[j * BLOCK] = fp;
        T[j] = dp;
        carry = knew;
        q1 = q; g1 = g; D1 = D;
    };
#pragma unroll
    for (int j = 1; j < NT; ++j) row(j);
    if (NB > NT) {
        if (bottom) row(NB - 1);
    }
    // where the sweeps meet: x_K = d'_K + f_K x_{K+1} (top), x_{K+1} = d"_{K+1} + f"_{K+1} x_K (bottom)
    X[(UX_MEET + 2 * role) * 32] = static_cast<double>(fp);
    X[(UX_MEET + 2 * role + 1) * 32] = static_cast<double>(dp);
    __syncthreads();
    const R fq = R(X[(UX_MEET + 2 * (role ^ 2)) * 32]), dq = R(X[(UX_MEET + 2 * (role ^ 2) + 1) * 32]);
    const R f_top = bottom ? fq : fp, d_top = bottom ? dq : dp, d_bot = bottom ? dp : dq;
    const R x_top = (d_top + f_top * d_bot) * r_rcp(R(1) - fp * fq);
    R x = bottom ? dp + fp * x_top : x_top;
    if (NB > NT) {
        if (bottom) {
            T[NB - 1] = udeb_cap(x, y.tmax);
            x = T[NT - 1] + F[(NT - 1) * BLOCK] * x;
        }
    }
    T[NT - 1] = udeb_cap(x, y.tmax);
#pragma unroll
    for (int j = NT - 2; j >= 0; --j) {
        x = T[j] + F[j * BLOCK] * x;
        T[j] = udeb_cap(x, y.tmax);
    }
}

// RSCM_NODE_CLOCKS (profiling aid, set by the JIT when RSCM_B200_NODE_CLOCKS is in the environment): thread 0 of CTA 0
// accumulates the cycles of the phases of a model year and prints the averages after the last step.
#ifdef RSCM_NODE_CLOCKS
#define UDEB_CLK(i) do { if (clk_on) { const long long now_ = clock64(); udeb_clk[i] += now_ - clk_t; clk_t = now_; } } while (0)
#else
#define UDEB_CLK(i) do { } while (0)
#endif

// in: [ERF at_start, ERF at_end, Surface Temperature[4] at_start]
// out: [Heat Uptake, Ocean Heat Content, Sea Surface Temperature, Surface Temperature[4]]
template <class R, int N>
__device__ inline bool climate_udeb_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef nr)
{
    // No early return: the four roles of a member meet at every barrier, and so must all members of the CTA.  A member
    // whose from_parameters failed (LAMCALC did not converge) computes on and reports failure at the end.
    const bool ok = S[US_OK] != R(0);
    const int steps_n = static_cast<int>(P[U_STEPS]);
    const int q = cx.role, h = q & 1;
#ifdef RSCM_NODE_CLOCKS
    __shared__ long long udeb_clk[8];
    const bool clk_on = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    if (clk_on && cx.N == 0) for (int i = 0; i < 8; ++i) udeb_clk[i] = 0;
    long long clk_t = clock64();
#endif
    const bool bottom = (q & 2) != 0;
    constexpr int K = (N - 2) >> 1, NT = K + 1, NB = N - K - 1;
    R *T = S + US_T;
    R *F = cx.sm + nr.sm * BLOCK * (8 / static_cast<int>(sizeof(R)));
    double *X = cx.xch + nr.xch * 32; // this member's exchange column, slot j at X[j * 32]
    const double *tab = cx.ctab + nr.ctab + q * (UDEB_MAXR * UDEB_CT);
    const R erf_start = in[0], erf_end = in[1];
    const R dt_year = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    const R steps = R(steps_n);
    const R dt_sub = dt_year / steps;

    // adjusted_ecs: sum of T*dt over the last `period` years
    const int nhist = static_cast<int>(S[US_NHIST]);
    {
        // Which entries the window covers follows from the time axis alone: entries [first, nhist) count fully, entry
        // first-1 with weight `partial` if the window ends inside it.  Each role sums every fourth entry, newest first,
        // with independent loads; role 0 adds the four partial sums.
        R partial = R(0);
        int first = nhist;
        const double *win = cx.gtab + nr.gt; // host-computed window per step for the graph's period (graph.cpp)
        if (static_cast<double>(P[U_FB_PERIOD]) == win[0]) {
            first = static_cast<int>(win[2 + 2 * nhist]);
            partial = R(win[3 + 2 * nhist]);
        } else { // period bound per member: walk the axis
            R rem = P[U_FB_PERIOD];
            for (int i = nhist - 1; i >= 0; --i) {
                if (rem <= R(0)) break;
                const R dt = R(cx.bounds[i + 1] - cx.bounds[i]);
                if (dt <= rem) { first = i; rem -= dt; }
                else { partial = rem / dt; rem = R(0); }
            }
        }
        const double *hist = cx.scratch + static_cast<long long>(nr.scr) * SCR_LD;
        R sum = R(0);
#pragma unroll 8
        for (int i = nhist - 1 - q; i >= first; i -= UDEB_LANES) sum += R(hist[i * SCR_LD]);
        if (q == 0 && partial > R(0) && first > 0) sum += R(hist[(first - 1) * SCR_LD]) * partial;
        X[(UX_HIST + q) * 32] = static_cast<double>(sum);
        if (q == 0) X[UX_T0 * 32] = static_cast<double>(T[0]); // the NH mixed-layer temperature (warm-start test below)
    }
    __syncthreads();
    UDEB_CLK(0); // history sums + barrier
    // Role 0 alone turns the history into this year's feedback parameters (LAMCALC: up to 40 secant iterations) and
    // publishes them; the other warps wait at the barrier.
    if (q == 0) {
        const R cum_t = ((R(X[UX_HIST * 32]) + R(X[(UX_HIST + 1) * 32])) + (R(X[(UX_HIST + 2) * 32]) + R(X[(UX_HIST + 3) * 32])));
        const R cumt_2x = P[U_ECS] * P[U_FB_PERIOD];
        const R cumt_factor = (r_abs(cumt_2x) > R(1e-15)) ? R(1) + P[U_FB_CUMT] * (cum_t - cumt_2x) / cumt_2x : R(1);
        const R erf_mid = (erf_start + erf_end) / R(2);
        const R q_factor = R(1) + P[U_FB_Q] * (udeb_floor(erf_mid, R(0)) - P[U_RF2X]);
        const R aecs = P[U_ECS] * cumt_factor * q_factor;
        R lo = S[US_LAMO], ll = S[US_LAML], ef = S[US_EFF];
        if (r_abs(aecs - P[U_ECS]) > R(1e-10)) {
            double lo2, ll2, ef2;
            if (udeb_lamcalc<double, R>(P, static_cast<double>(aecs), lo2, ll2, ef2)) { lo = R(lo2); ll = R(ll2); ef = R(ef2); }
        }
        X[UX_LAM * 32] = static_cast<double>(lo);
        X[(UX_LAM + 1) * 32] = static_cast<double>(ll);
        X[(UX_LAM + 2) * 32] = static_cast<double>(ef);
    }
    UDEB_CLK(1); // LAMCALC (role 0)
    __syncthreads();
    const R lam_o = R(X[UX_LAM * 32]), lam_l = R(X[(UX_LAM + 1) * 32]), co2_eff = R(X[(UX_LAM + 2) * 32]);
    if (R(X[UX_T0 * 32]) == R(0) && in[2] != R(0)) { // warm start from non-zero initial surface temperatures (mod.rs:439-448)
        if (!bottom) T[0] = h ? in[4] : in[2];
        S[US_LAND0] = in[3]; S[US_LAND1] = in[5];
        S[US_GR0] = S[US_LAND0]; S[US_GR1] = S[US_LAND1];
    }
    X[(UX_EDGE + q) * 32] = static_cast<double>(T[0]); // edge temperatures for the first sub-step
    __syncthreads();
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
        y.cM = dt_sub / (dz_mix * dz1);
        y.dt_mix = dt_sub / dz_mix;
        y.dtdz = dt_sub / dz;
        y.dt_cmix = dt_sub / c_mix;
        y.kc = P[U_KAPPA] * conv;
        y.kmin = P[U_KAPPA_MIN] * conv;
        if (!(y.kmin == y.kmin)) y.kmin = -R(RSCM_INF_F);
        y.dkdt_c = P[U_KAPPA_DKDT] * conv;
        y.pi_ratio = P[U_PI_RATIO];
        y.tmax = (P[U_TMAX] == P[U_TMAX]) ? P[U_TMAX] : R(RSCM_INF_F);
        y.kcA = y.kc * y.cA;
        y.kminA = y.kmin * y.cA;
        y.lhc = lhc;
        const R f_l = (h == 0 ? fgnl : fgsl), f_o = R(0.5) - f_l;
        const R denominator = f_o * (P[U_KLO] + f_l * lam_l);
        y.tfb_dt = (h ? S[US_AE1] : S[US_AE0]) / c_mix * (lam_o + lam_l * P[U_KLO] * P[U_AMP] * f_l / denominator) * dt_sub;
        y.famp = R(1) + P[U_KLO] * f_l / denominator;
        y.lhc_c = lhc ? P[U_KLG] / (c_mix * f_o) * dt_sub : R(0);
    }
    const R gr_c0 = (lhc && !(fgnl < R(1e-15))) ? P[U_KLG] / (fgnl * c_ground) * dt_sub : R(0);
    const R gr_c1 = (lhc && !(fgsl < R(1e-15))) ? P[U_KLG] / (fgsl * c_ground) * dt_sub : R(0);
    const UdebAir<R> air = udeb_air_constants(P);
    // efficacy scaling of the forcing (apply_efficacy_and_qfrac, mod.rs:253-270) and the land-box denominators are
    // constant over the year
    R e_scale = R(1);
    {
        const int mode = static_cast<int>(P[U_EFF_APPLY]);
        if (mode == 1) e_scale = P[U_EFF_CO2];
        else if (mode == 2 && (co2_eff - co2_eff) == R(0) && co2_eff > R(0)) e_scale = P[U_EFF_CO2] / co2_eff;
    }
    const R kla = P[U_KLO] * P[U_AMP];
    const R inv_den_n = R(1) / (lam_l * fgnl + P[U_KLO]), inv_den_s = R(1) / (lam_l * fgsl + P[U_KLO]);
    const R inv_steps = R(1) / steps;
    const R hx_c0 = (fgno > R(1e-15)) ? P[U_KNS] / fgno : R(0), hx_c1 = (fgso > R(1e-15)) ? P[U_KNS] / fgso : R(0);
    const R w0 = P[U_W0], fv = P[U_WVAR], wmin = w0 * (R(1) - fv);
    const R wmin_b = (wmin == wmin) ? wmin : -R(RSCM_INF_F);
    const R inv_wt_nh = R(1) / P[U_WT_NH], inv_wt_sh = R(1) / P[U_WT_SH];
    R sst_nh = R(0), sst_sh = R(0);
    UDEB_CLK(2); // the year's constants
    for (int step = 1; step <= steps_n; ++step) {
        const R frac = R(step) * inv_steps;
        const R erf = erf_start + frac * (erf_end - erf_start);
        const R e = (e_scale == R(1)) ? erf : erf * e_scale;
        if (lhc) {
            if (!(fgnl < R(1e-15))) S[US_GR0] += gr_c0 * (S[US_LAND0] - S[US_GR0]);
            if (!(fgsl < R(1e-15))) S[US_GR1] += gr_c1 * (S[US_LAND1] - S[US_GR1]);
        }
        udeb_substep<R, N>(y, P, S, tab, q, X, T, F, e * (h ? S[US_QF2] : S[US_QF0]));
        UDEB_CLK(3); // sweeps, rendezvous, back substitution
        // one rendezvous serves two purposes: the new sea-surface temperatures for the land / upwelling updates below,
        // and the edge temperatures (mixed layer, bottom layer) the next sub-step's sweeps start from
        X[(UX_EDGE + q) * 32] = static_cast<double>(T[0]);
        __syncthreads();
        UDEB_CLK(4); // edge rendezvous
        sst_nh = R(X[UX_EDGE * 32]);
        sst_sh = R(X[(UX_EDGE + 1) * 32]);
        const R air_nho = udeb_sst_to_air(air, sst_nh), air_sho = udeb_sst_to_air(air, sst_sh);
        // calculate_land_temperature (mod.rs:352-375) with the year's reciprocal denominators
        S[US_LAND0] = udeb_cap((e * S[US_QF1] * fgnl + kla * air_nho) * inv_den_n, y.tmax);
        S[US_LAND1] = udeb_cap((e * S[US_QF3] * fgsl + kla * air_sho) * inv_den_s, y.tmax);
        if (fgno > R(1e-15)) S[US_HX0] = hx_c0 * (air_sho - air_nho);
        if (fgso > R(1e-15)) S[US_HX1] = hx_c1 * (air_nho - air_sho);
        const R gt = air_nho * fgno + S[US_LAND0] * fgnl + air_sho * fgso + S[US_LAND1] * fgsl;
        S[US_W0] = udeb_floor(w0 * (R(1) - fv * udeb_cap(gt * inv_wt_nh, R(1))), wmin_b);
        S[US_W1] = udeb_floor(w0 * (R(1) - fv * udeb_cap(gt * inv_wt_sh, R(1))), wmin_b);
        UDEB_CLK(5); // land boxes, upwelling
    }
    S[US_AE0] = (r_abs(sst_nh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(air, sst_nh) / sst_nh;
    S[US_AE1] = (r_abs(sst_sh) < R(1e-15)) ? P[U_TA_ALPHA] : udeb_sst_to_air(air, sst_sh) / sst_sh;
    const R st[4] = {udeb_sst_to_air(air, sst_nh), S[US_LAND0], udeb_sst_to_air(air, sst_sh), S[US_LAND1]};
    const R gt = st[0] * fgno + st[1] * fgnl + st[2] * fgso + st[3] * fgsl;
    // (the other roles read this entry next year, after the barriers of the node's entry)
    if (q == 0 && cx.live) cx.scratch[static_cast<long long>(nr.scr + nhist) * SCR_LD] = static_cast<double>(gt * dt_year);
    S[US_NHIST] = R(nhist + 1);
    R f_end[4];
    udeb_apply_efficacy(P, S, erf_end, co2_eff, f_end);
    {
        const R lams[4] = {lam_o, lam_l, lam_o, lam_l};
        R qq = R(0), fb = R(0);
        for (int i = 0; i < 4; ++i) { qq += area[i] * f_end[i]; fb += area[i] * lams[i] * st[i]; }
        out[0] = qq - fb;
    }
    {
        // ocean heat content: every role weighs its rows (the mixed layer is the top roles' row 0), summed through the exchange
        const R rho_c = R(1026.0) * R(3985.0);
        R part = R(0);
        if (NB > NT) {
            if (bottom) part = T[NB - 1];
        }
#pragma unroll
        for (int j = NT - 1; j >= 1; --j) part += T[j];
        part = rho_c * P[U_DZ] * part + rho_c * (bottom ? P[U_DZ] : P[U_MLD]) * T[0];
        X[(UX_OHC + q) * 32] = static_cast<double>(part);
        __syncthreads();
        out[1] = ((R(X[UX_OHC * 32]) + R(X[(UX_OHC + 1) * 32])) + (R(X[(UX_OHC + 2) * 32]) + R(X[(UX_OHC + 3) * 32]))) / R(2);
    }
    out[2] = (sst_nh + sst_sh) / R(2);
    for (int i = 0; i < 4; ++i) out[3 + i] = st[i];
    UDEB_CLK(6); // outputs, heat content
#ifdef RSCM_NODE_CLOCKS
    if (clk_on && cx.N == cx.n_steps - 1)
        printf("udeb_clocks hist %lld lamcalc %lld constants %lld sweeps %lld edge_barrier %lld land %lld outputs %lld (cycles per year)\n",
               udeb_clk[0] / cx.n_steps, udeb_clk[1] / cx.n_steps, udeb_clk[2] / cx.n_steps, udeb_clk[3] / cx.n_steps,
               udeb_clk[4] / cx.n_steps, udeb_clk[5] / cx.n_steps, udeb_clk[6] / cx.n_steps);
#endif
    return ok;
}

} // namespace rscm_dev
