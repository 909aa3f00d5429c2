// magicc_boxes.cuh — scalar box components on the device (one call = one member, one time step, registers only).
//
// Reference: crates/rscm-components/src/components/four_box_ocean_heat_uptake.rs,
// .../ocean_carbon_cycle/ocean_surface_partial_pressure.rs, crates/rscm-magicc/src/carbon/budget.rs:121-205,
// carbon/terrestrial.rs:213-343 (+ parameters/terrestrial_carbon.rs turnover times), chemistry/ch4.rs:55-350,
// chemistry/n2o.rs:171-275.  History-dependent accessors (`previous()`, `at_offset(-k)`) are served from a short
// per-thread ring of the component's own state (S[]), refreshed at the end of every solve.
#pragma once

namespace rscm_dev {

// ---- FourBoxOceanHeatUptake: P = 4 regional ratios; in: ERF|Aggregated; out: FourBox heat uptake ----
template <class R> __device__ __forceinline__ void four_box_ohu_prepare(const R *, R *D) { D[0] = R(0); }
template <class R>
__device__ __forceinline__ bool four_box_ohu_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = in[0] * P[i];
    return true;
}

// ---- OceanSurfacePartialPressure: P = ospp_pi, sensitivity, sst_pi, offsets[5], coefficients[5] ----
template <class R> __device__ __forceinline__ void ocean_surface_pp_prepare(const R *, R *D) { D[0] = R(0); }
template <class R>
__device__ __forceinline__ bool ocean_surface_pp_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    const R d = in[1];
    const R d2 = d * d, d3 = d2 * d, d4 = d2 * d2;
    const R bits[5] = {d, d2 * R(10e-3), -d3 * R(10e-5), d4 * R(10e-7), -d4 * R(10e-10)};
    R dot = R(0);
#pragma unroll
    for (int i = 0; i < 5; ++i) dot += (P[3 + i] + P[8 + i] * P[2]) * bits[i];
    out[0] = (P[0] + dot) * r_exp_call<R>(P[1] * in[0]);
    return true;
}

// ---- CO2Budget: P = gtc_per_ppm, co2_pi; in: fossil, landuse, terrestrial flux, ocean flux, CO2 (state) ----
template <class R> __device__ __forceinline__ void co2_budget_prepare(const R *, R *D) { D[0] = R(0); }
template <class R>
__device__ __forceinline__ bool co2_budget_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &cx, R *, NodeRef)
{
    const R dt = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    const R total_emissions = in[0] + in[1];
    const R net = total_emissions - (in[2] + in[3]);
    out[0] = net;
    out[1] = (total_emissions > R(0)) ? net / total_emissions : R(0);
    out[2] = in[4] + (net * dt) / P[0];
    return true;
}

// ---- TerrestrialCarbon: 4 pools, implicit trapezoid; D = turnover times (plant, detritus, soil, humus), frac_npp_to_soil ----
template <class R> __device__ __forceinline__ void terrestrial_carbon_prepare(const R *P, R *D)
{
    const R frac_npp_to_soil = r_max(R(1) - P[13] - P[14], R(0));
    const R net_flux_plant = P[13] * P[0] - P[12];
    const R tau_plant = (net_flux_plant > R(1e-10)) ? P[8] / net_flux_plant : R(100);
    const R flux_into_det = P[14] * P[0] + P[15] * net_flux_plant;
    const R tau_det = (flux_into_det > R(1e-10)) ? P[9] / flux_into_det : R(3);
    const R flux_det_out = P[9] / tau_det;
    const R flux_into_soil = frac_npp_to_soil * P[0] + (R(1) - P[15]) * net_flux_plant + P[16] * flux_det_out;
    const R tau_soil = (flux_into_soil > R(1e-10)) ? P[10] / flux_into_soil : R(50);
    const R flux_soil_out = P[10] / tau_soil;
    const R flux_into_hum = P[17] * flux_soil_out;
    const R tau_hum = (flux_into_hum > R(1e-10)) ? P[11] / flux_into_hum : R(1000);
    D[0] = tau_plant; D[1] = tau_det; D[2] = tau_soil; D[3] = tau_hum; D[4] = frac_npp_to_soil;
}

template <class R> __device__ __forceinline__ void tc_pool_step(R pool, R tau, R flux_in, R temp_factor, R dt, R &new_pool, R &turnover)
{
    const R k_eff = temp_factor / tau;
    const R half_k = R(0.5) * k_eff * dt;
    new_pool = r_max(((R(1) - half_k) * pool + flux_in * dt) / (R(1) + half_k), R(0));
    turnover = R(0.5) * k_eff * (pool + new_pool);
}

template <class R>
__device__ __forceinline__ bool terrestrial_carbon_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &cx, R *, NodeRef)
{
    const R co2 = in[0], temperature = in[1], landuse = in[2];
    const R dt = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    R fert = R(1);
    if (P[18] != R(0) && !(co2 <= R(0))) fert = r_max(R(1) + P[2] * r_log_call<R>(co2 / P[1]), R(0.1));
    const bool tf = P[19] != R(0);
    const R npp = P[0] * fert * (tf ? r_exp_call<R>(P[3] * temperature) : R(1));
    const R respiration = P[12] * fert * (tf ? r_exp_call<R>(P[4] * temperature) : R(1));
    const R tf_det = tf ? r_exp_call<R>(P[5] * temperature) : R(1);
    const R tf_soil = tf ? r_exp_call<R>(P[6] * temperature) : R(1);
    const R tf_hum = tf ? r_exp_call<R>(P[7] * temperature) : R(1);
    R new_plant, to_plant, new_det, to_det, new_soil, to_soil, new_hum, to_hum;
    tc_pool_step<R>(in[3], D[0], npp * P[13] - respiration - landuse, R(1), dt, new_plant, to_plant);
    tc_pool_step<R>(in[4], D[1], npp * P[14] + P[15] * to_plant, tf_det, dt, new_det, to_det);
    tc_pool_step<R>(in[5], D[2], npp * D[4] + (R(1) - P[15]) * to_plant + P[16] * to_det, tf_soil, dt, new_soil, to_soil);
    tc_pool_step<R>(in[6], D[3], P[17] * to_soil, tf_hum, dt, new_hum, to_hum);
    const R total_resp = respiration + (R(1) - P[16]) * to_det + (R(1) - P[17]) * to_soil + to_hum;
    out[0] = npp - total_resp - landuse;
    out[1] = new_plant; out[2] = new_det; out[3] = new_soil; out[4] = new_hum;
    return true;
}

// ---- CH4Chemistry: 4 Prather iterations; S[0] = CH4 at index N-1 (previous()) ----
template <class R> __device__ __forceinline__ void ch4_chemistry_prepare(const R *P, R *D)
{
    D[0] = R(1) / (R(1) / P[3] + R(1) / P[4] + R(1) / P[5]); // tau_other
}
template <class R> __device__ __forceinline__ void ch4_chemistry_init_state(const R *, const R *, R *S, const StepCtx<R> &, NodeRef) { S[0] = R(0); }

template <class R>
__device__ __forceinline__ bool ch4_chemistry_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef)
{
    const R ch4_current = in[5];
    const R ch4_prev = (cx.N > 0) ? S[0] : ch4_current;
    S[0] = ch4_current;
    const R emissions = in[0], temperature = in[1];
    const R total_emissions = emissions + P[1];
    const R burden_prev = ch4_prev * P[14], burden_ref = P[0] * P[14];
    const R tau_other = D[0];
    R base = P[2];
    if (P[13] != R(0))
        base = P[2] * r_exp_call<R>(-P[7] * (P[8] * (in[2] - P[15]) + P[9] * (in[3] - P[16]) + P[10] * (in[4] - P[17])));
    const R x = -P[7] * P[6];
    R burden = ch4_current * P[14], delta_burden = R(0), tau_oh = P[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const R burden_mean = (burden + burden_prev) / R(2);
        R tau = base * r_pow<R>(r_max(burden_mean / burden_ref, R(1)), x);
        if (i > 0 && !(r_abs(burden_prev) < R(1e-10))) tau = tau * (R(1) - R(0.5) * x * delta_burden / burden_prev);
        if (P[12] != R(0) && !(r_abs(temperature) < R(1e-10))) tau = P[2] / (P[2] / tau + P[11] * r_max(temperature, R(0)));
        delta_burden = total_emissions - burden_mean / tau - burden_mean / tau_other;
        burden = burden_prev + delta_burden;
        tau_oh = tau;
    }
    out[0] = R(1) / (R(1) / tau_oh + R(1) / tau_other);
    out[1] = burden / P[14];
    return true;
}

// ---- N2OChemistry: S[j] = N2O at index N-1-j, j < 8 (previous(), at_offset(-delay), at_offset(-delay-1)) ----
constexpr int N2O_RING = 8;
template <class R> __device__ __forceinline__ void n2o_chemistry_prepare(const R *, R *D) { D[0] = R(0); }
template <class R> __device__ __forceinline__ void n2o_chemistry_init_state(const R *, const R *, R *S, const StepCtx<R> &, NodeRef)
{
#pragma unroll
    for (int j = 0; j < N2O_RING; ++j) S[j] = R(0);
}

template <class R>
__device__ __forceinline__ bool n2o_chemistry_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef)
{
    const R dt = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    const R cur = in[1];
    const int N = cx.N;
    auto back = [&](int k, R fallback) -> R { // value at index N-k (k >= 1) or the fallback when out of range
        R v = fallback;
#pragma unroll
        for (int j = 0; j < N2O_RING; ++j)
            if (j == k - 1 && N - k >= 0) v = S[j];
        return v;
    };
    const R prev = back(1, cur);
    int delay = static_cast<int>(P[4]);
    if (delay < 1) delay = 1;
    const R t_delay = back(delay, prev);
    const R t_delay_m1 = back(delay + 1, t_delay);
    const R lagged = (t_delay + t_delay_m1) / R(2);
#pragma unroll
    for (int j = N2O_RING - 1; j > 0; --j) S[j] = S[j - 1];
    S[0] = cur;
    const R total_emissions = in[0] + P[1];
    const R burden_prev = prev * P[5], burden_lagged = lagged * P[5], burden_ref = P[0] * P[5];
    R burden = cur * P[5], tau_eff = P[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const R mid = (burden_prev + burden) / R(2);
        tau_eff = P[2] * r_pow<R>(r_max(mid / burden_ref, R(1)), P[3]);
        burden = burden_prev + (total_emissions - burden_lagged / tau_eff) * dt;
    }
    out[0] = tau_eff;
    out[1] = burden / P[5];
    return true;
}

// ---- OceanCarbon — carbon/ocean.rs:167-260: monthly air-sea flux + impulse-response convolution --------------------
// The reference re-sums the whole monthly flux history against the (nonlinearly scaled) IRF every month:
// O((12 T)^2 / 2) multiply-adds per run.  Here the scaled IRF is a per-graph table (cx.gtab, lag in months; the IRF
// parameters are per-graph; zero from lag max_history_months on: the reference's bounded deque), the flux history lives
// in the member-interleaved global scratch, and the partial sums of all `steps` months of a year are advanced together
// (steps FMAs per load) in the reference's oldest-to-newest order, so the sums are the same up to FMA contraction.
//   * one thread per member (Prog::LANES == 1): the history that predates the current year is read once per year;
//   * lane groups (a program with ClimateUDEB: four warps = four roles per member, StepCtx): years are taken in
//     blocks of four.  At the first year of a block role q sums the whole history that predates the block against the
//     lags of the block's year q (12 prefix sums, kept in this thread's shared-memory column); role 0 then steps every
//     year from the prefix sums of that year's role and adds the block's own months.  The long history is read once per
//     FOUR years and ONCE PER CTA: the rows of the CTA's 32 members are contiguous in the global scratch (KArgs::scratch),
//     so tiles of OCEAN_KT months (8 KB) are staged in shared memory by bulk copies (cp.async.bulk + mbarrier, two tiles
//     in flight) and the four warps read them from there — a quarter of the L2 traffic of four warps loading the same
//     rows, and none of it through L1, which holds the spilled cells of these large programs.  The accumulation order
//     per month is unchanged (history before the block, then the block, oldest first).
// P: see include/rscm_b200.h (60 values); S[0] = months of history so far, S[1] = tiles staged so far (mbarrier phase).
// Shared memory of the node in lane-group programs: as one CTA-wide region, OCEAN_KT / 2 words x 128 threads are the two
// tiles [2][OCEAN_KT][32], two more words the tiles' IRF windows [2][OCEAN_WIN] (graph.cpp: n_smem_lanes); the tiles' two
// mbarriers are the node's exchange slot.  The prefix sums
// wait for their year in the first 4 x 16 rows of the node's global scratch (row 16 q + m: month m of the block's year q;
// 12 values per member and year through L2), the flux history follows from row OCEAN_HIST0 on.
constexpr int OCEAN_KT = 32;
constexpr int OCEAN_WIN = 128; // staged IRF window per tile: (32 + 4 x 16 - 1) entries at most, + 1 for the even start
constexpr int OCEAN_HIST0 = 64;
template <class R> __device__ __forceinline__ void ocean_carbon_prepare(const R *P, R *D)
{
    D[0] = P[3] / (P[4] * R(12));            // gas_exchange_rate
    D[1] = R(1.72e17) / (P[7] * P[8]);       // dic_conversion_factor
}
template <class R> __device__ __forceinline__ void ocean_carbon_init_state(const R *, const R *, R *S, const StepCtx<R> &cx, NodeRef nr)
{
    S[0] = R(0);
    S[1] = R(0);
    if (cx.lanes == 4 && threadIdx.x == 0) {
        unsigned long long *bars = reinterpret_cast<unsigned long long *>(cx.xch + nr.xch * 32); // (thread 0: lane 0's column = the slot's base)
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
    }
}

template <class R>
__device__ inline bool ocean_carbon_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &cx, R *S, NodeRef nr)
{
    constexpr int MAXS = 16;
    const int steps = nr.aux; // steps_per_year: a literal of the emitted program, so the loops below unroll
    const R co2 = in[0], dsst = in[1];
    R pco2 = in[2], cumulative = in[3];
    const R dt = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    const R dt_month = dt / R(steps);
    const R k_gas = D[0], dic_conv = D[1];
    const int n_old = static_cast<int>(S[0]);
    const double *irf = cx.gtab + nr.gt;
    double *pre = cx.scratch + static_cast<long long>(nr.scr) * SCR_LD;      // prefix sums of the block of four years (lane groups)
    double *hist = pre + OCEAN_HIST0 * SCR_LD;
    R acc[MAXS];
    // sum over history entries [i0, i1) (entry i at src[i * SCR_LD]: the global scratch, or a staged tile, which has the
    // same row length) against the lags of months `first_month + m` (m < steps), oldest entry first.  Chunks of 16 months:
    // sixteen independent loads in flight, and the IRF lags of a chunk form one sliding window of 15 + steps values
    // (uniform loads) instead of steps loads per month.
    // wt[lag]: the IRF table in global memory, or the staged window of a tile (shifted so that the lag indexes it).
    auto convolve = [&](const double *src, int i0, int i1, int first_month, const double *wt) {
        constexpr int CH = 16;
        int i = i0;
        for (; i + CH <= i1; i += CH) {
            R f[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) f[u] = R(src[(i + u) * SCR_LD]);
            const int l0 = first_month - i - (CH - 1); // wv[k]: lag l0 + k
            R wv[CH - 1 + MAXS];
#pragma unroll
            for (int k = 0; k < CH - 1 + MAXS; ++k)
                if (k < CH - 1 + steps) wv[k] = R(wt[l0 + k]);
#pragma unroll
            for (int u = 0; u < CH; ++u)
#pragma unroll
                for (int m = 0; m < MAXS; ++m)
                    if (m < steps) acc[m] += f[u] * wv[(CH - 1 - u) + m];
        }
        for (; i < i1; ++i) {
            const R f = R(src[i * SCR_LD]);
#pragma unroll
            for (int m = 0; m < MAXS; ++m)
                if (m < steps) acc[m] += f * R(wt[first_month - i + m]);
        }
    };
    // a flux older than max_history_months has left the reference's deque: no need to load it (its weights are zero)
    const int max_hist = static_cast<int>(P[11]);
    auto oldest = [&](int first_month) { const int lo = first_month + 1 - max_hist; return lo > 0 ? (lo < n_old ? lo : n_old) : 0; };
#ifdef RSCM_NODE_CLOCKS
    __shared__ long long ocean_clk[5];
    const bool clk_on = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    if (clk_on && cx.N == 0) for (int i = 0; i < 5; ++i) ocean_clk[i] = 0;
    long long clk_t = clock64();
#define OCEAN_CLK(i) do { if (clk_on) { const long long now_ = clock64(); ocean_clk[i] += now_ - clk_t; clk_t = now_; } } while (0)
#else
#define OCEAN_CLK(i) do { } while (0)
#endif
    if (cx.lanes == 4) {
        const int yb = (n_old / steps) & 3;                               // year within the block of four (CTA-uniform)
        if (yb == 0) {
            const int first_month = n_old + steps * cx.role;              // role q prepares year q of the block
#pragma unroll
            for (int m = 0; m < MAXS; ++m) acc[m] = R(0);
            double *cta = reinterpret_cast<double *>(cx.sm - threadIdx.x);
            double *tiles = cta + nr.sm * BLOCK;
            unsigned long long *bars = reinterpret_cast<unsigned long long *>(cx.xch - (threadIdx.x & 31) + nr.xch * 32);
            const int lo_cta = oldest(n_old), lo_role = oldest(first_month); // role 0 reaches furthest back
            const int nt = (n_old - lo_cta + OCEAN_KT - 1) / OCEAN_KT;
            const int tc = static_cast<int>(S[1]);
            // the IRF lags the four roles need for the months [ts, te) of a tile: one window of (te - ts) + 4 steps - 1 table
            // entries from lag n_old - te + 1 on, staged next to the tile (start rounded down to an even entry: 16-byte copies)
            double *wins = tiles + 2 * OCEAN_KT * SCR_LD; // [2][OCEAN_WIN]
            auto win0 = [&](int te) { return (n_old - te + 1) & ~1; };
            auto stage = [&](int t) { // thread 0 (lane 0 of the CTA's block: its `hist` is the block's row base)
                const int ts = lo_cta + t * OCEAN_KT, te = ts + OCEAN_KT < n_old ? ts + OCEAN_KT : n_old;
                const unsigned bytes = static_cast<unsigned>((te - ts) * SCR_LD * 8);
                const int w0 = win0(te);
                const unsigned wbytes = static_cast<unsigned>(((n_old + 4 * steps - ts - w0 + 1) & ~1) * 8);
                void *bar = bars + ((tc + t) & 1);
                mbar_expect_tx(bar, bytes + wbytes);
                tma_bulk_g2s_stream(tiles + ((tc + t) & 1) * OCEAN_KT * SCR_LD, hist + static_cast<long long>(ts) * SCR_LD, bytes, bar);
                tma_bulk_g2s(wins + ((tc + t) & 1) * OCEAN_WIN, irf + w0, wbytes, bar);
            };
            if (threadIdx.x == 0) {
                if (nt > 0) stage(0);
                if (nt > 1) stage(1);
            }
            for (int t = 0; t < nt; ++t) {
                const int k = tc + t, ts = lo_cta + t * OCEAN_KT, te = ts + OCEAN_KT < n_old ? ts + OCEAN_KT : n_old;
                mbar_wait(bars + (k & 1), (k >> 1) & 1);
                OCEAN_CLK(0); // waiting for a tile
                const double *win = wins + (k & 1) * OCEAN_WIN - win0(te); // win[lag]
                convolve(tiles + (k & 1) * OCEAN_KT * SCR_LD + cx.col - ts * SCR_LD, ts > lo_role ? ts : lo_role, te, first_month, win);
                OCEAN_CLK(1); // the tile's multiply-adds
                __syncthreads(); // every warp is done with this tile: its buffer can take the tile after next
                if (threadIdx.x == 0 && t + 2 < nt) stage(t + 2);
                OCEAN_CLK(2); // rendezvous of the four roles
            }
            S[1] = R(tc + nt);
            if (cx.live) {
#pragma unroll
                for (int m = 0; m < MAXS; ++m)
                    if (m < steps) pre[(16 * cx.role + m) * SCR_LD] = static_cast<double>(acc[m]);
            }
        }
        __syncthreads(); // (also makes the prefix sums, written by the member's other roles, visible)
        OCEAN_CLK(3); // prefix sums out, rendezvous
        // The block's own months so far (up to three years, just written by role 0: L2) are added by all four roles, four
        // of the year's sums each (months 4 q .. 4 q + 3, in the same order: history before the block, then the block,
        // oldest first), and handed to role 0.
        {
            const int m0 = 4 * cx.role;
            R part[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) part[u] = (m0 + u < steps) ? R(pre[(16 * yb + m0 + u) * SCR_LD]) : R(0); // the sums role yb prepared
#pragma unroll 4
            for (int i = n_old - yb * steps; i < n_old; ++i) {
                const R f = R(hist[i * SCR_LD]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (m0 + u < steps) part[u] += f * R(irf[n_old - i + m0 + u]);
            }
            // (handed over through the tile area, idle outside the staging above; like every per-thread shared-memory word it
            // holds nothing beyond this solve: ClimateUDEB's column uses the same words)
            double *xa = reinterpret_cast<double *>(cx.sm - threadIdx.x) + nr.sm * BLOCK + (threadIdx.x & 31); // sum m at xa[m * 32]
#pragma unroll
            for (int u = 0; u < 4; ++u) xa[(m0 + u) * 32] = static_cast<double>(part[u]);
            __syncthreads();
            if (cx.role != 0) { // the other roles only help with the history; role 0 steps the months
                S[0] = R(n_old + steps);
                return true;
            }
#pragma unroll
            for (int m = 0; m < MAXS; ++m) acc[m] = (m < steps) ? R(xa[m * 32]) : R(0);
        }
        OCEAN_CLK(3); // (+ the block's own months)
    } else {
#pragma unroll
        for (int m = 0; m < MAXS; ++m) acc[m] = R(0);
        convolve(hist, oldest(n_old), n_old, n_old, irf);
    }
    R fy[MAXS];
    R total_flux = R(0);
    const R tfac = (P[59] != R(0)) ? r_exp_call<R>(P[5] * dsst) : R(1);
#pragma unroll
    for (int m = 0; m < MAXS; ++m) {
        if (m < steps) {
            const R flux_ppm = k_gas * (co2 - pco2);
            if (cx.live) __stcs(hist + (n_old + m) * SCR_LD, static_cast<double>(flux_ppm)); // streaming: next read in 1..4 years, after ~1 GB of other traffic
            fy[m] = flux_ppm;
            const R flux_gtc_yr = flux_ppm * R(12) * R(2.124);
            total_flux += flux_gtc_yr / R(steps);
            cumulative += flux_gtc_yr * dt_month;
            R integral = acc[m];
#pragma unroll
            for (int j = 0; j < MAXS; ++j)
                if (j <= m) integral += fy[j] * R(__ldg(irf + (m - j)));
            const R ddic = integral * dic_conv;
            const R d2 = ddic * ddic, d3 = d2 * ddic, d4 = d2 * d2, d5 = d4 * ddic;
            const R pw[5] = {ddic, d2 * R(1e-3), -d3 * R(1e-5), d4 * R(1e-7), -d5 * R(1e-10)};
            R dp = R(0);
#pragma unroll
            for (int q = 0; q < 5; ++q) dp += (P[49 + q] + P[54 + q] * P[9]) * pw[q];
            pco2 = (P[2] + dp) * tfac;
        }
    }
    OCEAN_CLK(4); // role 0: the year's own months
#ifdef RSCM_NODE_CLOCKS
    if (clk_on && cx.N == cx.n_steps - 1)
        printf("ocean_clocks tile_wait %lld multiply_adds %lld tile_rendezvous %lld prefix_out %lld months_of_the_year %lld (cycles per year)\n",
               ocean_clk[0] / cx.n_steps, ocean_clk[1] / cx.n_steps, ocean_clk[2] / cx.n_steps, ocean_clk[3] / cx.n_steps, ocean_clk[4] / cx.n_steps);
#endif
    if (cx.lanes == 4) fence_proxy_async(); // the months just written are staged by bulk copies from the next block of years on
    S[0] = R(n_old + steps);
    out[0] = total_flux;
    out[1] = pco2;
    out[2] = cumulative;
    return true;
}

// ---- HalocarbonChemistry — crates/rscm-magicc/src/chemistry/halocarbon.rs -------------------------------------------
// decay_species :115-134, species_forcing :137-145, calculate_{total,fgas,montreal}_forcing :148-196, calculate_eesc
// :204-225, solve :295-352.  41 species (the reference's default list), all parameters per-graph: the per-species
// table {lifetime, conv, radiative_efficiency, concentration_pi, halogen loading, normalised release} comes from
// the host (graph.cpp halocarbon_const_table) through shared memory.
// in:  per species [emissions, concentration] as InputState::get_global returns them; out: 41 concentrations,
// Forcing|Halocarbons, Forcing|F-gases, Forcing|Montreal Gases, EESC.
// S[s] = the species' last non-NaN concentration: get_global on an endogenous series is Timeseries::latest_value
// (state/mod.rs:231-254), so a NaN written by one step (NaN emissions) does not stick.
constexpr int HALO_NS = 41, HALO_NF = 23, HALO_CT = 6;
constexpr int HALOCARBON_CHEMISTRY_NP = 0;
constexpr int HALOCARBON_CHEMISTRY_ND = 0;

template <class R> __device__ __forceinline__ void halocarbon_chemistry_prepare(const R *, R *) {}

template <class R> __device__ inline void halocarbon_chemistry_init_state(const R *, const R *, R *S, const StepCtx<R> &, NodeRef)
{
#pragma unroll
    for (int s = 0; s < HALO_NS; ++s) S[s] = r_nan<R>();
    S[HALO_NS] = R(0);
}

// One species, one step (decay_species :115-134, species_forcing :137-145)
template <class R>
__device__ __forceinline__ void halocarbon_species(const double *tab, int s, bool uniform_dt, R dt, R emissions, R conc_in, R &good, R &new_conc,
                                                   R &forcing)
{
    const double *t = tab + HALO_CT * s;
    const R lifetime = R(t[0]), conv = R(t[1]), rad_eff = R(t[2]), conc_pi = R(t[3]);
    R conc = conc_in;
    if (conc != conc) conc = good; // latest_value: fall back to the last non-NaN value (NaN if there never was one)
    good = conc;
    const R decay = uniform_dt ? R(tab[HALO_CT * HALO_NS + 1 + s]) : r_exp_call<R>(-dt / lifetime); // block-uniform choice
    const R emissions_ppt = emissions * conv;
    new_conc = conc * decay + emissions_ppt * lifetime * (R(1) - decay);
    forcing = (new_conc - conc_pi) * rad_eff / R(1000);
}

template <class R, class Out>
__device__ inline bool halocarbon_chemistry_solve(const R *, const R *, const R *in, Out out, const StepCtx<R> &cx, R *S, NodeRef nr)
{
    const double *tab = cx.ctab + nr.ctab;
    const R dt = R(cx.bounds[cx.N + 1] - cx.bounds[cx.N]);
    R total = R(0), fgas = R(0), montreal = R(0), eesc = R(0);
    // on a uniform time axis the decay factors are constants of the graph (host-computed, graph.cpp)
    const bool uniform_dt = static_cast<double>(dt) == tab[HALO_CT * HALO_NS];
    // Lane-group programs, all emissions exogenous (nr.aux): the 32 members of the warp belong to one scenario, so unless a
    // species' initial concentration is bound per member they all compute the same 41 species.  Then lane l computes species
    // l and l + 32 only (its own last-good values in S[0], S[1]), and every lane collects the results by shuffles, adding the
    // forcings in species order — the same operations on the same values as the member-by-member form, at a fraction of
    // its 200 local-memory operations per member-year.  Decided at the first step by comparing the lanes' concentrations.
    if (cx.lanes > 1 && nr.aux != 0) {
        const int lane = static_cast<int>(threadIdx.x) & 31;
        if (cx.N == 0) {
            bool same = true;
#pragma unroll 1
            for (int s = 0; s < HALO_NS; ++s) {
                const R c = in[2 * s + 1];
                const R c0 = __shfl_sync(0xffffffffu, c, 0); // (every lane takes part in every shuffle: not under the && below)
                same = same && c == c0;                      // (a NaN compares unequal: member-by-member form)
            }
            S[HALO_NS] = __all_sync(0xffffffffu, same) ? R(1) : R(0);
        }
        if (S[HALO_NS] != R(0)) {
#ifdef RSCM_NODE_CLOCKS
            __shared__ long long halo_clk[3];
            const bool clk_on = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
            if (clk_on && cx.N == 0) halo_clk[0] = halo_clk[1] = halo_clk[2] = 0;
            long long clk_t = clock64();
#define HALO_CLK(i) do { if (clk_on) { const long long now_ = clock64(); halo_clk[i] += now_ - clk_t; clk_t = now_; } } while (0)
#else
#define HALO_CLK(i) do { } while (0)
#endif
            const int sa = lane, sb = lane + 32 < HALO_NS ? lane + 32 : HALO_NS - 1; // (lanes 9.. repeat the last species, unused)
            R nca, fa, ncb, fb;
            {
                const R ea = in[2 * sa], ca = in[2 * sa + 1], eb = in[2 * sb], cb = in[2 * sb + 1];
                halocarbon_species<R>(tab, sa, uniform_dt, dt, ea, ca, S[0], nca, fa);
                halocarbon_species<R>(tab, sb, uniform_dt, dt, eb, cb, S[1], ncb, fb);
                // (the forcings are finished here, once per species: without this the compiler shuffles the numerators
                // and divides by 1000 on every lane of every turn of the loop below)
                r_keep(fa);
                r_keep(fb);
            }
            HALO_CLK(0); // this lane's two species
#pragma unroll 1
            for (int s = 0; s < HALO_NS; ++s) {
                const R new_conc = __shfl_sync(0xffffffffu, s < 32 ? nca : ncb, s & 31);
                const R forcing = __shfl_sync(0xffffffffu, s < 32 ? fa : fb, s & 31);
                const R loading = R(tab[HALO_CT * s + 4]), release = R(tab[HALO_CT * s + 5]);
                out[s] = new_conc;
                total += forcing;
                if (s < HALO_NF) fgas += forcing;
                else montreal += forcing;
                if (release > R(0)) eesc += new_conc * loading * release;
            }
            out[HALO_NS] = total;
            out[HALO_NS + 1] = fgas;
            out[HALO_NS + 2] = montreal;
            out[HALO_NS + 3] = eesc;
            HALO_CLK(1); // collecting the 41 species and the ordered sums
#ifdef RSCM_NODE_CLOCKS
            if (clk_on && cx.N == cx.n_steps - 1)
                printf("halocarbon_clocks own_species %lld collect %lld (cycles per year)\n", halo_clk[0] / cx.n_steps, halo_clk[1] / cx.n_steps);
#endif
            return true;
        }
    }
    // ROLLED over the 41 species (unrolled it is 3.5 k instructions of a program that is bound by instruction fetch), in
    // blocks of eight: the dynamic indices put in / out / S of this component into local memory, which in these programs
    // is served by L2 (shared memory takes most of the SM's L1), so the 24 loads of a block are issued together, before any
    // of the block's arithmetic and branches.  The sums keep the species order.
    constexpr int HB = 8;
#pragma unroll 1
    for (int s0 = 0; s0 < HALO_NS; s0 += HB) {
        R em[HB], cc[HB], sv[HB];
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const int s = s0 + u < HALO_NS ? s0 + u : HALO_NS - 1;
            em[u] = in[2 * s];
            cc[u] = in[2 * s + 1];
            sv[u] = S[s];
        }
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            const int s = s0 + u;
            if (s < HALO_NS) {
                R new_conc, forcing;
                halocarbon_species<R>(tab, s, uniform_dt, dt, em[u], cc[u], sv[u], new_conc, forcing);
                S[s] = sv[u];
                out[s] = new_conc;
                const R loading = R(tab[HALO_CT * s + 4]), release = R(tab[HALO_CT * s + 5]);
                total += forcing;
                if (s < HALO_NF) fgas += forcing;
                else montreal += forcing;
                if (release > R(0)) eesc += new_conc * loading * release;
            }
        }
    }
    out[HALO_NS] = total;
    out[HALO_NS + 1] = fgas;
    out[HALO_NS + 2] = montreal;
    out[HALO_NS + 3] = eesc;
    return true;
}

} // namespace rscm_dev
