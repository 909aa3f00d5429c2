// components.cuh — device-side physics of the component kinds on the ensemble
// hot path.  One call = one component solve for one member and one time step;
// all operands live in registers.  Included by the fused member-loop kernel
// (kernel.cuh) through the program structs the graph compiler emits.
//
// Calling convention (what graph.cpp's emitter generates):
//   <kind>_prepare<R>(P, D)            once per member: derived constants
//   <kind>_solve<R>(P, D, in, out, ..) once per step; `in` / `out` follow the
//       reference's inputs() / outputs() order (definitions order = inputs,
//       outputs, states: crates/rscm-macros/src/lib.rs:630-636), regions
//       expanded; returns false when the reference's solve would fail
//       (outputs then stay NaN, model/runtime.rs:493-495).
//
// Arithmetic follows the reference expression by expression except where noted
// ("hoisted" / "reciprocal"): those re-associations change results by O(1 ulp)
// per operation, far inside the 1e-9 relative parity bar (fp64).
#pragma once
#include "math_tables.cuh"

namespace rscm_dev {

constexpr int BLOCK = 128; // threads per CTA = stride of the per-thread shared-memory scratch
constexpr int SCR_LD = 32;  // global scratch: runs are blocked by 32, row j of a block at j * SCR_LD (KArgs::scratch)

// ---- mbarrier / TMA bulk-copy primitives (PTX) -----------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// The same copy with an L2 evict-first policy: for data that streams through once per pass and is far larger than L2 (the
// flux history of OceanCarbon), so that it does not push out the lines that are re-used (the spilled cells of the large
// programs live in local memory, i.e. in L2).
__device__ __forceinline__ void tma_bulk_g2s_stream(void *dst, const void *src, unsigned bytes, void *bar)
{
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// Global memory written with ordinary stores and later read by a bulk copy (the async proxy): the writer fences, then
// synchronises with the thread that issues the copy.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// What a component's solve sees beyond its own operands.
template <class R> struct StepCtx {
    const int *nsub;      // RK4 sub-step tables [n_rk][Tpad] (shared memory)
    const double *bounds; // time bounds (shared memory; valid when Prog::NEEDS_TIME)
    const double *ctab;   // per-graph constant tables (shared memory)
    const double *gtab;   // large per-graph constant tables (global memory, read through the read-only path)
    R *sm;                // this thread's shared-memory scratch: element j at sm[j * BLOCK]
    double *scratch;      // this run's global scratch: element j at scratch[j * SCR_LD] (blocks of 32 runs, see KArgs::scratch)
    int col;              // this run's column in its block of the global scratch (= member & 31)
    int Tpad;
    int N;                // current time index
    int n_steps;          // steps of the run (T - 1)
    // Lane groups (Prog::LANES == 4: programs with ClimateUDEB).  Each warp of the CTA is one ROLE of the CTA's 32
    // members: lane l of every warp belongs to member l.  Role 0 runs the component graph thread-per-member; the kinds
    // that spread a member over four threads (ClimateUDEB's half-sweeps, OceanCarbon's history sums) are entered by all
    // four warps, which exchange through `xch` (shared memory) and __syncthreads().
    int role;             // this thread's role = (warp + rot) & 3 (0 when Prog::LANES == 1)
    int rot;              // role rotation of this CTA; role r of a member is thread ((r - rot) & 3) * 32 + lane
    int lanes;            // Prog::LANES
    bool live;            // false for the padding members of the last CTA: they compute (barriers need every warp) but
                          // must not write global scratch
    double *xch;          // this member's exchange column: slot j at xch[j * 32]
};

// Output view for kinds with dozens of outputs (KindInfo::scatter_out): out[i] is the cell of the next time level the
// i-th output value belongs to, so the solve writes its results where they live.  The first LIN_N outputs sit at cells
// LIN_A + LIN_B i (a schema lists a component's series in order), which keeps the cell of a run-time index arithmetic; the
// others are addressed with compile-time indices, so their table look-up folds away.
template <class R, int LIN_N, int LIN_A, int LIN_B> struct ScatterOut {
    R *cells;
    const short *cell_of;
    __device__ __forceinline__ R &operator[](int i) const { return i < LIN_N ? cells[LIN_A + LIN_B * i] : cells[cell_of[i]]; }
};

// Per-node literals the emitter passes: RK4 table row, offsets into ctab / sm / scratch / gtab, one kind-specific
// integer (e.g. OceanCarbon's steps_per_year) that must be a compile-time constant, and the first exchange slot the
// kind may use (lane groups; the slots before it carry the node's broadcast inputs).
struct NodeRef {
    int rk, ctab, sm, scr, gt, aux, xch;
};

// ---------------------------------------------------------------------------
// exp / log in fp64.  The CUDA library versions materialise every polynomial coefficient with two uniform-register
// moves at the point of use (about 50 issue slots per exp + log pair in the member loop, which is bound by issue slots:
// an FP64 instruction takes two) and spend about 40 FP64 instructions on the pair.  These take their coefficients from
// the constant bank and are table-driven (math_tables.cuh: 2^(j/64) for exp, {c, 1/c, log c} on a 1/256 grid for log), so
// that degree-5 / degree-6 polynomials suffice: 10 + 13 FP64 instructions.  The rare arguments outside the plain
// range (|x| >= 700; zero, negative, subnormal, infinite or NaN y) go to the library functions.  Error <= 1 ulp against
// the exact value on the fast paths (tests/test_device_math.py compares both with the host library, which is within an
// ulp itself, over the ranges the components use).
// ---------------------------------------------------------------------------
// {64/ln2, ln2/64 (31 significant bits), its remainder, 1/2, 1/6, 1/24, 1/120}
__constant__ double RSCM_EXP_C[7] = {92.33248261689366, 0.01083042468962958, 6.619564634077006e-12, 0.5,
                                     0.16666666666666666, 0.041666666666666664, 0.0083333333333333332};
// {ln2 (31 significant bits), its remainder, -1/2, 1/3, -1/4, 1/5, -1/6}
__constant__ double RSCM_LOG_C[7] = {0.6931471801362932, 4.236521365809284e-10, -0.5, 0.33333333333333331, -0.25, 0.20000000000000001,
                                     -0.16666666666666666};

// the library versions, out of line: they are the rare path and must not bloat the member loop
__device__ __noinline__ double rscm_exp_slow(double x) { return exp(x); }
__device__ __noinline__ double rscm_log_slow(double y) { return log(y); }

__device__ __forceinline__ double rscm_exp(double x)
{
    const int hx = __double2hiint(x);
    if ((hx & 0x7fffffff) >= 0x4085e000) return rscm_exp_slow(x); // |x| >= 700, inf, NaN: scaling below would leave the normal range
    const double magic = 6755399441055744.0;           // 1.5 * 2^52: round to nearest integer in the low word
    const double t = fma(x, RSCM_EXP_C[0], magic);
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(kd, -RSCM_EXP_C[1], x);
    r = fma(kd, -RSCM_EXP_C[2], r);                     // |r| <= ln2/128
    double q = fma(r, RSCM_EXP_C[6], RSCM_EXP_C[5]);
    q = fma(r, q, RSCM_EXP_C[4]);
    q = fma(r, q, RSCM_EXP_C[3]);
    const double p = fma(r * r, q, r);                  // e^r - 1, truncation r^6/720 < 4e-17
    const double tj = __ldg(RSCM_EXP_TAB + (k & 63));         // per-lane index: through L1, not the constant bank
    const double v = fma(tj, p, tj);                    // in [1, 2 (1 + 2^-7))
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}

// log y = e ln2 + log c + log(1 + u): m = y 2^-e in [sqrt(1/2), sqrt(2)), c = the nearest multiple of 1/256 (m - c is
// exact), u = (m - c)/c with |u| < 2^-8.5, so that u - u^2/2 + ... - u^6/6 is exact to 2^-54 relative; log c comes as a
// high and a low part from the table (math_tables.cuh; per-lane indices: global memory through L1, not the constant
// bank, which would serialise them).  The cell c = 1 has log c = 0 and 1/c = 1: log is exact around 1 and log 1 = 0.
__device__ __forceinline__ double rscm_log(double y)
{
    int hy = __double2hiint(y);
    if (static_cast<unsigned>(hy - 0x00100000) >= 0x7fe00000u) return rscm_log_slow(y); // zero, negative, subnormal, inf, NaN
    int e = (hy >> 20) - 1023;
    hy = (hy & 0x000fffff) | 0x3ff00000;
    if (hy >= 0x3ff6a09f) { hy -= 0x00100000; e += 1; } // m in [sqrt(1/2), sqrt(2))
    const double m = __hiloint2double(hy, __double2loint(y));
    const int j = __double2loint(fma(m, 256.0, 6755399441055744.0)) - RSCM_LOG_J0;
    const double2 t01 = __ldg(reinterpret_cast<const double2 *>(RSCM_LOG_TAB[j]));
    const double2 t23 = __ldg(reinterpret_cast<const double2 *>(RSCM_LOG_TAB[j]) + 1);
    const double u = (m - t01.x) * t01.y;
    double p = fma(u, RSCM_LOG_C[6], RSCM_LOG_C[5]);
    p = fma(u, p, RSCM_LOG_C[4]);
    p = fma(u, p, RSCM_LOG_C[3]);
    p = fma(u, p, RSCM_LOG_C[2]);
    const double r = fma(u * u, p, u);
    const double ed = static_cast<double>(e);
    return fma(ed, RSCM_LOG_C[0], t23.x + (r + fma(ed, RSCM_LOG_C[1], t23.y)));
}

// Out-of-line copies for the components of large programs (the MAGICC boxes): a call is a handful of instructions where
// the inlined function is forty, and those programs are bound by instruction fetch.  The two-component
// programs of the headline path keep the inlined versions.
__device__ __noinline__ double rscm_exp_call(double x) { return rscm_exp(x); }
__device__ __noinline__ double rscm_log_call(double y) { return rscm_log(y); }
template <class R> __device__ __forceinline__ R r_exp_call(R x);
template <> __device__ __forceinline__ double r_exp_call<double>(double x) { return rscm_exp_call(x); }
template <> __device__ __forceinline__ float r_exp_call<float>(float x) { return expf(x); }
template <class R> __device__ __forceinline__ R r_log_call(R x);
template <> __device__ __forceinline__ double r_log_call<double>(double x) { return rscm_log_call(x); }
template <> __device__ __forceinline__ float r_log_call<float>(float x) { return logf(x); }

template <class R> __device__ __forceinline__ R r_exp(R x);
template <> __device__ __forceinline__ double r_exp<double>(double x) { return rscm_exp(x); }
template <> __device__ __forceinline__ float r_exp<float>(float x) { return expf(x); }
template <class R> __device__ __forceinline__ R r_log(R x);
template <> __device__ __forceinline__ double r_log<double>(double x) { return rscm_log(x); }
template <> __device__ __forceinline__ float r_log<float>(float x) { return logf(x); }
template <class R> __device__ __forceinline__ R r_sqrt(R x);
template <> __device__ __forceinline__ double r_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float r_sqrt<float>(float x) { return sqrtf(x); }
template <class R> __device__ __forceinline__ R r_pow(R x, R y);
// pow.  The library version is about three hundred instructions per call; the MAGICC chain makes fourteen calls per model
// year (the Prather iterations of CH4 / N2O, the CH4-N2O overlap terms, stratospheric ozone), most of them one after the
// other.  x^y = exp(y log x) from the tables above: log x is kept as an unevaluated sum hi + lo (the table's high part
// plus e ln2 with their rounding error recovered; everything else in lo: absolute error about 2^-61), y (hi + lo) likewise,
// and the part of the exponent that rounding to a double drops is applied to the result, e^(s + sl) = e^s (1 + sl).  About
// 1 ulp (tests/test_device_math.py: <= 2 ulp against the host library).  Zero, negative, subnormal, infinite or NaN x, a
// non-finite y and |y log x| >= 700 go to the library function.  One out-of-line copy: these programs are far larger than
// the instruction cache.
__device__ __noinline__ double rscm_pow_slow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double rscm_pow(double x, double y)
{
    int hx = __double2hiint(x);
    if (static_cast<unsigned>(hx - 0x00100000) >= 0x7fe00000u || (__double2hiint(y) & 0x7fffffff) >= 0x7ff00000) return rscm_pow_slow(x, y);
    int e = (hx >> 20) - 1023;
    hx = (hx & 0x000fffff) | 0x3ff00000;
    if (hx >= 0x3ff6a09f) { hx -= 0x00100000; e += 1; } // m in [sqrt(1/2), sqrt(2))
    const double m = __hiloint2double(hx, __double2loint(x));
    const int j = __double2loint(fma(m, 256.0, 6755399441055744.0)) - RSCM_LOG_J0;
    const double2 t01 = __ldg(reinterpret_cast<const double2 *>(RSCM_LOG_TAB[j]));
    const double2 t23 = __ldg(reinterpret_cast<const double2 *>(RSCM_LOG_TAB[j]) + 1);
    const double u = (m - t01.x) * t01.y;
    double p = fma(u, RSCM_LOG_C[6], RSCM_LOG_C[5]);
    p = fma(u, p, RSCM_LOG_C[4]);
    p = fma(u, p, RSCM_LOG_C[3]);
    p = fma(u, p, RSCM_LOG_C[2]);
    const double r = fma(u * u, p, u);
    const double ed = static_cast<double>(e);
    const double a = ed * RSCM_LOG_C[0];                // exact: 11 bits x 31 bits
    const double hi = a + t23.x;
    const double bb = hi - a;
    const double lo = ((a - (hi - bb)) + (t23.x - bb)) + (r + fma(ed, RSCM_LOG_C[1], t23.y));
    const double ph = y * hi;
    const double pl = fma(y, hi, -ph) + y * lo;
    const double sh = ph + pl;
    if (!(fabs(sh) < 700.0)) return rscm_pow_slow(x, y);
    const double sl = (ph - sh) + pl;
    const double ex = rscm_exp(sh);
    return fma(ex, sl, ex);
}
template <> __device__ __forceinline__ double r_pow<double>(double x, double y) { return rscm_pow(x, y); }
template <> __device__ __forceinline__ float r_pow<float>(float x, float y) { return powf(x, y); }
// an optimisation barrier for one value: what follows sees it as computed here
__device__ __forceinline__ void r_keep(double &x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void r_keep(float &x) { asm volatile("" : "+f"(x)); }
template <class R> __device__ __forceinline__ R r_nan();
template <> __device__ __forceinline__ double r_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <> __device__ __forceinline__ float r_nan<float>() { return __int_as_float(0x7fc00000); }

// ---------------------------------------------------------------------------
// TwoLayer — crates/rscm-two-layer/src/component.rs
//   P: lambda0, a, efficacy, eta, heat_capacity_surface, heat_capacity_deep
//   D: (h/6) k1, (h/6) k2, (h/6) k3 (below), (h/6)/Cs, (h/6) eta/Cd   (h = 0.1, the fixed RK4 step)
//   in : erf (get()), Ts (at_start), Td (at_start)      out: Ts, Td
// RK4 (ode_solvers 0.6.1 as called from rscm-core/src/ivp/mod.rs:245-253):
// nsub fixed steps of h = 0.1 (component.rs:240); the third state (cumulative
// heat, y0[2]=0, component.rs:236,186-187) is integrated and discarded by the
// reference, so it is not computed here.
// ---------------------------------------------------------------------------
constexpr int TWO_LAYER_NP = 6;
constexpr int TWO_LAYER_ND = 5;

// With x = efficacy*eta (same association as component.rs:176) the surface equation
//   dTs = (F - (lambda0 - a Ts) Ts - x (Ts - Td)) / Cs
// is evaluated in the expanded form  F/Cs + Ts (k1 + k3 Ts) + k2 Td  with
//   k1 = -(lambda0 + x)/Cs, k2 = x/Cs, k3 = a/Cs      (3 FMAs per evaluation)
// and dTd = (Ts - Td) * (eta/Cd).  Same polynomial, different association: O(1 ulp).
// The fixed RK4 step h = 0.1 (component.rs:238-243) is folded into the coefficients as h/6 (two_layer_prepare).
template <class R>
__device__ __forceinline__ void two_layer_prepare(const R *P, R *D)
{
    // h/6 is folded in: a right-hand side returns (h/6) f, the RK4 stage points are y + 3 K, y + 3 K, y + 6 K and the
    // update is y + (K0 + 2 K1 + 2 K2 + K3) — every weight is an instruction immediate (1/6 is not: it would be
    // re-materialised with two moves wherever it is used)
    const R h6 = R(0.1) / R(6);
    const R x = P[2] * P[3];
    const R inv_cs = R(1) / P[4];
    D[0] = -(P[0] + x) * inv_cs * h6;
    D[1] = x * inv_cs * h6;
    D[2] = P[1] * inv_cs * h6;
    D[3] = inv_cs * h6;
    D[4] = P[3] / P[5] * h6;
}

template <class R>
__device__ __forceinline__ void two_layer_rhs(R k0, R k1, R k2, R k3, R eta_cd, R ts, R td, R &dts, R &dtd)
{
    // calculate_dy_dt — component.rs:160-188 (times h/6)
    dts = ts * (k3 * ts + k1) + (td * k2 + k0);
    dtd = (ts - td) * eta_cd;
}

template <class R>
__device__ __forceinline__ void two_layer_rk4_step(R k0, R k1, R k2, R k3, R eta_cd, R &ts, R &td)
{
    R a0, b0, a1, b1, a2, b2, a3, b3;
    two_layer_rhs(k0, k1, k2, k3, eta_cd, ts, td, a0, b0);
    two_layer_rhs(k0, k1, k2, k3, eta_cd, ts + a0 * R(3), td + b0 * R(3), a1, b1);
    two_layer_rhs(k0, k1, k2, k3, eta_cd, ts + a1 * R(3), td + b1 * R(3), a2, b2);
    two_layer_rhs(k0, k1, k2, k3, eta_cd, ts + a2 * R(6), td + b2 * R(6), a3, b3);
    ts = ts + ((a0 + a1 * R(2)) + (a2 * R(2) + a3));
    td = td + ((b0 + b1 * R(2)) + (b2 * R(2) + b3));
}

template <class R>
__device__ __forceinline__ bool two_layer_solve(const R *, const R *D, const R *in, R *out, const StepCtx<R> &cx, R *, NodeRef nr)
{
    const int nsub = cx.nsub[nr.rk * cx.Tpad + cx.N];
    if (nsub < 0) return false; // get_last_step assertion (ivp/mod.rs:94-97) would fire
    const R k1 = D[0], k2 = D[1], k3 = D[2], eta_cd = D[4];
    const R k0 = in[0] * D[3]; // (h/6) F / Cs, F frozen over the step
    R ts = in[1], td = in[2];
    if (nsub == 10) { // annual steps: straight-line code (block-uniform branch)
#pragma unroll
        for (int s = 0; s < 10; ++s) two_layer_rk4_step(k0, k1, k2, k3, eta_cd, ts, td);
    } else {
        for (int s = 0; s < nsub; ++s) two_layer_rk4_step(k0, k1, k2, k3, eta_cd, ts, td);
    }
    out[0] = ts;
    out[1] = td;
    return true;
}

// ---------------------------------------------------------------------------
// CarbonCycle — crates/rscm-components/src/components/carbon_cycle.rs
//   P: tau, conc_pi, alpha_temperature, step_size
//   D: 1/tau, step_size/6
//   in : emissions (get), temperature (get), concentration, cumulative_emissions,
//        cumulative_uptake (at_start)      out: same three states
//
// The inputs are frozen over a model step (:144-145), so with x = C - C_pi, k = exp(-alpha*T)/tau and
// e = E/GTC_PER_PPM the right-hand side (:134-158) is LINEAR with constant coefficients:
//     x' = e - k x,      U' = GTC k x,      E_cum' = E.
// One classical RK4 step of size h on x' = e - k x is the affine map
//     x <- P x + c,   q = 1 - z/2 + z^2/6 - z^3/24,  P = 1 - z q,  c = h q e,  z = h k
// (the degree-4 Taylor polynomial of exp(-z) and its integral), so the reference's n sub-steps are the n-fold
// composition of that map — formed here by repeated squaring of (P, c) instead of n x 4 right-hand sides.  The uptake
// needs no stages either: each RK4 step adds h/6 (u0 + 2u1 + 2u2 + u3) GTC with u = k x(stage), and the x update of the
// same step is x += h e - h/6 (u0 + 2u1 + 2u2 + u3); summed over the n sub-steps:  dU = GTC (n h e - (x_n - x_0)).
// Same numbers as stepping up to rounding (parity <= 1e-13 against the stepwise oracle), ~30 FP64 operations instead
// of 170 per model year; exact at equilibrium (x = 0, e = 0 stays 0) and for k = 0.
// ---------------------------------------------------------------------------
constexpr int CARBON_CYCLE_NP = 4;
constexpr int CARBON_CYCLE_ND = 2;

template <class R>
__device__ __forceinline__ void carbon_cycle_prepare(const R *P, R *D)
{
    D[0] = R(1) / P[0];
    D[1] = P[3] / R(6); // h/6
}

template <class R>
__device__ __forceinline__ bool carbon_cycle_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &cx, R *, NodeRef nr)
{
    const int nsub = cx.nsub[nr.rk * cx.Tpad + cx.N];
    if (nsub < 0) return false;
    const R gtc = R(2.13); // GTC_PER_PPM, crates/rscm-components/src/constants.rs:37
    const R conc_pi = P[1], alpha = P[2], h = P[3];
    const R emissions = in[0], temperature = in[1];
    R conc = in[2], cum_e = in[3], cum_u = in[4];
    const R inv_life = r_exp<R>(-(alpha * temperature)) * D[0];
    const R e_ppm = emissions * (R(1) / gtc); // reciprocal of the constant: 1 ulp from E/GTC_PER_PPM
    const R z = h * inv_life;
    const R q = R(1) + z * (R(-0.5) + z * (R(1.0 / 6.0) + z * R(-1.0 / 24.0)));
    R pa = R(1) - z * q, pb = h * q * e_ppm; // one RK4 step: x <- pa x + pb
    R A, B; // composition of the n steps
    if (nsub == 10) { // annual steps: f^10 = f^8 o f^2, straight-line (block-uniform branch)
        const R b2 = pa * pb + pb, a2 = pa * pa;
        const R b4 = a2 * b2 + b2, a4 = a2 * a2;
        const R b8 = a4 * b4 + b4, a8 = a4 * a4;
        A = a8 * a2;
        B = a8 * b2 + b8;
    } else {
        A = R(1);
        B = R(0);
        for (int m = nsub; m > 0; m >>= 1) { // block-uniform trip count
            if (m & 1) { B = pa * B + pb; A = pa * A; }
            pb = pa * pb + pb;
            pa = pa * pa;
        }
    }
    // The concentration is integrated as the anomaly x = C - C_pi (exact subtraction for C within a factor 2 of
    // C_pi), so that uptake is exactly 0 at C = C_pi as in the reference.
    const R x0 = conc - conc_pi;
    const R x = A * x0 + B;
    const R span = R(nsub) * h;
    cum_u = cum_u + gtc * (span * e_ppm - (x - x0));
    cum_e = cum_e + span * emissions;
    conc = x + conc_pi;
    out[0] = conc;
    out[1] = cum_e;
    out[2] = cum_u;
    return true;
}

// ---------------------------------------------------------------------------
// CO2ERF — crates/rscm-components/src/components/co2_erf.rs:57-60
//   P: erf_2xco2, conc_pi      D: erf_2xco2 / ln 2, 1/conc_pi
// ---------------------------------------------------------------------------
constexpr int CO2_ERF_NP = 2;
constexpr int CO2_ERF_ND = 2;

template <class R>
__device__ __forceinline__ void co2_erf_prepare(const R *P, R *D)
{
    D[0] = P[0] / R(0.6931471805599453094172321);
    D[1] = R(1) / P[1];
}

template <class R>
__device__ __forceinline__ bool co2_erf_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    out[0] = D[0] * r_log<R>(R(1) + (in[0] - P[1]) * D[1]);
    return true;
}

// ---------------------------------------------------------------------------
// GhgForcing — crates/rscm-magicc/src/forcing/ghg.rs:122-125,164-279
//   P: see include/rscm_b200.h (21 values, method first)
// ---------------------------------------------------------------------------
constexpr int GHG_FORCING_NP = 21;
constexpr int GHG_FORCING_ND = 3;

template <class R> __device__ __forceinline__ R ghg_overlap_f(R ch4, R n2o);
// D: the terms of the pre-industrial concentrations, which every step subtracts: overlap(ch4_pi, n2o_pi) (Ipcctar only:
// two pow and a log), sqrt(ch4_pi), sqrt(n2o_pi)
template <class R> __device__ __forceinline__ void ghg_forcing_prepare(const R *P, R *D)
{
    D[0] = (P[0] == R(0)) ? ghg_overlap_f<R>(P[2], P[3]) : R(0);
    D[1] = r_sqrt<R>(P[2]);
    D[2] = r_sqrt<R>(P[3]);
}

template <class R> __device__ __forceinline__ R ghg_overlap_f(R ch4, R n2o)
{
    const R mn = ch4 * n2o;
    return R(0.47) * r_log_call<R>(R(1) + R(2.01e-5) * r_pow<R>(mn, R(0.75)) +
                              R(5.31e-15) * ch4 * r_pow<R>(mn, R(1.52)));
}

template <class R>
__device__ __forceinline__ bool ghg_forcing_solve(const R *P, const R *D, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    const R co2 = in[0], ch4 = in[1], n2o = in[2];
    R co2_raw, ch4_raw, n2o_raw;
    if (P[0] == R(0)) { // Ipcctar, ghg.rs:164-200
        co2_raw = (P[4] / R(0.6931471805599453094172321)) * r_log_call<R>(co2 / P[1]);
        ch4_raw = P[5] * (r_sqrt<R>(ch4) - D[1]) - (ghg_overlap_f<R>(ch4, P[3]) - D[0]);
        n2o_raw = P[6] * (r_sqrt<R>(n2o) - D[2]) - (ghg_overlap_f<R>(P[2], n2o) - D[0]);
    } else { // Olbl, ghg.rs:205-259
        const R co2_pi = P[1], a1 = P[7], b1 = P[8], c1 = P[9], d1 = P[10];
        const R delta = co2 - co2_pi;
        const R n2o_overlap = c1 * r_sqrt<R>(n2o);
        const R c_max = co2_pi - b1 / (R(2) * a1);
        R alpha;
        if (co2 >= c_max) alpha = -b1 * b1 / (R(4) * a1) + d1 + n2o_overlap;
        else if (co2 <= co2_pi) alpha = d1 + n2o_overlap;
        else alpha = a1 * delta * delta + b1 * delta + d1 + n2o_overlap;
        co2_raw = alpha * r_log_call<R>(co2 / co2_pi);
        const R s_ch4 = r_sqrt<R>(ch4), s_n2o = r_sqrt<R>(n2o), s_co2 = r_sqrt<R>(co2);
        ch4_raw = (P[11] * s_ch4 + P[12] * s_n2o + P[13]) * (s_ch4 - D[1]);
        n2o_raw = (P[14] * s_co2 + P[15] * s_n2o + P[16] * s_ch4 + P[17]) * (s_n2o - D[2]);
    }
    out[0] = co2_raw * P[18];
    out[1] = ch4_raw * P[19];
    out[2] = n2o_raw * P[20];
    return true;
}

// ---------------------------------------------------------------------------
// OzoneForcing — crates/rscm-magicc/src/forcing/ozone.rs:187-262
//   P: eesc_reference, strat_o3_scale, strat_cl_exponent, trop_radeff, trop_oz_ch4, trop_oz_nox,
//      trop_oz_co, trop_oz_voc, ch4_pi, nox_pi, co_pi, nmvoc_pi, temp_feedback_scale
//   in: EESC, CH4, NOx, CO, NMVOC, temperature    out: strat, trop, temperature-feedback ERF
// ---------------------------------------------------------------------------
template <class R> __device__ __forceinline__ void ozone_forcing_prepare(const R *, R *D) { D[0] = R(0); }

template <class R>
__device__ __forceinline__ bool ozone_forcing_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    const R delta_eesc = in[0] - P[0];
    out[0] = (delta_eesc <= R(0)) ? R(0) : P[1] * r_pow<R>(delta_eesc / R(100), P[2]);
    const R ch4_term = (in[1] > R(0) && P[8] > R(0)) ? P[4] * r_log_call<R>(in[1] / P[8]) : R(0);
    const R precursor = P[5] * (in[2] - P[9]) + P[6] * (in[3] - P[10]) + P[7] * (in[4] - P[11]);
    out[1] = P[3] * (ch4_term + precursor);
    out[2] = P[12] * in[5];
    return true;
}

// ---------------------------------------------------------------------------
// AerosolDirect — crates/rscm-magicc/src/forcing/aerosol_direct.rs:150-192 (FourBox output)
//   P: 4 species coefficients, 4 x regional[4] (SOx, BC, OC, nitrate), sox_pi, bc_pi, oc_pi, nox_pi,
//      harmonize, harmonize_year, harmonize_target (carried; unused by the reference's solve)
// ---------------------------------------------------------------------------
template <class R> __device__ __forceinline__ void aerosol_direct_prepare(const R *, R *D) { D[0] = R(0); }

template <class R> __device__ __forceinline__ R r_abs(R x) { return x < R(0) ? -x : x; }

template <class R>
__device__ __forceinline__ bool aerosol_direct_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    const R f_sox = P[0] * (in[0] - P[20]), f_bc = P[1] * (in[1] - P[21]);
    const R f_oc = P[2] * (in[2] - P[22]), f_nit = P[3] * (in[3] - P[23]);
    const R total = f_sox + f_bc + f_oc + f_nit;
    const R total_abs = r_abs(f_sox) + r_abs(f_bc) + r_abs(f_oc) + r_abs(f_nit);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const R pattern = (r_abs(f_sox) * P[4 + i] + r_abs(f_bc) * P[8 + i] + r_abs(f_oc) * P[12 + i] +
                           r_abs(f_nit) * P[16 + i]) / total_abs;
        R v = total * pattern;
        if (total_abs < R(1e-15)) v = total / R(4);
        if (r_abs(total) < R(1e-15)) v = R(0);
        out[i] = v;
    }
    return true;
}

// ---------------------------------------------------------------------------
// AerosolIndirect — crates/rscm-magicc/src/forcing/aerosol_indirect.rs:152-188
//   P: cloud_albedo_coefficient, reference_burden, sox_weight, oc_weight, sox_pi, oc_pi, harmonize*
// ---------------------------------------------------------------------------
template <class R> __device__ __forceinline__ void aerosol_indirect_prepare(const R *, R *D) { D[0] = R(0); }

template <class R>
__device__ __forceinline__ bool aerosol_indirect_solve(const R *P, const R *, const R *in, R *out, const StepCtx<R> &, R *, NodeRef)
{
    const R burden = P[2] * in[0] + P[3] * in[1];
    const R burden_pi = P[2] * P[4] + P[3] * P[5];
    const R delta = burden - burden_pi;
    out[0] = (delta <= R(0)) ? R(0) : P[0] * r_log_call<R>(R(1) + delta / P[1]);
    return true;
}

// ---------------------------------------------------------------------------
// Schema aggregates — crates/rscm-core/src/schema.rs:760-806 (NaN contributors
// are dropped; all-NaN -> NaN) and read-side grid aggregation
// (state/aggregating.rs:162-177: weighted sum over non-NaN regions).
// ---------------------------------------------------------------------------
template <class R, int N>
__device__ __forceinline__ R agg_sum(const R (&v)[N])
{
    R s = R(0);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (v[i] == v[i]) { s += v[i]; ++cnt; }
    return cnt ? s : r_nan<R>();
}

template <class R, int N>
__device__ __forceinline__ R agg_mean(const R (&v)[N])
{
    R s = R(0);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (v[i] == v[i]) { s += v[i]; ++cnt; }
    return cnt ? s / R(cnt) : r_nan<R>();
}

template <class R, int N>
__device__ __forceinline__ R agg_weighted(const R (&v)[N], const R (&w)[N])
{
    R s = R(0);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (v[i] == v[i]) { s += v[i] * w[i]; ++cnt; }
    return cnt ? s : r_nan<R>();
}

// read transform with the grid's default weights: aggregate_global = plain sum of v*w (NaN propagates)
template <class R, int N>
__device__ __forceinline__ R read_plain(const R (&v)[N], const R (&w)[N])
{
    R s = R(0);
#pragma unroll
    for (int i = 0; i < N; ++i) s += v[i] * w[i];
    return s;
}

// read transform FourBox/Hemispheric -> Scalar with custom weights (ModelBuilder::with_grid_weights):
// NaN regions skipped, no renormalisation
template <class R, int N>
__device__ __forceinline__ R read_weighted(const R (&v)[N], const R (&w)[N])
{
    R s = R(0);
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (v[i] == v[i]) s += v[i] * w[i];
    return s;
}

} // namespace rscm_dev
