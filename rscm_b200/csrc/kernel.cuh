// kernel.cuh — the fused per-member time-loop kernel (sm_100a).
//
// One thread = one run (member m of scenario s = blockIdx.y).  The compiled
// component graph arrives as a `Prog` struct emitted by the host graph compiler
// (graph.cpp: emit_program): straight-line calls into components.cuh with literal
// cell / parameter indices, so every parameter, derived constant and state cell
// is a register.  There are no per-step launches: the whole `Model::run` loop
// (crates/rscm-core/src/model/runtime.rs:515-527) is the `for N` loop below.
//
// Per block, the scenario's exogenous series, the RK4 sub-step tables and the
// observation tables are staged once into shared memory with TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx); parameter loads
// are issued while the copies are in flight.
//
// Outputs are written structure-of-arrays, out[row][run] with run = s*M + m, so a
// warp's store of one (variable, time) row is one contiguous 256 B segment;
// stores carry the streaming (.cs) hint because outputs are never re-read.
#pragma once
// (no standard headers: this file is also compiled by NVRTC at run time)
#include "components.cuh"
#include "climate_udeb.cuh"
#include "magicc_boxes.cuh"
#define RSCM_INF (__longlong_as_double(0x7ff0000000000000LL))

namespace rscm_dev {

constexpr int MAX_SLOTS = 224;  // component parameter slots per program
constexpr int MAX_CELLS = 160;  // scalar storage cells (variables x regions); HalocarbonChemistry alone brings 86
constexpr int MAX_OBS_ROWS = 4; // dense observation tables (one per observed variable)
constexpr int MAX_PEERS = 8;    // GPUs of one NVSwitch domain that share a log-posterior buffer

struct PriorDev {
    int kind;
    int pad;
    double a, b, low, high;
};

struct BlockPartial {
    double max_lp;
    long long argmax;
    double sum_finite;
    long long n_finite;
};

struct SummaryDev {
    double max_logpost;
    long long argmax;
    double sum_finite;
    long long n_finite;
    long long n_runs;
};

struct KArgs {
    // parameter matrix: element (col j, member m) = params[j*ld_col + m*ld_mem]
    const double *params;
    long long M;
    long long ld_col, ld_mem;
    int n_cols;
    int T, Tpad;
    // staged tables (global, 16-byte aligned, sizes multiples of 16 B)
    const double *exo;   // [S][n_exo_rows][Tpad]
    const int *nsub;     // [n_rk][Tpad]   RK4 sub-steps per step; <0: get_last_step would assert
    const double *obs;   // [2][n_obs_rows][Tpad]: values then sigmas (sigma<=0: no observation)
    const double *bounds; // [Tpad + 4] time bounds (T + 1 used), staged only for programs that need the time axis
    const double *ctab;   // [n_ctab] per-graph constant tables of stateful components (host-computed)
    const double *gtab;   // large per-graph tables that stay in global memory (e.g. the ocean IRF by lag)
    // global scratch of stateful components, or null: [S][ceil(M/32)][scratch_rows][32] — the rows of 32 consecutive
    // members of one scenario are one contiguous block, so a warp's access to a row is one 256 B segment and a range of
    // rows of a CTA's 32 members (lane-group programs) is one contiguous piece that a bulk copy can stage
    double *scratch;
    int scratch_rows;
    int n_ctab;
    int n_exo_rows, n_rk, n_obs_rows;
    int normalize;
    int obs_cell[MAX_OBS_ROWS];
    // outputs
    double *out;
    long long runs;
    unsigned char *status;
    double *logpost;
    const PriorDev *priors; // [n_cols] or null
    BlockPartial *partials; // [gridDim.x*gridDim.y] or null
    SummaryDev *summary;
    unsigned int *ticket;
    int t_start, t_stop, t_step; // selected time indices: t_start + k*t_step < t_stop
    // log-posterior placement: run (s, m) goes to logpost[s * lp_ld + m]  (lp_ld = M unless this launch evaluates a member
    // block of a larger ensemble in place: rscm_b200_logpost_sharded_device)
    long long lp_ld;
    // fused all-gather over peer memory: the same element is also stored into every peer's copy of the buffer
    // (peer_lp[p] already offset like `logpost`), and the last CTA raises this rank's flag in every peer's flag array
    int n_peers;
    double *peer_lp[MAX_PEERS];
    unsigned long long *peer_flag[MAX_PEERS]; // &flags_of_peer_p[this rank]
    const unsigned long long *epoch;          // device counter: the flag value to raise is *epoch + 1
    // slot tables (compile-time indexed after unrolling -> constant-bank loads)
    int slot_col[MAX_SLOTS];
    double slot_def[MAX_SLOTS];
    int init_col[MAX_CELLS];
    double init_def[MAX_CELLS];
    long long out_off[MAX_CELLS]; // byte offset of the cell's first output row (row * runs * 8) or -1
};

static_assert(sizeof(KArgs) <= 32764, "kernel parameter block must stay within the launch limit (32764 B since CUDA 12.1, sm_70+)");

__device__ __forceinline__ void store_stream(double *p, double v) { __stcs(p, v); }

// Distribution::ln_pdf — crates/rscm-calibrate/src/distribution.rs:157-163,256-259,353-360,490-497
__device__ inline double prior_ln_pdf(const PriorDev &p, double x)
{
    const double ln_2pi_half = 0.5 * 1.8378770664093454835606594728112;
    int kind = p.kind;
    if (kind >= 4) {
        if (x < p.low || x > p.high) return -RSCM_INF;
        kind = (kind == 4) ? 2 : (kind == 5 ? 3 : 1);
    }
    if (kind == 1) {
        if (x < p.a || x > p.b) return -RSCM_INF;
        return -log(p.b - p.a);
    }
    if (kind == 2) {
        const double z = (x - p.a) / p.b;
        return -0.5 * z * z - log(p.b) - ln_2pi_half;
    }
    if (kind == 3) {
        if (x <= 0.0) return -RSCM_INF;
        const double ln_x = log(x);
        const double z = (ln_x - p.a) / p.b;
        return -0.5 * z * z - ln_x - log(p.b) - ln_2pi_half;
    }
    return 0.0;
}

// GaussianLikelihood::observation_ln_likelihood — likelihood.rs:186-199
template <class R, int NC>
__device__ __forceinline__ void obs_accumulate(const KArgs &a, const double *s_obs, int tidx,
                                               const R *vals, double (&ll)[MAX_OBS_ROWS], bool &bad)
{
#pragma unroll
    for (int j = 0; j < MAX_OBS_ROWS; ++j) {
        if (j < a.n_obs_rows) {
            const double sig = s_obs[(a.n_obs_rows + j) * a.Tpad + tidx];
            if (sig > 0.0) { // block-uniform
                const double val = s_obs[j * a.Tpad + tidx];
                double model = 0.0;
                const int cell = a.obs_cell[j];
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c == cell) model = static_cast<double>(vals[c]);
                // missing (NaN) or non-finite model value => Err => -inf (likelihood.rs:209-221)
                if (!(fabs(model) <= 1.7976931348623157e308)) bad = true;
                const double residual = val - model;
                double l = -0.5 * ((residual * residual) / (sig * sig));
                if (a.normalize) {
                    l -= 0.5 * 1.8378770664093454835606594728112;
                    l -= log(sig);
                }
                ll[j] += l;
            }
        }
    }
}

// Binding specialisation (engine.cu: binding_spec): a run-time compiled program may state which parameter slots are fed by
// a column; the others are literals.  Ahead-of-time programs carry no such statement and keep the run-time slot table.
template <class Prog, class = void> struct ProgSpec {
    static constexpr bool on = false;
    __host__ __device__ static constexpr bool bound(int) { return true; }
    __host__ __device__ static constexpr double value(int) { return 0.0; }
};
template <class Prog> struct ProgSpec<Prog, decltype(void(Prog::SPECIALIZED))> {
    static constexpr bool on = true;
    __host__ __device__ static constexpr bool bound(int i) { return Prog::bound_slot(i); }
    __host__ __device__ static constexpr double value(int i) { return Prog::slot_value(i); }
};

template <class R, class Prog, bool WRITE, bool LOGP>
__global__ void __launch_bounds__(BLOCK, (LOGP && Prog::MIN_BLOCKS > 4) ? Prog::MIN_BLOCKS - 1
                                         : (Prog::LANES > 1 && sizeof(R) == 4 && Prog::MIN_BLOCKS == 4) ? 5 // fp32 rows are half the registers
                                                                                                        : Prog::MIN_BLOCKS) ensemble_kernel(const __grid_constant__ KArgs a)
{
    constexpr int NC = Prog::NC;
    constexpr int NP = Prog::NP;
    constexpr int ND = Prog::ND;

    // shared memory map (all sections 16-byte multiples):
    //   mbarrier | exogenous rows | observation tables | time bounds | graph constant tables | RK4 sub-step tables |
    //   per-thread scratch of stateful components ([Prog::NSM][BLOCK], conflict-free)
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    double *s_exo = reinterpret_cast<double *>(smem + 16);
    // programs with per-thread shared-memory scratch (so that the scratch of two CTAs fits one SM) or with very many
    // exogenous rows leave them in global memory: block-uniform broadcast loads that hit L1/L2
    constexpr bool STAGE_EXO = Prog::STAGE_EXO;
    double *s_obs = s_exo + (STAGE_EXO ? static_cast<unsigned long long>(a.n_exo_rows) * a.Tpad : 0ull);
    double *s_bounds = s_obs + static_cast<unsigned long long>(LOGP ? 2 * a.n_obs_rows : 0) * a.Tpad;
    double *s_ctab = s_bounds + (Prog::NEEDS_TIME ? a.Tpad + 4 : 0);
    int *s_nsub = reinterpret_cast<int *>(s_ctab + a.n_ctab);
    R *s_thread0 = reinterpret_cast<R *>(s_nsub + static_cast<unsigned long long>(a.n_rk) * a.Tpad);
    R *s_thread = s_thread0 + threadIdx.x;
    // exchange area of lane-group programs: [Prog::NXCH][32] doubles after the per-thread scratch ([NSM][BLOCK] 8-byte words)
    double *s_xch = reinterpret_cast<double *>(s_thread0) + static_cast<unsigned long long>(Prog::NSM) * BLOCK;

    const unsigned exo_bytes = STAGE_EXO ? static_cast<unsigned>(a.n_exo_rows) * a.Tpad * 8u : 0u;
    const double *x_exo = STAGE_EXO ? s_exo : a.exo + static_cast<unsigned long long>(blockIdx.y) * a.n_exo_rows * a.Tpad;
    const unsigned obs_bytes = LOGP ? static_cast<unsigned>(2 * a.n_obs_rows) * a.Tpad * 8u : 0u;
    const unsigned bounds_bytes = Prog::NEEDS_TIME ? static_cast<unsigned>(a.Tpad + 4) * 8u : 0u;
    const unsigned ctab_bytes = static_cast<unsigned>(a.n_ctab) * 8u;
    const unsigned nsub_bytes = static_cast<unsigned>(a.n_rk) * a.Tpad * 4u;
    const unsigned total_bytes = exo_bytes + obs_bytes + bounds_bytes + ctab_bytes + nsub_bytes;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0 && total_bytes) {
        mbar_expect_tx(bar, total_bytes);
        if (exo_bytes)
            tma_bulk_g2s(s_exo, a.exo + static_cast<unsigned long long>(blockIdx.y) * a.n_exo_rows * a.Tpad, exo_bytes, bar);
        if (obs_bytes) tma_bulk_g2s(s_obs, a.obs, obs_bytes, bar);
        if (bounds_bytes) tma_bulk_g2s(s_bounds, a.bounds, bounds_bytes, bar);
        if (ctab_bytes) tma_bulk_g2s(s_ctab, a.ctab, ctab_bytes, bar);
        if (nsub_bytes) tma_bulk_g2s(s_nsub, a.nsub, nsub_bytes, bar);
    }

    // ---- per-member parameters (coalesced SoA loads overlap the bulk copies) ----
    // Prog::LANES threads work on one member (programs with ClimateUDEB: 4, see StepCtx in components.cuh): warp w of the
    // CTA is role w of the CTA's 32 members, warp 0 also runs the rest of the component graph thread-per-member
    constexpr int LANES = Prog::LANES;
    static_assert(LANES == 1 || BLOCK / LANES == 32, "a lane group is one lane of each warp of the CTA");
    // Warp w goes to scheduler w % 4 of its SM: were role 0 always warp 0, one scheduler would carry the graph work of
    // every resident CTA.  The roles are rotated by CTA (148 SMs: the CTAs that share an SM differ by multiples of 148).
    const int rot = LANES == 1 ? 0 : static_cast<int>((blockIdx.x + blockIdx.x / 148u) & 3u);
    const int role = LANES == 1 ? 0 : ((static_cast<int>(threadIdx.x) >> 5) + rot) & 3;
    const long long m_raw = LANES == 1 ? static_cast<long long>(blockIdx.x) * BLOCK + threadIdx.x
                                       : static_cast<long long>(blockIdx.x) * 32 + (threadIdx.x & 31);
    const bool active = m_raw < a.M;
    const bool writer = active && role == 0; // the thread that stores the member's outputs
    const long long m = active ? m_raw : a.M - 1;
    const long long run = static_cast<long long>(blockIdx.y) * a.M + m;
    const double *pm = a.params + m * a.ld_mem;

    R P[NP > 0 ? NP : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        using Spec = ProgSpec<Prog>;
        if (Spec::on) P[i] = Spec::bound(i) ? static_cast<R>(__ldg(pm + a.slot_col[i] * a.ld_col)) : static_cast<R>(Spec::value(i));
        else P[i] = a.slot_col[i] >= 0 ? static_cast<R>(__ldg(pm + a.slot_col[i] * a.ld_col)) : static_cast<R>(a.slot_def[i]);
    }
    R D[ND > 0 ? ND : 1];
    Prog::template prepare<R>(P, D);

    double lp = 0.0;
    if (LOGP && a.priors) {
        // ParameterSet::log_prior — parameter_set.rs:255-270
        for (int j = 0; j < a.n_cols; ++j) lp += prior_ln_pdf(a.priors[j], __ldg(pm + j * a.ld_col));
    }

    // The cells of the two time levels.  Small programs keep them in registers and copy next -> current after every step
    // (register moves the compiler folds away); the large MAGICC programs spill them to local memory, where that copy
    // would be two memory operations per cell and step: Prog::SWAP_CELLS exchanges the two pointers instead.
    R cells0[NC], cells1[NC];
    R *cur = cells0, *nxt = cells1;
    // lane-group programs: only role 0 runs the component graph and owns cells
    const bool cells_on = LANES == 1 || role == 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        cur[c] = a.init_col[c] >= 0 ? static_cast<R>(__ldg(pm + a.init_col[c] * a.ld_col)) : static_cast<R>(a.init_def[c]);
        nxt[c] = r_nan<R>();
    }

    if (total_bytes) mbar_wait(bar, 0);
    // Padding threads of the last CTA (clamped to run M-1) are only needed for the block-wide reductions of the
    // log-posterior variants; they must not step (they would race with the real run on its global scratch rows).
    // (lane-group programs synchronise their warps inside the step: their padding members step too, without side effects)
    if (LANES == 1 && !LOGP && !active) return;

#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (Prog::exo_row(c) >= 0) cur[c] = static_cast<R>(x_exo[Prog::exo_row(c) * a.Tpad]);

    // stateful components: small per-thread state in registers/local (S), large in the per-thread
    // shared-memory scratch and the member-interleaved global scratch (cx.scratch[j*runs + run])
    StepCtx<R> cx;
    cx.nsub = s_nsub;
    cx.bounds = s_bounds;
    cx.ctab = s_ctab;
    cx.gtab = a.gtab;
    cx.sm = s_thread;
    cx.scratch = a.scratch ? a.scratch + ((static_cast<long long>(blockIdx.y) * ((a.M + 31) >> 5) + (m >> 5)) * a.scratch_rows) * SCR_LD + (m & 31) : nullptr;
    cx.col = static_cast<int>(m & 31);
    cx.Tpad = a.Tpad;
    cx.N = 0;
    cx.n_steps = a.T - 1;
    cx.role = role;
    cx.lanes = LANES;
    cx.live = active;
    cx.rot = rot;
    cx.xch = s_xch + (threadIdx.x & 31);
    R S[Prog::NS > 0 ? Prog::NS : 1];
    if (LANES > 1 || !LOGP || active) Prog::template init_state<R>(P, D, S, cx);

    double ll[MAX_OBS_ROWS] = {0.0, 0.0, 0.0, 0.0};
    bool bad = false;
    unsigned fail = 0;

    // index 0: initial values / exogenous echo (model/builder.rs:771-781)
    // Output addressing: rows of a variable with R regions advance by R*runs per selected time,
    // so one running byte pointer per region class (1, 2, 4) plus a host-precomputed per-cell
    // byte offset (constant bank) gives each store two integer adds.
    int tnext = a.t_start;
    const long long row_bytes = a.runs * 8;
    char *p1 = reinterpret_cast<char *>(a.out) + run * 8, *p2 = p1, *p4 = p1;
    if (WRITE) {
        if (tnext == 0 && tnext < a.t_stop) {
            if (writer) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (a.out_off[c] >= 0)
                        store_stream(reinterpret_cast<double *>((Prog::regions(c) == 1 ? p1 : Prog::regions(c) == 2 ? p2 : p4) + a.out_off[c]),
                                     static_cast<double>(cur[c]));
            }
            p1 += row_bytes; p2 += 2 * row_bytes; p4 += 4 * row_bytes;
            tnext += a.t_step;
        }
    }
    if (LOGP && cells_on) obs_accumulate<R, NC>(a, s_obs, 0, cur, ll, bad);

    const int T = a.T;
#ifdef RSCM_NODE_CLOCKS // profiling aid (tools/node_clocks.py): where thread 0 of CTA 0 spends a model year outside the nodes
    long long loop_clk[3] = {0, 0, 0}, loop_t = clock64();
#define LOOP_CLK(i) do { const long long now_ = clock64(); loop_clk[i] += now_ - loop_t; loop_t = now_; } while (0)
#else
#define LOOP_CLK(i) do { } while (0)
#endif
    for (int N = 0; N < T - 1; ++N) {
        // Programs whose step is far larger than the instruction cache (the 124-variable MAGICC chain: 128 KB of SASS per
        // model year) keep the warps of a CTA in step, so that they fetch the same instructions together instead of
        // streaming four copies.  (Padding threads have exited or skip the step; a barrier counts non-exited threads.)
        if (Prog::SYNC_STEPS && LANES == 1) __syncthreads();
        if (cells_on) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (Prog::exo_row(c) >= 0) nxt[c] = static_cast<R>(x_exo[Prog::exo_row(c) * a.Tpad + N + 1]);
                else if (Prog::nan_init(c)) nxt[c] = r_nan<R>(); // other cells are overwritten on every path of the step
            }
        }
        cx.N = N;
        LOOP_CLK(0);
        if (LANES > 1 || !LOGP || active) Prog::template step<R>(P, D, cur, nxt, S, cx, fail);
        LOOP_CLK(1);

        if (WRITE) {
            if (N + 1 == tnext && tnext < a.t_stop) { // block-uniform
                if (writer) {
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if (a.out_off[c] >= 0)
                            store_stream(reinterpret_cast<double *>((Prog::regions(c) == 1 ? p1 : Prog::regions(c) == 2 ? p2 : p4) + a.out_off[c]),
                                         static_cast<double>(nxt[c]));
                }
                p1 += row_bytes; p2 += 2 * row_bytes; p4 += 4 * row_bytes;
                tnext += a.t_step;
            }
        }
        if (LOGP && cells_on) obs_accumulate<R, NC>(a, s_obs, N + 1, nxt, ll, bad);
        if (Prog::SWAP_CELLS) {
            R *t = cur; cur = nxt; nxt = t;
        } else if (cells_on) {
#pragma unroll
            for (int c = 0; c < NC; ++c) cur[c] = nxt[c];
        }
        LOOP_CLK(2);
    }
#ifdef RSCM_NODE_CLOCKS
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("loop_clocks exogenous+nan %lld step %lld outputs+likelihood+cells %lld (cycles per year)\n", loop_clk[0] / (T - 1), loop_clk[1] / (T - 1),
               loop_clk[2] / (T - 1));
#endif

    if (a.status && writer) {
        bool nonfinite = false;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (Prog::endogenous(c) && !(fabs(static_cast<double>(cur[c])) <= 1.7976931348623157e308)) nonfinite = true;
        a.status[run] = static_cast<unsigned char>((fail ? 1u : 0u) | (nonfinite ? 2u : 0u));
    }

    if (LOGP) {
        // EnsembleSampler::log_posterior_batch — sampler/ensemble.rs:143-178
        double total = 0.0;
#pragma unroll
        for (int j = 0; j < MAX_OBS_ROWS; ++j)
            if (j < a.n_obs_rows) total += ll[j];
        double post = lp + total;
        if (bad || !(fabs(lp) <= 1.7976931348623157e308)) post = -RSCM_INF;
        const long long lp_idx = static_cast<long long>(blockIdx.y) * a.lp_ld + m;
        if (writer) {
            a.logpost[lp_idx] = post;
            // fused all-gather: 8 B per run straight into every peer's buffer over NVLink (a warp's stores of one peer
            // are one contiguous 256 B segment)
            for (int p = 0; p < a.n_peers; ++p)
                if (a.peer_lp[p] != a.logpost) a.peer_lp[p][lp_idx] = post;
        }

        __shared__ BlockPartial s_part[BLOCK / 32];
        __shared__ bool s_last;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (a.summary) {
            // warp-shuffle + block reduction of the ensemble summary
            const bool fin = writer && (fabs(post) <= 1.7976931348623157e308);
            double vmax = fin ? post : -RSCM_INF;
            long long amax = fin ? run : -1;
            double vsum = fin ? post : 0.0;
            long long cnt = fin ? 1 : 0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double omax = __shfl_down_sync(0xffffffffu, vmax, off);
                const long long oarg = __shfl_down_sync(0xffffffffu, amax, off);
                if (omax > vmax || (omax == vmax && oarg >= 0 && (amax < 0 || oarg < amax))) { vmax = omax; amax = oarg; }
                vsum += __shfl_down_sync(0xffffffffu, vsum, off);
                cnt += __shfl_down_sync(0xffffffffu, cnt, off);
            }
            if (lane == 0) s_part[warp] = BlockPartial{vmax, amax, vsum, cnt};
        }
        if (a.summary || a.n_peers > 0) {
            // every thread's (remote) stores are ordered before this block's ticket
            if (a.n_peers > 0) __threadfence_system();
            __syncthreads();
            const unsigned nblocks = gridDim.x * gridDim.y;
            const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
            if (threadIdx.x == 0) {
                if (a.summary) {
                    BlockPartial r = s_part[0];
                    for (int w = 1; w < BLOCK / 32; ++w) {
                        const BlockPartial o = s_part[w];
                        if (o.max_lp > r.max_lp || (o.max_lp == r.max_lp && o.argmax >= 0 && (r.argmax < 0 || o.argmax < r.argmax))) {
                            r.max_lp = o.max_lp; r.argmax = o.argmax;
                        }
                        r.sum_finite += o.sum_finite;
                        r.n_finite += o.n_finite;
                    }
                    a.partials[bid] = r;
                }
                __threadfence();
                const unsigned t = atomicAdd(a.ticket, 1u);
                s_last = (t == nblocks - 1);
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                if (a.summary) {
                    // last block: deterministic tree over the per-block partials
                    double bmax = -RSCM_INF, bsum = 0.0;
                    long long barg = -1, bcnt = 0;
                    for (unsigned i = threadIdx.x; i < nblocks; i += BLOCK) {
                        const BlockPartial o = a.partials[i];
                        if (o.max_lp > bmax || (o.max_lp == bmax && o.argmax >= 0 && (barg < 0 || o.argmax < barg))) { bmax = o.max_lp; barg = o.argmax; }
                        bsum += o.sum_finite;
                        bcnt += o.n_finite;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double omax = __shfl_down_sync(0xffffffffu, bmax, off);
                        const long long oarg = __shfl_down_sync(0xffffffffu, barg, off);
                        if (omax > bmax || (omax == bmax && oarg >= 0 && (barg < 0 || oarg < barg))) { bmax = omax; barg = oarg; }
                        bsum += __shfl_down_sync(0xffffffffu, bsum, off);
                        bcnt += __shfl_down_sync(0xffffffffu, bcnt, off);
                    }
                    __syncthreads(); // s_part is reused
                    if (lane == 0) s_part[warp] = BlockPartial{bmax, barg, bsum, bcnt};
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        BlockPartial r = s_part[0];
                        for (int w = 1; w < BLOCK / 32; ++w) {
                            const BlockPartial o = s_part[w];
                            if (o.max_lp > r.max_lp || (o.max_lp == r.max_lp && o.argmax >= 0 && (r.argmax < 0 || o.argmax < r.argmax))) {
                                r.max_lp = o.max_lp; r.argmax = o.argmax;
                            }
                            r.sum_finite += o.sum_finite;
                            r.n_finite += o.n_finite;
                        }
                        a.summary->max_logpost = r.max_lp;
                        a.summary->argmax = r.argmax;
                        a.summary->sum_finite = r.sum_finite;
                        a.summary->n_finite = r.n_finite;
                        a.summary->n_runs = a.runs;
                    }
                }
                if (a.n_peers > 0 && threadIdx.x < a.n_peers) {
                    // every block of this rank has stored and fenced (its ticket came after a system-scope fence): tell
                    // each peer that this rank's block of the buffer is complete for this epoch
                    const unsigned long long e = *a.epoch + 1ull;
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flag[threadIdx.x]), "l"(e) : "memory");
                }
                if (threadIdx.x == 0) *a.ticket = 0u; // re-arm for the next launch
            }
        }
    }
}

// Completion of the fused all-gather: wait until every peer has raised its flag for epoch *epoch + 1, then advance the
// epoch.  One warp; lane q watches peer q's flag (acquire at system scope).  A peer that never arrives (crashed process)
// must not hang the GPU: after `timeout_ns` the kernel gives up and records the failure in *error.
__global__ void peer_wait_kernel(const unsigned long long *flags, int n_peers, unsigned long long *epoch, unsigned long long timeout_ns,
                                 int *error)
{
    const unsigned long long want = *epoch + 1ull;
    if (threadIdx.x < n_peers) {
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v >= want) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(error, 1); break; }
            __nanosleep(200);
        }
    }
    __syncwarp();
    __threadfence_system();
    if (threadIdx.x == 0) *epoch = want;
}

// scenario packing: user layout [S][var][T][R_v]  ->  staged rows [S][cell][Tpad]
__global__ void pack_scenarios_kernel(const double *__restrict__ src, double *__restrict__ dst, int n_rows,
                                      int T, int Tpad, long long S, const int *__restrict__ row_src_off,
                                      const int *__restrict__ row_src_stride, long long src_scen_stride)
{
    const long long total = S * n_rows * static_cast<long long>(Tpad);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % Tpad);
        const long long rs = i / Tpad;
        const int row = static_cast<int>(rs % n_rows);
        const long long s = rs / n_rows;
        double v = 0.0;
        if (t < T) v = src[s * src_scen_stride + row_src_off[row] + static_cast<long long>(t) * row_src_stride[row]];
        dst[i] = v;
    }
}

// params transpose [M][n_cols] -> [n_cols][M] is not needed: the kernel takes strides.

// Timeseries::interpolate_into on the device (crates/rscm-core/src/timeseries.rs:586-611 -> Interp1d over the axis VALUES;
// strategies interpolate/strategies/{linear_spline,previous,next}.rs, segment search strategies/mod.rs:24-68 with the
// is_close! boundary snap, relative tolerance 1e-8).  One thread per (series, target time); extrapolation allowed, as in
// interpolate_into.  strategy: 0 Linear (the last source time is trimmed before the search; beyond the ends the first /
// last segment is extended), 1 Next, 2 Previous.  src [n_series][K][R], dst [n_series][T][R].
__device__ __forceinline__ bool interp_is_close(double a, double b)
{
    return fabs(a - b) <= 1e-8 * fmax(fabs(a), fabs(b));
}

__global__ void interpolate_kernel(const double *__restrict__ src_t, long long K, const double *__restrict__ src_v, long long n_series, int R,
                                   const double *__restrict__ dst_t, long long T, int strategy, double *__restrict__ dst_v)
{
    const long long total = n_series * T;
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total; g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long s = g / T, i = g % T;
        const double target = dst_t[i];
        const long long nb = strategy == 0 ? K - 1 : K; // searched boundaries
        long long lo = 0, hi = nb;                      // lower bound: first index with src_t[idx] >= target
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (src_t[mid] < target) lo = mid + 1; else hi = mid;
        }
        const long long idx = lo;
        const bool fwd = idx == nb, bwd = !fwd && idx == 0;
        const bool boundary = !fwd && interp_is_close(src_t[idx], target);
        const double *y = src_v + s * K * R;
        double *out = dst_v + (s * T + i) * R;
        for (int r = 0; r < R; ++r) {
            double v;
            if (strategy == 0) {
                const long long e = idx < K - 1 ? idx : K - 1;
                if (boundary) v = y[e * R + r];
                else {
                    long long i1, i2;
                    if (bwd) { i1 = 0; i2 = 1; }
                    else if (fwd) { i1 = K - 2; i2 = K - 1; }
                    else { i1 = e - 1; i2 = e; }
                    const double t1 = src_t[i1], t2 = src_t[i2], y1 = y[i1 * R + r], y2 = y[i2 * R + r];
                    const double m = __ddiv_rn(y2 - y1, t2 - t1);
                    v = __dadd_rn(__dmul_rn(m, target - t1), y1); // m * (target - t1) + y1, not contracted (host arithmetic)
                }
            } else if (strategy == 2) { // Previous
                if (boundary) v = y[idx * R + r];
                else if (bwd) v = y[r];
                else if (fwd) v = y[(K - 1) * R + r];
                else v = y[(idx - 1) * R + r];
            } else { // Next
                const long long e = idx < K - 1 ? idx : K - 1;
                if (bwd) v = y[r];
                else if (fwd) v = y[(K - 1) * R + r];
                else v = y[e * R + r];
            }
            out[r] = v;
        }
    }
}

// self-test hook for the device exp / log (components.cuh): y[i] = f(x[i])
__global__ void device_math_kernel(int op, const double *__restrict__ x, long long n, double *__restrict__ y)
{
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] = op == 0 ? rscm_exp(x[i]) : op == 1 ? rscm_log(x[i]) : rscm_pow(x[2 * i], x[2 * i + 1]);
}

// FMA-pipe saturating micro-benchmark (roofline denominator for the FP64/FP32 bound)
template <class R>
__global__ void __launch_bounds__(256) fma_peak_kernel(R *sink, int iters, R seed)
{
    R a0 = seed + R(threadIdx.x), a1 = a0 + R(1), a2 = a0 + R(2), a3 = a0 + R(3);
    R a4 = a0 + R(4), a5 = a0 + R(5), a6 = a0 + R(6), a7 = a0 + R(7);
    const R b = R(0.999999), c = R(1e-6);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
            a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
        }
    }
    const R s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == R(-1)) sink[0] = s; // never true; keeps the chain live
}

} // namespace rscm_dev
