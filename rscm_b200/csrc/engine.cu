// engine.cu — C-ABI implementation of the B200 ensemble engine (include/rscm_b200.h).
//
// Host side: graph compile (graph.cpp) -> look the emitted program up in the
// ahead-of-time registry (aot_programs.inc, generated at build time by
// tools/gen_aot from the same emitter) -> launch the fused member-loop kernel
// (kernel.cuh).  There is no CPU execution path in this library.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "graph.hpp"
#include "jit.hpp"
#include "kernel.cuh"
#include "sampler.cuh"
#include "summary.cuh"

using rscm_dev::KArgs;

namespace {

typedef cudaError_t (*launch_fn)(int dtype, bool write, bool logp, dim3 grid, size_t smem, cudaStream_t st,
                                 const KArgs &a);

struct AotEntry {
    const char *name;
    const char *signature;
    launch_fn launch;
};

template <class Prog>
cudaError_t launch_prog(int dtype, bool write, bool logp, dim3 grid, size_t smem, cudaStream_t st, const KArgs &a)
{
    using namespace rscm_dev;
    const dim3 block(BLOCK);
#define RSCM_LAUNCH(R, W, L)                                                                                                   \
    do {                                                                                                                       \
        if (smem > 48u * 1024u)                                                                                                \
            cudaFuncSetAttribute(ensemble_kernel<R, Prog, W, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
        ensemble_kernel<R, Prog, W, L><<<grid, block, smem, st>>>(a);                                                          \
    } while (0)
    if (dtype == 0) {
        if (write && !logp) RSCM_LAUNCH(double, true, false);
        else if (!write && logp) RSCM_LAUNCH(double, false, true);
        else RSCM_LAUNCH(double, true, true);
    } else {
        if (write && !logp) RSCM_LAUNCH(float, true, false);
        else if (!write && logp) RSCM_LAUNCH(float, false, true);
        else RSCM_LAUNCH(float, true, true);
    }
#undef RSCM_LAUNCH
    return cudaGetLastError();
}

#include "aot_programs.inc"

thread_local std::string g_global_err;

} // namespace

struct rscm_b200_ensemble {
    rscm::Graph g;
    int dtype = 0;
    int device = 0;
    const AotEntry *prog = nullptr;
    bool use_jit = false; // program compiled at run time (jit.cpp) instead of taken from the AOT registry
    std::string jit_cubin[3], jit_names[3]; // per kernel variant (write / logpost / both), compiled on first use
    std::string jit_spec;                   // binding specialisation the loaded variants were compiled for (binding_spec)
    rscm::JitProgram jit;
    std::string err;
    int Tpad = 0;

    // bindings
    int n_cols = 0;
    std::vector<int> slot_col, init_col;
    // output selection
    std::vector<int> sel_vars;
    int t_start = 0, t_stop = 0, t_step = 1, n_tsel = 0;
    std::vector<int> out_base, out_tmul;
    int64_t rows = 0;
    // target / priors
    int n_obs_rows = 0;
    int obs_cell[rscm_dev::MAX_OBS_ROWS] = {0, 0, 0, 0};
    int normalize = 0;
    bool has_target = false;
    int n_priors = 0;

    // device scratch
    double *d_exo = nullptr;
    int64_t exo_capacity_S = 0;
    int *d_nsub = nullptr;
    double *d_bounds = nullptr, *d_ctab = nullptr, *d_gtab = nullptr;
    double *d_scratch[2] = {nullptr, nullptr};
    int64_t cap_scratch[2] = {0, 0};
    double *d_obs = nullptr;
    rscm_dev::PriorDev *d_priors = nullptr;
    rscm_dev::BlockPartial *d_partials = nullptr;
    int64_t partials_capacity = 0;
    unsigned *d_ticket = nullptr;
    int *d_row_off = nullptr, *d_row_stride = nullptr;
    int64_t scen_stride = 0;
    // host pipeline
    cudaStream_t streams[2] = {nullptr, nullptr};
    double *d_params[2] = {nullptr, nullptr};
    double *d_out[2] = {nullptr, nullptr};
    unsigned char *d_status[2] = {nullptr, nullptr};
    double *d_scen = nullptr;
    double *d_logpost = nullptr;
    rscm_dev::SummaryDev *d_summary = nullptr;
    int64_t cap_params = 0, cap_out = 0, cap_status = 0, cap_scen = 0, cap_logpost = 0;

    // ordering between launches of this handle on different streams: d_exo, the scratch slots, the summary partials and
    // the ticket are per handle, so a launch on another stream first waits for the previous launch's kernel
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;

    // stats
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t ev_next = 0;
    std::vector<char> ev_pending;
    double kernel_ms_sum = 0.0;
    int64_t kernel_ms_n = 0;
};

// A communicator over the ranks that share one ensemble evaluation (one process per GPU).  NCCL carries the generic
// collectives; buffers obtained from rscm_b200_comm_symmetric_alloc are additionally mapped into every peer (CUDA IPC over
// NVLink / NVSwitch), which is what the fused log-posterior + all-gather kernel stores through.
struct rscm_b200_comm {
    void *nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    bool p2p = false; // every peer's symmetric memory is mapped here
    std::string err;
    struct Segment {
        void *local = nullptr;
        size_t bytes = 0;
        void *peer[rscm_dev::MAX_PEERS] = {};
    };
    std::vector<Segment> segments; // segments[0] = control block: flags[MAX_PEERS], epoch, error
    unsigned long long *flags = nullptr, *d_epoch = nullptr;
    int *d_error = nullptr;
    unsigned long long timeout_ns = 20000000000ull;
    // sampler iteration recorded as a CUDA graph (rscm_b200_sampler_iterate)
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<uintptr_t> graph_key;
    unsigned *d_iteration = nullptr;
};

namespace {

int fail(rscm_b200_ensemble *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    g_global_err = msg;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, RSCM_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

void recompute_selection(rscm_b200_ensemble *h)
{
    const rscm::Graph &g = h->g;
    h->out_base.assign(g.n_cells, -1);
    h->out_tmul.assign(g.n_cells, 1);
    h->n_tsel = 0;
    for (int t = h->t_start; t < h->t_stop; t += h->t_step) ++h->n_tsel;
    int64_t row = 0;
    for (int v : h->sel_vars) {
        const rscm::Variable &var = g.vars[v];
        for (int r = 0; r < var.n_regions; ++r) {
            h->out_base[var.cell0 + r] = static_cast<int>(row + r);
            h->out_tmul[var.cell0 + r] = var.n_regions;
        }
        row += static_cast<int64_t>(h->n_tsel) * var.n_regions;
    }
    h->rows = row;
}

void harvest_event(rscm_b200_ensemble *h, size_t i)
{
    if (!h->ev_pending[i]) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->events[i].first, h->events[i].second) == cudaSuccess) {
        h->kernel_ms_sum += ms;
        h->kernel_ms_n += 1;
    }
    h->ev_pending[i] = 0;
}

size_t smem_bytes(const rscm_b200_ensemble *h, bool logp)
{
    size_t b = 16;
    if (h->g.stage_exo) b += static_cast<size_t>(h->g.n_exo_rows) * h->Tpad * 8; // Prog::STAGE_EXO
    if (logp) b += static_cast<size_t>(2 * h->n_obs_rows) * h->Tpad * 8;
    if (h->g.needs_time) b += static_cast<size_t>(h->Tpad + 4) * 8;
    b += h->g.ctab.size() * 8;
    b += static_cast<size_t>(h->g.n_rk) * h->Tpad * 4;
    b += static_cast<size_t>(h->g.n_smem) * rscm_dev::BLOCK * 8; // n_smem counts 8-byte words per thread in both dtypes
    b += static_cast<size_t>(h->g.n_xch) * 32 * 8;                // exchange area of lane-group programs
    return b;
}

// Run-time compiled programs are specialised on the parameter binding: a slot that no column feeds is a literal of the
// program instead of a register loaded through the kernel's slot table.  The MAGICC kinds carry dozens of parameters per
// component of which an ensemble varies a handful; without this every one of them is a live per-thread value (config 4:
// 107 doubles, spilled to local memory).  The text is appended to the emitted `Prog` body, so the disk cache keys on it.
std::string binding_spec(const rscm_b200_ensemble *h)
{
    const rscm::Graph &g = h->g;
    for (int i = 0; i < g.n_slots; ++i)
        if (!std::isfinite(g.slot_default[i])) return std::string(); // literals cannot carry NaN / inf: stay generic
    std::string s = "    static constexpr bool SPECIALIZED = true;\n    __host__ __device__ static constexpr bool bound_slot(int i) { return ";
    for (int i = 0; i < g.n_slots; ++i)
        if (h->slot_col[i] >= 0) s += "i == " + std::to_string(i) + " || ";
    s += "false; }\n    __host__ __device__ static constexpr double slot_value(int i) { return ";
    char buf[64];
    for (int i = 0; i < g.n_slots; ++i) {
        if (h->slot_col[i] >= 0 || g.slot_default[i] == 0.0) continue;
        std::snprintf(buf, sizeof buf, "%.17g", g.slot_default[i]);
        std::string lit = buf;
        if (lit.find_first_of(".eEn") == std::string::npos) lit += ".0";
        s += "i == " + std::to_string(i) + " ? " + lit + " : ";
    }
    s += "0.0; }\n";
    return s;
}

// What a launch that evaluates a member block of a larger ensemble in place needs beyond the plain arguments.
struct LaunchOpts {
    int64_t ld_col = 0; // layout 0: leading dimension of the parameter matrix (0 = M)
    int64_t lp_ld = 0;  // scenario stride of the log-posterior array (0 = M)
    int n_peers = 0;    // fused all-gather over peer memory (kernel.cuh)
    double *peer_lp[rscm_dev::MAX_PEERS] = {};
    unsigned long long *peer_flag[rscm_dev::MAX_PEERS] = {};
    const unsigned long long *epoch = nullptr;
};

// enqueue scenario packing + the fused kernel on `st`
int enqueue(rscm_b200_ensemble *h, const double *d_params, int64_t M, int layout, const double *d_scen, int64_t S,
            double *d_out, unsigned char *d_status, double *d_logpost, rscm_dev::SummaryDev *d_summary, bool write,
            bool logp, cudaStream_t st, bool pack, const LaunchOpts *opts = nullptr)
{
    const rscm::Graph &g = h->g;
    if (M <= 0) return fail(h, RSCM_B200_EINVAL, "M must be positive");
    if (g.n_exo_rows > 0 && (S <= 0 || !d_scen)) return fail(h, RSCM_B200_EINVAL, "this graph needs scenarios for its exogenous variables");
    if (g.n_exo_rows == 0 && S <= 0) S = 1;
    if (S > 65535) return fail(h, RSCM_B200_EINVAL, "at most 65535 scenarios per call");
    if (h->n_cols > 0 && !d_params) return fail(h, RSCM_B200_EINVAL, "parameter matrix required (columns are bound)");
    if (logp && !h->has_target) return fail(h, RSCM_B200_EINVAL, "set a target before evaluating the log-posterior");

    // inside a stream capture (rscm_b200_sampler_iterate records its iteration as a CUDA graph) nothing may allocate,
    // synchronise or time: the buffers were sized by the eager iteration that precedes every capture
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(st, &cap));
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    if (!capturing && h->has_last && h->last_stream != st) CU(cudaStreamWaitEvent(st, h->last_done, 0));

    if (g.n_exo_rows > 0) {
        if (S > h->exo_capacity_S) {
            if (capturing) return fail(h, RSCM_B200_EINVAL, "scenario table would have to grow inside a stream capture");
            if (h->has_last) CU(cudaEventSynchronize(h->last_done)); // the old table may still be read
            if (h->d_exo) cudaFree(h->d_exo);
            h->d_exo = nullptr;
            CU(cudaMalloc(&h->d_exo, static_cast<size_t>(S) * g.n_exo_rows * h->Tpad * 8));
            h->exo_capacity_S = S;
        }
        if (pack) {
            const long long total = S * g.n_exo_rows * static_cast<long long>(h->Tpad);
            const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
            rscm_dev::pack_scenarios_kernel<<<blocks, 256, 0, st>>>(d_scen, h->d_exo, g.n_exo_rows, g.T, h->Tpad, S,
                                                                   h->d_row_off, h->d_row_stride, h->scen_stride);
            CU(cudaGetLastError());
            h->launches++;
        }
    }

    KArgs a;
    std::memset(&a, 0, sizeof a);
    a.params = d_params;
    a.M = M;
    if (layout == 0) { a.ld_col = (opts && opts->ld_col) ? opts->ld_col : M; a.ld_mem = 1; } else { a.ld_col = 1; a.ld_mem = h->n_cols; }
    a.lp_ld = (opts && opts->lp_ld) ? opts->lp_ld : M;
    if (opts && opts->n_peers > 0) {
        a.n_peers = opts->n_peers;
        for (int p = 0; p < opts->n_peers; ++p) { a.peer_lp[p] = opts->peer_lp[p]; a.peer_flag[p] = opts->peer_flag[p]; }
        a.epoch = opts->epoch;
        a.ticket = h->d_ticket;
    }
    a.n_cols = h->n_cols;
    a.T = g.T;
    a.Tpad = h->Tpad;
    a.exo = h->d_exo;
    a.nsub = h->d_nsub;
    a.bounds = h->d_bounds;
    a.ctab = h->d_ctab;
    a.gtab = h->d_gtab;
    a.n_ctab = static_cast<int>(g.ctab.size());
    if (g.n_scratch_rows > 0) {
        // global scratch of stateful components, one buffer per launch stream slot
        const int64_t need = static_cast<int64_t>(g.n_scratch_rows) * S * ((M + 31) / 32) * 32; // blocks of 32 members (KArgs::scratch)
        a.scratch_rows = g.n_scratch_rows;
        const int slot = (st == h->streams[1] && st) ? 1 : 0;
        if (need > h->cap_scratch[slot]) {
            if (capturing) return fail(h, RSCM_B200_EINVAL, "scratch would have to grow inside a stream capture");
            if (h->d_scratch[slot]) cudaFree(h->d_scratch[slot]); // (cudaFree synchronises the device)
            h->d_scratch[slot] = nullptr;
            h->cap_scratch[slot] = 0;
            CU(cudaMalloc(&h->d_scratch[slot], static_cast<size_t>(need) * 8));
            h->cap_scratch[slot] = need;
        }
        a.scratch = h->d_scratch[slot];
    }
    a.obs = h->d_obs;
    a.n_exo_rows = g.n_exo_rows;
    a.n_rk = g.n_rk;
    a.n_obs_rows = logp ? h->n_obs_rows : 0;
    a.normalize = h->normalize;
    for (int j = 0; j < rscm_dev::MAX_OBS_ROWS; ++j) a.obs_cell[j] = h->obs_cell[j];
    a.out = d_out;
    a.runs = S * M;
    a.status = d_status;
    a.logpost = d_logpost;
    a.priors = (logp && h->n_priors > 0) ? h->d_priors : nullptr;
    a.t_start = h->t_start;
    a.t_stop = h->t_stop;
    a.t_step = h->t_step;
    for (int i = 0; i < g.n_slots; ++i) { a.slot_col[i] = h->slot_col[i]; a.slot_def[i] = g.slot_default[i]; }
    for (int c = 0; c < g.n_cells; ++c) {
        a.init_col[c] = h->init_col[c];
        const rscm::Variable &var = g.vars[g.cell_var[c]];
        a.init_def[c] = var.has_initial ? var.initial : std::numeric_limits<double>::quiet_NaN();
        a.out_off[c] = (write && h->out_base[c] >= 0) ? static_cast<long long>(h->out_base[c]) * a.runs * 8 : -1;
    }
    const int64_t members_per_cta = rscm_dev::BLOCK / g.lanes; // Prog::LANES threads (one per warp of the CTA) work on one member
    const dim3 grid(static_cast<unsigned>((M + members_per_cta - 1) / members_per_cta), static_cast<unsigned>(S));
    if (logp && d_summary) {
        const int64_t nb = static_cast<int64_t>(grid.x) * grid.y;
        if (nb > h->partials_capacity) {
            if (capturing) return fail(h, RSCM_B200_EINVAL, "summary partials would have to grow inside a stream capture");
            if (h->d_partials) cudaFree(h->d_partials);
            h->d_partials = nullptr;
            CU(cudaMalloc(&h->d_partials, static_cast<size_t>(nb) * sizeof(rscm_dev::BlockPartial)));
            h->partials_capacity = nb;
        }
        a.partials = h->d_partials;
        a.summary = d_summary;
        a.ticket = h->d_ticket;
    }

    if (h->use_jit) { // compile / load before the timing events: compilation is host time, not kernel time
        const int variant = (write && !logp) ? 0 : ((!write && logp) ? 1 : 2);
        const std::string spec = binding_spec(h);
        if (spec != h->jit_spec) { // the binding changed: the loaded variants were specialised for another one
            if (capturing) return fail(h, RSCM_B200_EINVAL, "program would have to be recompiled inside a stream capture");
            cudaDeviceSynchronize();
            rscm::jit_unload(h->jit);
            for (int v = 0; v < 3; ++v) { h->jit_cubin[v].clear(); h->jit_names[v].clear(); }
            h->jit_spec = spec;
        }
        if (!h->jit.fn[variant]) { // first use of this kernel variant: compile (disk-cached) and load it
            if (capturing) return fail(h, RSCM_B200_EINVAL, "program would have to be compiled inside a stream capture");
            std::string jerr;
            if (h->jit_cubin[variant].empty() &&
                !rscm::jit_compile_cubin(g.program_source + spec, h->dtype, variant, h->jit_cubin[variant], h->jit_names[variant], jerr))
                return fail(h, RSCM_B200_EUNSUPPORTED, "run-time compilation failed: " + jerr);
            if (!rscm::jit_load(h->jit_cubin[variant], h->jit_names[variant], variant, h->jit, jerr))
                return fail(h, RSCM_B200_ECUDA, "loading the run-time compiled program failed: " + jerr);
        }
    }

    // CUDA-event timing of the fused kernel on its launch stream
    size_t ei = h->ev_next;
    if (capturing) {
        // no timing and no ordering event inside a capture
    } else if (h->events.size() < 64) {
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        h->events.push_back({e0, e1});
        h->ev_pending.push_back(0);
        ei = h->events.size() - 1;
    } else {
        if (h->ev_pending[ei]) {
            cudaEventSynchronize(h->events[ei].second);
            harvest_event(h, ei);
        }
    }
    if (!capturing) {
        h->ev_next = (ei + 1) % 64;
        CU(cudaEventRecord(h->events[ei].first, st));
    }
    if (h->use_jit) {
        const int variant = (write && !logp) ? 0 : ((!write && logp) ? 1 : 2);
        const int rc = rscm::jit_launch(h->jit, variant, grid.x, grid.y, rscm_dev::BLOCK, static_cast<unsigned>(smem_bytes(h, logp)), st, &a);
        if (rc != 0) return fail(h, RSCM_B200_ECUDA, "cuLaunchKernel failed with CUresult " + std::to_string(rc));
    } else {
        cudaError_t e = h->prog->launch(h->dtype, write, logp, grid, smem_bytes(h, logp), st, a);
        if (e != cudaSuccess) return fail(h, RSCM_B200_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    }
    if (!capturing) {
        CU(cudaEventRecord(h->events[ei].second, st));
        h->ev_pending[ei] = 1;
        h->last_done = h->events[ei].second;
        h->last_stream = st;
        h->has_last = true;
    }
    h->launches++;
    return RSCM_B200_OK;
}

// ---- NCCL, loaded at run time (the library links against nothing but the CUDA runtime) -----------------------------------
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};
constexpr int NCCL_INT8 = 0, NCCL_FLOAT64 = 8;

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    // 1. an explicit path; 2. the copy already mapped into this process (a Python host has torch's); 3. the system's
    const char *env = getenv("RSCM_B200_NCCL_LIB");
    void *lib = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
#define RSCM_SYM(field, name)                                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name));                         \
    if (!api.field) { api.err = std::string("libnccl lacks ") + name; return nullptr; }
    RSCM_SYM(GetUniqueId, "ncclGetUniqueId")
    RSCM_SYM(CommInitRank, "ncclCommInitRank")
    RSCM_SYM(CommDestroy, "ncclCommDestroy")
    RSCM_SYM(AllGather, "ncclAllGather")
    RSCM_SYM(Broadcast, "ncclBroadcast")
    RSCM_SYM(GroupStart, "ncclGroupStart")
    RSCM_SYM(GroupEnd, "ncclGroupEnd")
    RSCM_SYM(GetErrorString, "ncclGetErrorString")
#undef RSCM_SYM
    api.lib = lib;
    return &api;
}

template <class T> int ensure(rscm_b200_ensemble *h, T **p, int64_t *cap, int64_t need)
{
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CU(cudaMalloc(reinterpret_cast<void **>(p), static_cast<size_t>(need) * sizeof(T)));
    *cap = need;
    return 0;
}

} // namespace

extern "C" {

int rscm_b200_abi_version(void) { return RSCM_B200_ABI_VERSION; }

int rscm_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *rscm_b200_last_global_error(void) { return g_global_err.c_str(); }
const char *rscm_b200_last_error(const rscm_b200_ensemble *h) { return h ? h->err.c_str() : g_global_err.c_str(); }

int rscm_b200_ensemble_create(const rscm_b200_graph_desc *desc, rscm_b200_ensemble **out)
{
    if (!desc || !out) return fail(nullptr, RSCM_B200_EINVAL, "null argument");
    *out = nullptr;
    rscm_b200_ensemble *h = new rscm_b200_ensemble();
    std::string err;
    if (!rscm::compile_graph(*desc, h->g, err)) {
        delete h;
        return fail(nullptr, RSCM_B200_EINVAL, err);
    }
    const rscm::Graph &g = h->g;
    if (g.n_cells > rscm_dev::MAX_CELLS || g.n_slots > rscm_dev::MAX_SLOTS) {
        delete h;
        return fail(nullptr, RSCM_B200_EUNSUPPORTED, "graph exceeds the engine's cell/slot limits");
    }
    h->dtype = desc->compute_dtype ? 1 : 0;
    for (const AotEntry &e : g_aot)
        if (g.signature == e.signature) { h->prog = &e; break; }
    if (!h->prog) {
        // not one of the ahead-of-time graphs: compile the emitted program at run time (NVRTC)
        std::string jerr;
        // the plain `write` variant is compiled now (a program that does not compile is refused at creation, also for
        // host-only handles); the log-posterior variants follow on first use
        // Device handles compile on first use, specialised on the parameter binding (binding_spec); a host-only handle
        // (and RSCM_B200_EAGER_JIT=1) compiles the unspecialised program now, so that a graph whose program does not
        // compile is refused at creation.
        if ((desc->device == -2 || std::getenv("RSCM_B200_EAGER_JIT")) &&
            !rscm::jit_compile_cubin(g.program_source, h->dtype, 0, h->jit_cubin[0], h->jit_names[0], jerr)) {
            delete h;
            return fail(nullptr, RSCM_B200_EUNSUPPORTED,
                        "no ahead-of-time device program for this component graph and run-time compilation failed "
                        "(there is no CPU fallback): " + jerr);
        }
        h->jit_cubin[0].clear();
        h->jit_names[0].clear();
        h->use_jit = true;
    }
    h->Tpad = (g.T + 3) & ~3;
    h->slot_col.assign(g.n_slots, -1);
    h->init_col.assign(g.n_cells, -1);
    for (size_t v = 0; v < g.vars.size(); ++v) h->sel_vars.push_back(static_cast<int>(v));
    h->t_start = 0;
    h->t_stop = g.T;
    h->t_step = 1;
    recompute_selection(h);
    if (desc->device == -2) { // host-only handle: graph compile + introspection, never runs
        h->device = -2;
        *out = h;
        return RSCM_B200_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        delete h;
        return fail(nullptr, RSCM_B200_ENODEVICE, "no CUDA device available; the engine has no CPU fallback");
    }
    if (desc->device >= 0) {
        if (desc->device >= ndev || cudaSetDevice(desc->device) != cudaSuccess) {
            delete h;
            return fail(nullptr, RSCM_B200_ENODEVICE, "cannot select the requested CUDA device");
        }
    }
    cudaGetDevice(&h->device);
    if (h->use_jit) cudaFree(nullptr); // make the primary context current for the driver-API module loads

    // RK4 sub-step tables
    if (g.n_rk > 0) {
        std::vector<int> tab(static_cast<size_t>(g.n_rk) * h->Tpad, 0);
        for (int r = 0; r < g.n_rk; ++r)
            for (int t = 0; t < g.T; ++t) tab[static_cast<size_t>(r) * h->Tpad + t] = g.rk_nsub[r][t];
        if (cudaMalloc(&h->d_nsub, tab.size() * sizeof(int)) != cudaSuccess ||
            cudaMemcpy(h->d_nsub, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
            std::string m = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
            rscm_b200_ensemble_destroy(h);
            return fail(nullptr, RSCM_B200_ECUDA, m);
        }
    }
    if (g.needs_time) {
        std::vector<double> b(static_cast<size_t>(h->Tpad) + 4, 0.0);
        for (int t = 0; t <= g.T; ++t) b[t] = g.bounds[t];
        cudaMalloc(&h->d_bounds, b.size() * 8);
        cudaMemcpy(h->d_bounds, b.data(), b.size() * 8, cudaMemcpyHostToDevice);
    }
    if (!g.ctab.empty()) {
        cudaMalloc(&h->d_ctab, g.ctab.size() * 8);
        cudaMemcpy(h->d_ctab, g.ctab.data(), g.ctab.size() * 8, cudaMemcpyHostToDevice);
    }
    if (!g.gtab.empty()) {
        cudaMalloc(&h->d_gtab, g.gtab.size() * 8);
        cudaMemcpy(h->d_gtab, g.gtab.data(), g.gtab.size() * 8, cudaMemcpyHostToDevice);
    }
    // scenario row map: user layout [S][exo var][T][R] -> staged rows
    if (g.n_exo_rows > 0) {
        std::vector<int> off(g.n_exo_rows), stride(g.n_exo_rows);
        int64_t base = 0;
        int row = 0;
        for (int v : g.exo_vars) {
            const rscm::Variable &var = g.vars[v];
            for (int r = 0; r < var.n_regions; ++r) {
                off[row] = static_cast<int>(base + r);
                stride[row] = var.n_regions;
                ++row;
            }
            base += static_cast<int64_t>(g.T) * var.n_regions;
        }
        h->scen_stride = base;
        cudaMalloc(&h->d_row_off, off.size() * sizeof(int));
        cudaMalloc(&h->d_row_stride, off.size() * sizeof(int));
        cudaMemcpy(h->d_row_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(h->d_row_stride, stride.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    cudaMalloc(&h->d_ticket, sizeof(unsigned));
    cudaMemset(h->d_ticket, 0, sizeof(unsigned));
    if (cudaGetLastError() != cudaSuccess) {
        rscm_b200_ensemble_destroy(h);
        return fail(nullptr, RSCM_B200_ECUDA, "device allocation failed");
    }
    *out = h;
    return RSCM_B200_OK;
}

void rscm_b200_ensemble_destroy(rscm_b200_ensemble *h)
{
    if (!h) return;
    cudaFree(h->d_exo); cudaFree(h->d_nsub); cudaFree(h->d_obs); cudaFree(h->d_priors); cudaFree(h->d_partials);
    cudaFree(h->d_bounds); cudaFree(h->d_ctab); cudaFree(h->d_gtab); cudaFree(h->d_scratch[0]); cudaFree(h->d_scratch[1]);
    cudaFree(h->d_ticket); cudaFree(h->d_row_off); cudaFree(h->d_row_stride); cudaFree(h->d_scen);
    cudaFree(h->d_logpost); cudaFree(h->d_summary);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_params[i]); cudaFree(h->d_out[i]); cudaFree(h->d_status[i]);
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    for (auto &e : h->events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (h->use_jit) rscm::jit_unload(h->jit);
    cudaGetLastError();
    delete h;
}

int rscm_b200_n_variables(const rscm_b200_ensemble *h) { return static_cast<int>(h->g.vars.size()); }
const char *rscm_b200_variable_name(const rscm_b200_ensemble *h, int v)
{
    return (v >= 0 && v < static_cast<int>(h->g.vars.size())) ? h->g.vars[v].name.c_str() : nullptr;
}
int rscm_b200_variable_grid(const rscm_b200_ensemble *h, int v) { return h->g.vars[v].grid; }
int rscm_b200_variable_is_endogenous(const rscm_b200_ensemble *h, int v) { return h->g.vars[v].endogenous ? 1 : 0; }
int rscm_b200_variable_index(const rscm_b200_ensemble *h, const char *name) { return h->g.find_var(name); }
int rscm_b200_n_exogenous(const rscm_b200_ensemble *h) { return static_cast<int>(h->g.exo_vars.size()); }
int rscm_b200_exogenous_variable(const rscm_b200_ensemble *h, int i) { return h->g.exo_vars[i]; }
int rscm_b200_n_nodes(const rscm_b200_ensemble *h) { return static_cast<int>(h->g.nodes.size()); }
int rscm_b200_execution_order(const rscm_b200_ensemble *h, int *order, int capacity)
{
    const int n = static_cast<int>(h->g.order.size());
    for (int i = 0; i < n && i < capacity; ++i) order[i] = h->g.order[i];
    return n;
}
int rscm_b200_variable_source(const rscm_b200_ensemble *h, int component, const char *variable)
{
    if (component < 0 || component >= static_cast<int>(h->g.nodes.size())) return -1;
    const rscm::Node &n = h->g.nodes[component];
    const int v = h->g.find_var(variable);
    for (size_t i = 0; i < n.in_var.size(); ++i)
        if (n.in_var[i] == v) return n.in_src[i];
    return -1;
}
const char *rscm_b200_program_signature(const rscm_b200_ensemble *h)
{
    if (!h->use_jit) return h->g.signature.c_str();
    // run-time compiled programs: the text that is compiled, i.e. with the current binding specialisation
    thread_local std::string text;
    text = h->g.program_source + binding_spec(h);
    return text.c_str();
}
int rscm_b200_program_is_jit(const rscm_b200_ensemble *h) { return h->use_jit ? 1 : 0; }
int rscm_b200_time_index(const rscm_b200_ensemble *h, double time) { return rscm::time_index_for(h->g, time); }

int rscm_b200_bind_parameters(rscm_b200_ensemble *h, int n_bindings, const char *const *slots, const int32_t *columns,
                              int n_columns)
{
    if (!h) return RSCM_B200_EINVAL;
    std::vector<int> slot_col(h->g.n_slots, -1), init_col(h->g.n_cells, -1);
    for (int i = 0; i < n_bindings; ++i) {
        if (columns[i] < 0 || columns[i] >= n_columns) return fail(h, RSCM_B200_EINVAL, "binding column out of range");
        std::string err;
        const int s = h->g.resolve_slot(slots[i], err);
        if (s == -1000000000) return fail(h, RSCM_B200_EINVAL, err);
        if (s >= 0) slot_col[s] = columns[i];
        else {
            const int cell0 = -(s + 1);
            const rscm::Variable &var = h->g.vars[h->g.cell_var[cell0]];
            // a scalar initial value is broadcast to every region (builder.rs:797-803,821-824)
            for (int r = 0; r < var.n_regions; ++r) init_col[cell0 + r] = columns[i];
        }
    }
    h->slot_col = slot_col;
    h->init_col = init_col;
    if (n_columns != h->n_cols && h->n_priors > 0) {
        // the priors were given per column of the previous binding: a different column count invalidates them
        if (h->d_priors) cudaFree(h->d_priors);
        h->d_priors = nullptr;
        h->n_priors = 0;
    }
    h->n_cols = n_columns;
    return RSCM_B200_OK;
}

int rscm_b200_select_outputs(rscm_b200_ensemble *h, int n_vars, const int32_t *vars, int32_t t_start, int32_t t_stop,
                             int32_t t_step)
{
    if (!h) return RSCM_B200_EINVAL;
    if (t_step < 1 || t_start < 0 || t_stop > h->g.T || t_start > t_stop) return fail(h, RSCM_B200_EINVAL, "bad time selection");
    std::vector<int> sel;
    for (int i = 0; i < n_vars; ++i) {
        if (vars[i] < 0 || vars[i] >= static_cast<int>(h->g.vars.size())) return fail(h, RSCM_B200_EINVAL, "bad variable index");
        for (int s : sel) if (s == vars[i]) return fail(h, RSCM_B200_EINVAL, "variable selected twice");
        sel.push_back(vars[i]);
    }
    h->sel_vars = sel;
    h->t_start = t_start;
    h->t_stop = t_stop;
    h->t_step = t_step;
    recompute_selection(h);
    return RSCM_B200_OK;
}

int64_t rscm_b200_output_rows(const rscm_b200_ensemble *h) { return h->rows; }

int rscm_b200_set_target(rscm_b200_ensemble *h, const rscm_b200_obs *obs, int64_t n_obs, int normalize)
{
    if (!h) return RSCM_B200_EINVAL;
    const rscm::Graph &g = h->g;
    // dense tables: one row per observed variable (a repeated (variable, time) opens another row)
    std::vector<int> row_var;
    std::vector<std::vector<double>> val, sig;
    for (int64_t i = 0; i < n_obs; ++i) {
        const rscm_b200_obs &o = obs[i];
        if (o.variable < 0 || o.variable >= static_cast<int>(g.vars.size())) return fail(h, RSCM_B200_EINVAL, "observation: bad variable");
        if (g.vars[o.variable].grid != RSCM_B200_SCALAR)
            return fail(h, RSCM_B200_EINVAL, "observation on a grid variable (reference: 'Grid variables not yet supported')");
        if (!(o.sigma > 0.0)) return fail(h, RSCM_B200_EINVAL, "observation uncertainty must be positive");
        if (o.time_index < 0 || o.time_index >= g.T) return fail(h, RSCM_B200_EINVAL, "observation time index out of range (reference: missing time => Err)");
        int row = -1;
        for (size_t r = 0; r < row_var.size(); ++r)
            if (row_var[r] == o.variable && sig[r][o.time_index] == 0.0) { row = static_cast<int>(r); break; }
        if (row < 0) {
            if (row_var.size() >= static_cast<size_t>(rscm_dev::MAX_OBS_ROWS))
                return fail(h, RSCM_B200_EUNSUPPORTED, "too many observed variables / duplicate observation times");
            row_var.push_back(o.variable);
            val.push_back(std::vector<double>(h->Tpad, 0.0));
            sig.push_back(std::vector<double>(h->Tpad, 0.0));
            row = static_cast<int>(row_var.size()) - 1;
        }
        val[row][o.time_index] = o.value;
        sig[row][o.time_index] = o.sigma;
    }
    h->n_obs_rows = static_cast<int>(row_var.size());
    for (int r = 0; r < h->n_obs_rows; ++r) h->obs_cell[r] = g.vars[row_var[r]].cell0;
    h->normalize = normalize ? 1 : 0;
    if (h->device < 0) { h->has_target = true; return RSCM_B200_OK; }
    if (h->d_obs) cudaFree(h->d_obs);
    h->d_obs = nullptr;
    if (h->n_obs_rows > 0) {
        std::vector<double> flat;
        for (auto &v : val) flat.insert(flat.end(), v.begin(), v.end());
        for (auto &s : sig) flat.insert(flat.end(), s.begin(), s.end());
        CU(cudaMalloc(&h->d_obs, flat.size() * sizeof(double)));
        CU(cudaMemcpy(h->d_obs, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->has_target = true;
    return RSCM_B200_OK;
}

int rscm_b200_set_priors(rscm_b200_ensemble *h, const rscm_b200_prior *priors, int n_columns)
{
    if (!h) return RSCM_B200_EINVAL;
    if (h->d_priors) cudaFree(h->d_priors);
    h->d_priors = nullptr;
    h->n_priors = 0;
    if (n_columns <= 0) return RSCM_B200_OK;
    if (n_columns != h->n_cols) return fail(h, RSCM_B200_EINVAL, "one prior per bound parameter column required");
    if (h->device < 0) { h->n_priors = n_columns; return RSCM_B200_OK; }
    std::vector<rscm_dev::PriorDev> p(n_columns);
    for (int i = 0; i < n_columns; ++i) {
        p[i].kind = priors[i].kind; p[i].pad = 0;
        p[i].a = priors[i].a; p[i].b = priors[i].b; p[i].low = priors[i].low; p[i].high = priors[i].high;
    }
    CU(cudaMalloc(&h->d_priors, p.size() * sizeof(rscm_dev::PriorDev)));
    CU(cudaMemcpy(h->d_priors, p.data(), p.size() * sizeof(rscm_dev::PriorDev), cudaMemcpyHostToDevice));
    h->n_priors = n_columns;
    return RSCM_B200_OK;
}

int rscm_b200_run_device(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                         const double *scenarios, int64_t S, double *out, uint8_t *status, void *stream)
{
    if (!h || !out) return fail(h, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return fail(h, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    CU(cudaSetDevice(h->device));
    return enqueue(h, params, M, params_layout, scenarios, S, out, status, nullptr, nullptr, true, false,
                   static_cast<cudaStream_t>(stream), true);
}

int rscm_b200_logpost_device(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                             const double *scenarios, int64_t S, double *logpost,
                             rscm_b200_logpost_summary *summary, void *stream)
{
    if (!h || !logpost) return fail(h, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return fail(h, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    CU(cudaSetDevice(h->device));
    static_assert(sizeof(rscm_b200_logpost_summary) == sizeof(rscm_dev::SummaryDev), "summary layout");
    return enqueue(h, params, M, params_layout, scenarios, S, nullptr, nullptr, logpost,
                   reinterpret_cast<rscm_dev::SummaryDev *>(summary), false, true, static_cast<cudaStream_t>(stream), true);
}

int rscm_b200_run_host(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                       const double *scenarios, int64_t S, double *out, uint8_t *status)
{
    if (!h || !out) return fail(h, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return fail(h, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    CU(cudaSetDevice(h->device));
    const rscm::Graph &g = h->g;
    if (M <= 0) return fail(h, RSCM_B200_EINVAL, "M must be positive");
    if (g.n_exo_rows == 0 && S <= 0) S = 1;
    if (g.n_exo_rows > 0 && (S <= 0 || !scenarios)) return fail(h, RSCM_B200_EINVAL, "this graph needs scenarios for its exogenous variables");
    for (int i = 0; i < 2; ++i)
        if (!h->streams[i]) CU(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    // scenarios: one H2D + one pack, shared by all chunks
    if (g.n_exo_rows > 0) {
        if (ensure(h, &h->d_scen, &h->cap_scen, S * h->scen_stride)) return RSCM_B200_ECUDA;
        CU(cudaMemcpyAsync(h->d_scen, scenarios, static_cast<size_t>(S) * h->scen_stride * 8, cudaMemcpyHostToDevice, h->streams[0]));
    }
    // chunk size: bound each device output buffer to ~1 GiB and keep >= 4 chunks in flight for overlap
    const int64_t bytes_per_member = std::max<int64_t>(1, h->rows * S * 8);
    int64_t Mc = std::max<int64_t>(rscm_dev::BLOCK, (int64_t(1) << 30) / bytes_per_member);
    Mc = std::min(Mc, std::max<int64_t>(rscm_dev::BLOCK, (M + 3) / 4));
    Mc = (Mc + rscm_dev::BLOCK - 1) / rscm_dev::BLOCK * rscm_dev::BLOCK;
    Mc = std::min(Mc, M);
    // capacities are shared between the two pipeline buffers: allocate both explicitly
    {
        const int64_t need_p = std::max<int64_t>(1, Mc * std::max(1, h->n_cols));
        const int64_t need_o = std::max<int64_t>(1, h->rows * S * Mc);
        const int64_t need_s = std::max<int64_t>(1, S * Mc);
        if (need_p > h->cap_params || need_o > h->cap_out || need_s > h->cap_status) {
            for (int i = 0; i < 2; ++i) {
                cudaFree(h->d_params[i]); cudaFree(h->d_out[i]); cudaFree(h->d_status[i]);
                h->d_params[i] = nullptr; h->d_out[i] = nullptr; h->d_status[i] = nullptr;
                CU(cudaMalloc(&h->d_params[i], static_cast<size_t>(need_p) * 8));
                CU(cudaMalloc(&h->d_out[i], static_cast<size_t>(need_o) * 8));
                CU(cudaMalloc(&h->d_status[i], static_cast<size_t>(need_s)));
            }
            h->cap_params = need_p; h->cap_out = need_o; h->cap_status = need_s;
        }
    }
    cudaEvent_t scen_ready;
    CU(cudaEventCreateWithFlags(&scen_ready, cudaEventDisableTiming));
    bool first = true;
    int rc = RSCM_B200_OK;
    int k = 0;
    for (int64_t m0 = 0; m0 < M && rc == RSCM_B200_OK; m0 += Mc, ++k) {
        const int64_t mc = std::min(Mc, M - m0);
        const int b = k & 1;
        cudaStream_t st = h->streams[b];
        if (!first) CU(cudaStreamWaitEvent(st, scen_ready, 0));
        if (h->n_cols > 0) {
            if (params_layout == 0)
                CU(cudaMemcpy2DAsync(h->d_params[b], static_cast<size_t>(mc) * 8, params + m0, static_cast<size_t>(M) * 8,
                                     static_cast<size_t>(mc) * 8, h->n_cols, cudaMemcpyHostToDevice, st));
            else
                CU(cudaMemcpyAsync(h->d_params[b], params + m0 * h->n_cols, static_cast<size_t>(mc) * h->n_cols * 8,
                                   cudaMemcpyHostToDevice, st));
        }
        rc = enqueue(h, h->d_params[b], mc, params_layout, h->d_scen, S, h->d_out[b], status ? h->d_status[b] : nullptr,
                     nullptr, nullptr, true, false, st, first);
        if (rc != RSCM_B200_OK) break;
        if (first) { CU(cudaEventRecord(scen_ready, st)); first = false; }
        // device chunk [rows][S][mc] -> host [rows][S][M] columns m0..m0+mc
        CU(cudaMemcpy2DAsync(out + m0, static_cast<size_t>(M) * 8, h->d_out[b], static_cast<size_t>(mc) * 8,
                             static_cast<size_t>(mc) * 8, static_cast<size_t>(h->rows * S), cudaMemcpyDeviceToHost, st));
        if (status)
            CU(cudaMemcpy2DAsync(status + m0, static_cast<size_t>(M), h->d_status[b], static_cast<size_t>(mc),
                                 static_cast<size_t>(mc), static_cast<size_t>(S), cudaMemcpyDeviceToHost, st));
    }
    cudaError_t e0 = cudaStreamSynchronize(h->streams[0]);
    cudaError_t e1 = cudaStreamSynchronize(h->streams[1]);
    cudaEventDestroy(scen_ready);
    if (rc != RSCM_B200_OK) return rc;
    if (e0 != cudaSuccess || e1 != cudaSuccess)
        return fail(h, RSCM_B200_ECUDA, std::string("pipeline: ") + cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    return RSCM_B200_OK;
}

int rscm_b200_logpost_host(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                           const double *scenarios, int64_t S, double *logpost, rscm_b200_logpost_summary *summary)
{
    if (!h || !logpost) return fail(h, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return fail(h, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    CU(cudaSetDevice(h->device));
    const rscm::Graph &g = h->g;
    if (M <= 0) return fail(h, RSCM_B200_EINVAL, "M must be positive");
    if (g.n_exo_rows == 0 && S <= 0) S = 1;
    if (g.n_exo_rows > 0 && (S <= 0 || !scenarios)) return fail(h, RSCM_B200_EINVAL, "this graph needs scenarios for its exogenous variables");
    if (!h->streams[0]) CU(cudaStreamCreateWithFlags(&h->streams[0], cudaStreamNonBlocking));
    cudaStream_t st = h->streams[0];
    if (g.n_exo_rows > 0) {
        if (ensure(h, &h->d_scen, &h->cap_scen, S * h->scen_stride)) return RSCM_B200_ECUDA;
        CU(cudaMemcpyAsync(h->d_scen, scenarios, static_cast<size_t>(S) * h->scen_stride * 8, cudaMemcpyHostToDevice, st));
    }
    const int64_t need_p = std::max<int64_t>(1, M * std::max(1, h->n_cols));
    if (need_p > h->cap_params) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->d_params[i]); h->d_params[i] = nullptr;
            CU(cudaMalloc(&h->d_params[i], static_cast<size_t>(need_p) * 8));
        }
        h->cap_params = need_p;
    }
    if (ensure(h, &h->d_logpost, &h->cap_logpost, S * M)) return RSCM_B200_ECUDA;
    if (!h->d_summary) CU(cudaMalloc(&h->d_summary, sizeof(rscm_dev::SummaryDev)));
    if (h->n_cols > 0)
        CU(cudaMemcpyAsync(h->d_params[0], params, static_cast<size_t>(M) * h->n_cols * 8, cudaMemcpyHostToDevice, st));
    int rc = enqueue(h, h->d_params[0], M, params_layout, h->d_scen, S, nullptr, nullptr, h->d_logpost,
                     summary ? h->d_summary : nullptr, false, true, st, true);
    if (rc != RSCM_B200_OK) return rc;
    CU(cudaMemcpyAsync(logpost, h->d_logpost, static_cast<size_t>(S) * M * 8, cudaMemcpyDeviceToHost, st));
    if (summary) CU(cudaMemcpyAsync(summary, h->d_summary, sizeof(rscm_dev::SummaryDev), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RSCM_B200_OK;
}

int64_t rscm_b200_launch_count(const rscm_b200_ensemble *h) { return h ? h->launches : 0; }

int64_t rscm_b200_shared_bytes(const rscm_b200_ensemble *h, int log_posterior) { return h ? static_cast<int64_t>(smem_bytes(h, log_posterior != 0)) : 0; }

double rscm_b200_kernel_ms(rscm_b200_ensemble *h, int reset)
{
    if (!h) return 0.0;
    for (size_t i = 0; i < h->events.size(); ++i) {
        if (!h->ev_pending[i]) continue;
        cudaEventSynchronize(h->events[i].second);
        harvest_event(h, i);
    }
    const double avg = h->kernel_ms_n ? h->kernel_ms_sum / static_cast<double>(h->kernel_ms_n) : 0.0;
    if (reset) { h->kernel_ms_sum = 0.0; h->kernel_ms_n = 0; }
    return avg;
}

int rscm_b200_measure_fma_peak(int device, int dtype, double *tflops)
{
    rscm_b200_ensemble *h = nullptr;
    if (!tflops) return fail(nullptr, RSCM_B200_EINVAL, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, RSCM_B200_ENODEVICE, "no CUDA device available");
    }
    if (device >= 0) CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaGetDeviceProperties(&prop, dev));
    void *sink = nullptr;
    CU(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double best = 0.0;
    int iters = 2000;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0));
        if (dtype == 0) rscm_dev::fma_peak_kernel<double><<<blocks, threads>>>(static_cast<double *>(sink), iters, 1.0);
        else rscm_dev::fma_peak_kernel<float><<<blocks, threads>>>(static_cast<float *>(sink), iters, 1.0f);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 16.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        if (ms < 20.f) iters *= 4;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return RSCM_B200_OK;
}

int rscm_b200_interpolate_device(const double *d_src_times, int64_t K, const double *d_src_values, int64_t n_series, int R,
                                 const double *d_dst_times, int64_t T, int strategy, double *d_out, void *stream)
{
    rscm_b200_ensemble *h = nullptr;
    if (!d_src_times || !d_src_values || !d_dst_times || !d_out) return fail(nullptr, RSCM_B200_EINVAL, "interpolate: null argument");
    if (K < 2 || T < 1 || n_series < 1 || (R != 1 && R != 2 && R != 4)) return fail(nullptr, RSCM_B200_EINVAL, "interpolate: need at least 2 source times, 1, 2 or 4 regions");
    if (strategy < 0 || strategy > 2) return fail(nullptr, RSCM_B200_EINVAL, "interpolate: strategy is 0 (Linear), 1 (Next) or 2 (Previous)");
    const long long total = n_series * T;
    const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 16));
    rscm_dev::interpolate_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_src_times, K, d_src_values, n_series, R, d_dst_times, T,
                                                                                      strategy, d_out);
    CU(cudaGetLastError());
    return RSCM_B200_OK;
}

int rscm_b200_device_math(int op, const double *d_x, int64_t n, double *d_y, void *stream)
{
    rscm_b200_ensemble *h = nullptr;
    if (op < 0 || op > 2 || !d_x || !d_y || n < 0) return fail(nullptr, RSCM_B200_EINVAL, "device_math: bad argument");
    if (n == 0) return RSCM_B200_OK;
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 148 * 8));
    rscm_dev::device_math_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(op, d_x, n, d_y);
    CU(cudaGetLastError());
    return RSCM_B200_OK;
}

// ---- stretch move (sampler/moves.rs, sampler/ensemble.rs:489-546) -------------------------------------------------
int rscm_b200_stretch_propose(const double *d_positions, int64_t ld, int n_cols, int64_t active_begin, int64_t n_active,
                              int64_t comp_begin, int64_t n_comp, double a, uint64_t seed, uint32_t step, double *d_proposals,
                              int64_t ld_proposals, double *d_z, void *stream)
{
    rscm_b200_ensemble *h = nullptr;
    if (!d_positions || !d_proposals || !d_z) return fail(nullptr, RSCM_B200_EINVAL, "null argument");
    if (n_cols < 1 || n_active < 1 || n_comp < 1 || ld < 1 || ld_proposals < n_active)
        return fail(nullptr, RSCM_B200_EINVAL, "stretch_propose: empty walker set or bad leading dimension");
    if (!(a > 1.0)) return fail(nullptr, RSCM_B200_EINVAL, "stretch parameter must be > 1"); // moves.rs:36-40
    const unsigned blocks = static_cast<unsigned>((n_active + 255) / 256);
    rscm_dev::stretch_propose_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_positions, ld, n_cols, active_begin, n_active, comp_begin, n_comp, a, seed, step, d_proposals, ld_proposals, d_z, nullptr);
    CU(cudaGetLastError());
    return RSCM_B200_OK;
}

int rscm_b200_stretch_accept(double *d_positions, int64_t ld, int n_cols, int64_t active_begin, int64_t n_active,
                             const double *d_proposals, int64_t ld_proposals, const double *d_z, const double *d_logpost_new,
                             double *d_logpost, uint64_t seed, uint32_t step, unsigned long long *d_n_accepted, void *stream)
{
    rscm_b200_ensemble *h = nullptr;
    if (!d_positions || !d_proposals || !d_z || !d_logpost_new || !d_logpost) return fail(nullptr, RSCM_B200_EINVAL, "null argument");
    if (n_cols < 1 || n_active < 1 || ld < 1 || ld_proposals < n_active)
        return fail(nullptr, RSCM_B200_EINVAL, "stretch_accept: empty walker set or bad leading dimension");
    const unsigned blocks = static_cast<unsigned>((n_active + 255) / 256);
    rscm_dev::stretch_accept_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_positions, ld, n_cols, active_begin, n_active, d_proposals, ld_proposals, d_z, d_logpost_new, d_logpost, seed, step,
        d_n_accepted, nullptr);
    CU(cudaGetLastError());
    return RSCM_B200_OK;
}

// ---- across-member quantiles of an output block (summary.cuh) -------------------------------------------------------
int rscm_b200_member_quantiles(const double *d_out, int64_t rows, int64_t S, int64_t M, const double *q, int nq, double *d_result,
                               void *stream)
{
    rscm_b200_ensemble *h = nullptr;
    if (!d_out || !q || !d_result) return fail(nullptr, RSCM_B200_EINVAL, "null argument");
    if (rows < 1 || S < 1 || M < 1) return fail(nullptr, RSCM_B200_EINVAL, "member_quantiles: empty output block");
    if (nq < 1 || nq > rscm_dev::Q_MAXQ) return fail(nullptr, RSCM_B200_EINVAL, "member_quantiles: 1 to 5 quantiles per call");
    if (rows * S > 2147483647LL) return fail(nullptr, RSCM_B200_EINVAL, "member_quantiles: too many (row, scenario) segments");
    rscm_dev::QArgs a{};
    a.data = d_out;
    a.M = M;
    a.runs = S * M;
    a.S = static_cast<int>(S);
    a.nq = nq;
    for (int k = 0; k < nq; ++k) {
        if (!(q[k] >= 0.0 && q[k] <= 1.0)) return fail(nullptr, RSCM_B200_EINVAL, "quantiles must be in the range [0, 1]");
        a.q[k] = q[k];
    }
    a.result = d_result;
    a.rows = rows;
    const size_t smem = sizeof(rscm_dev::QShared);
    CU(cudaFuncSetAttribute(rscm_dev::member_quantiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    rscm_dev::member_quantiles_kernel<<<static_cast<unsigned>(rows * S), rscm_dev::Q_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    CU(cudaGetLastError());
    return RSCM_B200_OK;
}

// =====================================================================================================================
// Multi-GPU: member sharding + the all-gather of per-member log-posteriors
// =====================================================================================================================
namespace {

int cfail(rscm_b200_comm *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    g_global_err = msg;
    return code;
}

#define CUC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return cfail(c, RSCM_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)
#define NCC(call)                                                                                        \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0) return cfail(c, RSCM_B200_ECOMM, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
    } while (0)

void shard_of(int64_t M, int rank, int world, int64_t *lo, int64_t *hi)
{
    *lo = (static_cast<int64_t>(rank) * M) / world;
    *hi = (static_cast<int64_t>(rank + 1) * M) / world;
}

// map `bytes` of fresh device memory into every peer: cudaMalloc + cudaIpcGetMemHandle, handles exchanged with NCCL
int symmetric_alloc(rscm_b200_comm *c, size_t bytes, rscm_b200_comm::Segment *out)
{
    NcclApi *nc = nccl_api();
    rscm_b200_comm::Segment seg;
    seg.bytes = bytes;
    CUC(cudaMalloc(&seg.local, bytes));
    CUC(cudaMemset(seg.local, 0, bytes));
    seg.peer[c->rank] = seg.local;
    if (c->world > 1 && c->p2p) {
        cudaIpcMemHandle_t mine;
        bool ok = cudaIpcGetMemHandle(&mine, seg.local) == cudaSuccess;
        if (!ok) cudaGetLastError();
        // exchange (handle, ok) of every rank
        const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
        std::vector<char> host(rec * c->world, 0);
        std::memcpy(host.data() + rec * c->rank, &mine, sizeof mine);
        host[rec * c->rank + sizeof mine] = ok ? 1 : 0;
        char *d_x = nullptr;
        CUC(cudaMalloc(&d_x, rec * c->world));
        CUC(cudaMemcpy(d_x, host.data(), rec * c->world, cudaMemcpyHostToDevice));
        NCC(nc->AllGather(d_x + rec * c->rank, d_x, rec, NCCL_INT8, c->nccl, nullptr));
        CUC(cudaStreamSynchronize(nullptr));
        CUC(cudaMemcpy(host.data(), d_x, rec * c->world, cudaMemcpyDeviceToHost));
        cudaFree(d_x);
        bool all = true;
        for (int p = 0; p < c->world; ++p) all = all && host[rec * p + sizeof mine];
        for (int p = 0; p < c->world && all; ++p) {
            if (p == c->rank) continue;
            cudaIpcMemHandle_t hp;
            std::memcpy(&hp, host.data() + rec * p, sizeof hp);
            if (cudaIpcOpenMemHandle(&seg.peer[p], hp, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                all = false;
            }
        }
        // every rank must take the same path: agree on the outcome
        char *d_ok = nullptr;
        CUC(cudaMalloc(&d_ok, c->world));
        const char mine_ok = all ? 1 : 0;
        CUC(cudaMemcpy(d_ok + c->rank, &mine_ok, 1, cudaMemcpyHostToDevice));
        NCC(nc->AllGather(d_ok + c->rank, d_ok, 1, NCCL_INT8, c->nccl, nullptr));
        CUC(cudaStreamSynchronize(nullptr));
        std::vector<char> oks(c->world);
        CUC(cudaMemcpy(oks.data(), d_ok, c->world, cudaMemcpyDeviceToHost));
        cudaFree(d_ok);
        for (char o : oks) all = all && o;
        if (!all) {
            for (int p = 0; p < c->world; ++p)
                if (p != c->rank && seg.peer[p]) { cudaIpcCloseMemHandle(seg.peer[p]); seg.peer[p] = nullptr; }
            cudaGetLastError();
            c->p2p = false; // NCCL carries the all-gather instead
        }
    }
    *out = seg;
    return RSCM_B200_OK;
}

const rscm_b200_comm::Segment *segment_of(const rscm_b200_comm *c, const void *p, size_t bytes)
{
    for (size_t i = 1; i < c->segments.size(); ++i) {
        const auto &sg = c->segments[i];
        const char *b = static_cast<const char *>(sg.local), *q = static_cast<const char *>(p);
        if (q >= b && q + bytes <= b + sg.bytes) return &sg;
    }
    return nullptr;
}

// all-gather of member blocks inside `buf` [S][M] (block of rank r = columns shard_of(r)), in place, over NCCL
int allgather_in_place(rscm_b200_comm *c, double *buf, int64_t M, int64_t S, cudaStream_t st)
{
    NcclApi *nc = nccl_api();
    if (c->world == 1) return RSCM_B200_OK;
    if (S == 1 && M % c->world == 0) {
        const int64_t n = M / c->world;
        NCC(nc->AllGather(buf + n * c->rank, buf, static_cast<size_t>(n), NCCL_FLOAT64, c->nccl, st));
        return RSCM_B200_OK;
    }
    // ragged blocks or several scenario rows: one broadcast per (row, owner) in a group
    NCC(nc->GroupStart());
    for (int64_t s = 0; s < S; ++s)
        for (int r = 0; r < c->world; ++r) {
            int64_t lo, hi;
            shard_of(M, r, c->world, &lo, &hi);
            if (hi > lo) {
                int rc = nc->Broadcast(buf + s * M + lo, buf + s * M + lo, static_cast<size_t>(hi - lo), NCCL_FLOAT64, r, c->nccl, st);
                if (rc != 0) { nc->GroupEnd(); return cfail(c, RSCM_B200_ECOMM, std::string("ncclBroadcast: ") + nc->GetErrorString(rc)); }
            }
        }
    NCC(nc->GroupEnd());
    return RSCM_B200_OK;
}

// one sharded log-posterior evaluation: kernel on this rank's member block + all-gather (fused over peer memory when the
// destination is symmetric memory, NCCL otherwise)
int logpost_sharded(rscm_b200_ensemble *h, rscm_b200_comm *c, const double *params, int64_t M, int layout, int64_t ld,
                    const double *scen, int64_t S, double *lp_global, cudaStream_t st, bool pack)
{
    int64_t lo, hi;
    shard_of(M, c->rank, c->world, &lo, &hi);
    const int64_t Ml = hi - lo;
    const int64_t S_eff = S > 0 ? S : 1;
    LaunchOpts o;
    o.lp_ld = M;
    const double *p_local = params;
    if (layout == 0) { o.ld_col = ld > 0 ? ld : M; p_local = params + lo; }
    else p_local = params + lo * h->n_cols;
    const rscm_b200_comm::Segment *sg = (c->world > 1 && c->p2p) ? segment_of(c, lp_global, static_cast<size_t>(S_eff * M) * 8) : nullptr;
    if (sg) {
        const size_t off = static_cast<const char *>(static_cast<const void *>(lp_global)) - static_cast<const char *>(sg->local);
        o.n_peers = c->world;
        for (int p = 0; p < c->world; ++p) {
            o.peer_lp[p] = reinterpret_cast<double *>(static_cast<char *>(sg->peer[p]) + off) + lo;
            o.peer_flag[p] = reinterpret_cast<unsigned long long *>(c->segments[0].peer[p]) + c->rank;
        }
        o.epoch = c->d_epoch;
    }
    if (Ml > 0) {
        int rc = enqueue(h, p_local, Ml, layout, scen, S, nullptr, nullptr, lp_global + lo, nullptr, false, true, st, pack, &o);
        if (rc != RSCM_B200_OK) { c->err = h->err; return rc; }
    } else if (sg) {
        return cfail(c, RSCM_B200_EINVAL, "fused all-gather needs at least one member per rank");
    }
    if (sg) {
        rscm_dev::peer_wait_kernel<<<1, 32, 0, st>>>(c->flags, c->world, c->d_epoch, c->timeout_ns, c->d_error);
        CUC(cudaGetLastError());
        h->launches++;
        return RSCM_B200_OK;
    }
    return allgather_in_place(c, lp_global, M, S_eff, st);
}

} // namespace

int rscm_b200_comm_unique_id(void *unique_id)
{
    rscm_b200_comm *c = nullptr;
    NcclApi *nc = nccl_api();
    if (!nc) return cfail(nullptr, RSCM_B200_ECOMM, "NCCL is not available (set RSCM_B200_NCCL_LIB to libnccl.so.2)");
    if (!unique_id) return cfail(nullptr, RSCM_B200_EINVAL, "null argument");
    static_assert(RSCM_B200_UNIQUE_ID_BYTES == sizeof(NcclUniqueId), "unique id size");
    NCC(nc->GetUniqueId(static_cast<NcclUniqueId *>(unique_id)));
    return RSCM_B200_OK;
}

int rscm_b200_comm_init(const void *unique_id, int rank, int world, int device, rscm_b200_comm **out)
{
    if (!out) return cfail(nullptr, RSCM_B200_EINVAL, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return cfail(nullptr, RSCM_B200_EINVAL, "bad rank / world size");
    if (world > rscm_dev::MAX_PEERS) return cfail(nullptr, RSCM_B200_EUNSUPPORTED, "at most 8 ranks (one NVSwitch domain) per communicator");
    rscm_b200_comm *c = new rscm_b200_comm();
    c->rank = rank;
    c->world = world;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        delete c;
        return cfail(nullptr, RSCM_B200_ENODEVICE, "no CUDA device available");
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { delete c; return cfail(nullptr, RSCM_B200_ENODEVICE, "cannot select the requested CUDA device"); }
    cudaGetDevice(&c->device);
    if (world > 1) {
        NcclApi *nc = nccl_api();
        if (!nc || !unique_id) { delete c; return cfail(nullptr, RSCM_B200_ECOMM, "NCCL is not available or no unique id was given"); }
        NcclUniqueId id;
        std::memcpy(&id, unique_id, sizeof id);
        const int r = nc->CommInitRank(&c->nccl, world, id, rank);
        if (r != 0) { std::string m = std::string("ncclCommInitRank: ") + nc->GetErrorString(r); delete c; return cfail(nullptr, RSCM_B200_ECOMM, m); }
        c->p2p = getenv("RSCM_B200_NO_P2P") == nullptr;
    }
    // control block: per-peer arrival flags, the epoch counter and the error word
    rscm_b200_comm::Segment ctl;
    int rc = symmetric_alloc(c, 256, &ctl);
    if (rc != RSCM_B200_OK) { std::string m = c->err; rscm_b200_comm_destroy(c); return cfail(nullptr, rc, m); }
    c->segments.push_back(ctl);
    c->flags = static_cast<unsigned long long *>(ctl.local);
    c->d_epoch = c->flags + rscm_dev::MAX_PEERS;
    c->d_error = reinterpret_cast<int *>(c->flags + rscm_dev::MAX_PEERS + 1);
    if (cudaMalloc(&c->d_iteration, sizeof(unsigned)) != cudaSuccess) { rscm_b200_comm_destroy(c); return cfail(nullptr, RSCM_B200_ECUDA, "device allocation failed"); }
    cudaMemset(c->d_iteration, 0, sizeof(unsigned));
    *out = c;
    return RSCM_B200_OK;
}

void rscm_b200_comm_destroy(rscm_b200_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (auto &sg : c->segments) {
        for (int p = 0; p < c->world; ++p)
            if (p != c->rank && sg.peer[p]) cudaIpcCloseMemHandle(sg.peer[p]);
        cudaFree(sg.local);
    }
    cudaFree(c->d_iteration);
    if (c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
    cudaGetLastError();
    delete c;
}

const char *rscm_b200_comm_last_error(const rscm_b200_comm *c) { return c ? c->err.c_str() : g_global_err.c_str(); }
int rscm_b200_comm_rank(const rscm_b200_comm *c) { return c ? c->rank : 0; }
int rscm_b200_comm_world(const rscm_b200_comm *c) { return c ? c->world : 1; }
int rscm_b200_comm_peer_access(const rscm_b200_comm *c) { return (c && c->world > 1 && c->p2p) ? 1 : 0; }

int rscm_b200_comm_shard(const rscm_b200_comm *c, int64_t M, int64_t *begin, int64_t *end)
{
    if (!c || !begin || !end) return RSCM_B200_EINVAL;
    shard_of(M, c->rank, c->world, begin, end);
    return RSCM_B200_OK;
}

int rscm_b200_comm_symmetric_alloc(rscm_b200_comm *c, size_t bytes, void **d_ptr)
{
    if (!c || !d_ptr || bytes == 0) return cfail(c, RSCM_B200_EINVAL, "null argument");
    CUC(cudaSetDevice(c->device));
    rscm_b200_comm::Segment sg;
    int rc = symmetric_alloc(c, (bytes + 255) & ~size_t(255), &sg);
    if (rc != RSCM_B200_OK) return rc;
    c->segments.push_back(sg);
    *d_ptr = sg.local;
    return RSCM_B200_OK;
}

int rscm_b200_allgather_f64(rscm_b200_comm *c, const double *d_local, int64_t n_local, double *d_global, void *stream)
{
    if (!c || !d_global || (n_local > 0 && !d_local) || n_local < 0) return cfail(c, RSCM_B200_EINVAL, "null argument");
    CUC(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c->world == 1) {
        if (d_local != d_global) CUC(cudaMemcpyAsync(d_global, d_local, static_cast<size_t>(n_local) * 8, cudaMemcpyDeviceToDevice, st));
        return RSCM_B200_OK;
    }
    NCC(nccl_api()->AllGather(d_local, d_global, static_cast<size_t>(n_local), NCCL_FLOAT64, c->nccl, st));
    return RSCM_B200_OK;
}

int rscm_b200_logpost_sharded_device(rscm_b200_ensemble *h, rscm_b200_comm *c, const double *params, int64_t M, int params_layout,
                                     const double *scenarios, int64_t S, double *logpost_global, void *stream)
{
    if (!h || !c || !logpost_global) return cfail(c, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return cfail(c, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    if (h->device != c->device) return cfail(c, RSCM_B200_EINVAL, "ensemble and communicator live on different devices");
    if (M < 1) return cfail(c, RSCM_B200_EINVAL, "M must be positive");
    CUC(cudaSetDevice(h->device));
    return logpost_sharded(h, c, params, M, params_layout, 0, scenarios, S, logpost_global, static_cast<cudaStream_t>(stream), true);
}

int rscm_b200_comm_check(rscm_b200_comm *c)
{
    if (!c) return RSCM_B200_EINVAL;
    int e = 0;
    CUC(cudaMemcpy(&e, c->d_error, sizeof e, cudaMemcpyDeviceToHost));
    if (e) return cfail(c, RSCM_B200_ECOMM, "a peer did not arrive at the fused all-gather within the timeout");
    return RSCM_B200_OK;
}

// ---- the sampler loop behind the ABI: n iterations of (propose -> sharded log-posterior -> accept) x 2 halves, each
// iteration replayed from one CUDA graph (sampler/ensemble.rs:412-546) -------------------------------------------------------
int rscm_b200_sampler_iterate(rscm_b200_ensemble *h, rscm_b200_comm *c, const rscm_b200_sampler_state *s, const double *scenarios,
                              int64_t S, int n_iterations, int use_graph, void *stream)
{
    if (!h || !c || !s) return cfail(c, RSCM_B200_EINVAL, "null argument");
    if (h->device < 0) return cfail(c, RSCM_B200_ENODEVICE, "host-only handle (device = -2) cannot run; there is no CPU fallback");
    if (!s->positions || !s->logpost || !s->proposals || !s->z || !s->logpost_new[0] || !s->logpost_new[1])
        return cfail(c, RSCM_B200_EINVAL, "sampler state has null buffers");
    const int64_t W = s->n_walkers, half = W / 2;
    if (W < 2 || (W & 1)) return cfail(c, RSCM_B200_EINVAL, "the number of walkers must be even and at least 2"); // ensemble.rs:420-431
    if (s->n_cols != h->n_cols) return cfail(c, RSCM_B200_EINVAL, "sampler state and parameter binding disagree on the column count");
    if (!(s->a > 1.0)) return cfail(c, RSCM_B200_EINVAL, "stretch parameter must be > 1");
    if (S > 1) return cfail(c, RSCM_B200_EINVAL, "the sampler evaluates one scenario per walker");
    if (n_iterations < 1) return RSCM_B200_OK;
    CUC(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>((half + 255) / 256);
    const unsigned first = s->first_iteration;

    // one iteration = two half-updates; `iter` is either the host's count (eager) or the device counter (graph)
    auto one_iteration = [&](unsigned step_base, const unsigned *d_iter, bool pack) -> int {
        for (int hidx = 0; hidx < 2; ++hidx) {
            const int64_t a0 = hidx == 0 ? 0 : half, c0 = hidx == 0 ? half : 0;
            const unsigned step = step_base + static_cast<unsigned>(hidx);
            double *lpn = s->logpost_new[hidx];
            rscm_dev::stretch_propose_kernel<<<blocks, 256, 0, st>>>(s->positions, s->ld, s->n_cols, a0, half, c0, half, s->a, s->seed, step,
                                                                     s->proposals, half, s->z, d_iter);
            CUC(cudaGetLastError());
            int rc = logpost_sharded(h, c, s->proposals, half, 0, half, scenarios, S, lpn, st, pack && hidx == 0);
            if (rc != RSCM_B200_OK) return rc;
            rscm_dev::stretch_accept_kernel<<<blocks, 256, 0, st>>>(s->positions, s->ld, s->n_cols, a0, half, s->proposals, half, s->z, lpn,
                                                                    s->logpost, s->seed, step, s->n_accepted, d_iter);
            CUC(cudaGetLastError());
            h->launches += 2;
        }
        return RSCM_B200_OK;
    };

    // Chain::push (sampler/chain.rs:63): every thin-th iteration's ensemble is kept, in device memory
    auto record = [&](unsigned iteration) -> int {
        if (!s->thin || !s->chain_positions || iteration % s->thin) return RSCM_B200_OK;
        const int64_t k = iteration / s->thin;
        if (k >= s->chain_capacity) return cfail(c, RSCM_B200_EINVAL, "chain buffers are too small for this iteration");
        CUC(cudaMemcpy2DAsync(s->chain_positions + k * s->n_cols * W, static_cast<size_t>(W) * 8, s->positions, static_cast<size_t>(s->ld) * 8,
                              static_cast<size_t>(W) * 8, static_cast<size_t>(s->n_cols), cudaMemcpyDeviceToDevice, st));
        if (s->chain_logpost)
            CUC(cudaMemcpyAsync(s->chain_logpost + k * W, s->logpost, static_cast<size_t>(W) * 8, cudaMemcpyDeviceToDevice, st));
        return RSCM_B200_OK;
    };

    int done = 0;
    // the first iteration of a call always runs eagerly: it packs the scenarios and sizes every buffer
    {
        int rc = one_iteration(2u * first, nullptr, true);
        if (rc == RSCM_B200_OK) rc = record(first);
        if (rc != RSCM_B200_OK) return rc;
        done = 1;
    }
    if (!use_graph) {
        for (; done < n_iterations; ++done) {
            int rc = one_iteration(2u * (first + static_cast<unsigned>(done)), nullptr, false);
            if (rc == RSCM_B200_OK) rc = record(first + static_cast<unsigned>(done));
            if (rc != RSCM_B200_OK) return rc;
        }
        return RSCM_B200_OK;
    }
    if (done == n_iterations) return RSCM_B200_OK;
    // graph: keyed by everything baked into the captured kernel arguments
    std::vector<uintptr_t> key = {reinterpret_cast<uintptr_t>(h), reinterpret_cast<uintptr_t>(s->positions), reinterpret_cast<uintptr_t>(s->logpost),
                                  reinterpret_cast<uintptr_t>(s->proposals), reinterpret_cast<uintptr_t>(s->z),
                                  reinterpret_cast<uintptr_t>(s->logpost_new[0]), reinterpret_cast<uintptr_t>(s->logpost_new[1]),
                                  reinterpret_cast<uintptr_t>(s->n_accepted), static_cast<uintptr_t>(s->ld), static_cast<uintptr_t>(W),
                                  static_cast<uintptr_t>(s->seed), reinterpret_cast<uintptr_t>(scenarios), static_cast<uintptr_t>(S),
                                  static_cast<uintptr_t>(h->n_obs_rows), static_cast<uintptr_t>(h->n_priors), reinterpret_cast<uintptr_t>(h->d_obs),
                                  reinterpret_cast<uintptr_t>(h->d_priors), static_cast<uintptr_t>(std::hash<double>()(s->a))};
    if (!c->graph_exec || key != c->graph_key) {
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        cudaStream_t cs;
        CUC(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaStream_t user = st;
        st = cs;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
        int rc = RSCM_B200_OK;
        if (e == cudaSuccess) {
            const int64_t before = h->launches;
            rc = one_iteration(0u, c->d_iteration, false);
            if (rc == RSCM_B200_OK) {
                rscm_dev::advance_iteration_kernel<<<1, 1, 0, cs>>>(c->d_iteration);
                if (cudaGetLastError() != cudaSuccess) rc = RSCM_B200_ECUDA;
            }
            h->launches = before; // counted per replay below
            e = cudaStreamEndCapture(cs, &graph);
        }
        st = user;
        cudaStreamDestroy(cs);
        if (rc != RSCM_B200_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || !graph) { cudaGetLastError(); return cfail(c, RSCM_B200_ECUDA, std::string("stream capture of the sampler iteration failed: ") + cudaGetErrorString(e)); }
        e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { c->graph_exec = nullptr; return cfail(c, RSCM_B200_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        c->graph_key = key;
    }
    const unsigned start = first + static_cast<unsigned>(done);
    CUC(cudaMemcpyAsync(c->d_iteration, &start, sizeof start, cudaMemcpyHostToDevice, st)); // pageable source: copied before return
    const bool fused = c->world > 1 && c->p2p && segment_of(c, s->logpost_new[0], static_cast<size_t>(half) * 8);
    const int per_iter = 2 * (3 + (fused ? 1 : 0)) + 1; // propose, log-posterior, (peer wait), accept per half + the counter
    for (; done < n_iterations; ++done) {
        CUC(cudaGraphLaunch(c->graph_exec, st));
        h->launches += per_iter;
        int rc = record(first + static_cast<unsigned>(done));
        if (rc != RSCM_B200_OK) return rc;
    }
    return RSCM_B200_OK;
}

} // extern "C"
