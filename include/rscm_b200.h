/*
 * rscm_b200.h — C ABI of the B200-native RSCM ensemble engine.
 *
 * This is the drop-in boundary for ONE path of lewisjared/rscm (v0.5.0): many
 * independent `Model::run` calls (parameter sets x scenarios) of a component
 * graph, i.e. `ModelRunner::run_batch` and what it calls.  A Rust `-sys` crate,
 * a ctypes/cffi module or C++ code binds exactly these symbols; there are no
 * torch / C++ types in any signature.  `INTEGRATION.md` shows the reference-side
 * bindings.  Citations are relative to the reference checkout.
 *
 * Conventions
 *  - every function returns 0 on success, a negative RSCM_B200_E* code on error;
 *    `rscm_b200_last_error(h)` (or `rscm_b200_last_global_error()` when no handle
 *    exists yet) returns the message.  Nothing throws or aborts across the ABI.
 *  - per-member numerical failure is DATA, not an error: outputs are NaN, the
 *    status byte is set, the log-posterior is -inf (reference:
 *    crates/rscm-core/src/model/runtime.rs:493-495,
 *    crates/rscm-calibrate/src/sampler/ensemble.rs:152-176).
 *  - the caller owns every buffer it passes; the handle owns device scratch and
 *    the compiled graph.  A handle is not re-entrant (one call in flight) but
 *    may be moved between host threads.
 *  - there is NO CPU fallback: creation fails with RSCM_B200_ENODEVICE when no
 *    CUDA device is usable, and RSCM_B200_EUNSUPPORTED for a graph the engine
 *    has no device code for.
 */
#ifndef RSCM_B200_H
#define RSCM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSCM_B200_ABI_VERSION 1

/* error codes */
#define RSCM_B200_OK 0
#define RSCM_B200_EINVAL (-1)       /* bad argument / graph description */
#define RSCM_B200_EUNSUPPORTED (-2) /* graph has no device program */
#define RSCM_B200_ENODEVICE (-3)    /* no usable CUDA device */
#define RSCM_B200_ECUDA (-4)        /* CUDA runtime error */
#define RSCM_B200_ENOMEM (-5)
#define RSCM_B200_ECOMM (-6)        /* NCCL / peer-memory error (multi-GPU entry points) */

/* Component kinds.  Replaces `Arc<dyn Component>` objects added through
 * ModelBuilder::with_component (crates/rscm-core/src/model/builder.rs:45-60);
 * parameter block order per kind is documented at rscm_b200_component_desc. */
typedef enum {
    RSCM_B200_TWO_LAYER = 1,    /* crates/rscm-two-layer/src/component.rs:38-90 */
    RSCM_B200_CARBON_CYCLE = 2, /* crates/rscm-components/src/components/carbon_cycle.rs:24-40 */
    RSCM_B200_CO2_ERF = 3,      /* crates/rscm-components/src/components/co2_erf.rs:18-25 */
    RSCM_B200_GHG_FORCING = 5,  /* crates/rscm-magicc/src/parameters/ghg_forcing.rs */
    RSCM_B200_OZONE_FORCING = 6,    /* crates/rscm-magicc/src/parameters/ozone_forcing.rs */
    RSCM_B200_AEROSOL_DIRECT = 7,   /* crates/rscm-magicc/src/parameters/aerosol.rs (AerosolDirectParameters) */
    RSCM_B200_AEROSOL_INDIRECT = 8, /* crates/rscm-magicc/src/parameters/aerosol.rs (AerosolIndirectParameters) */
    RSCM_B200_CLIMATE_UDEB = 9,     /* crates/rscm-magicc/src/parameters/climate_udeb.rs */
    RSCM_B200_FOUR_BOX_OHU = 10,    /* crates/rscm-components/src/components/four_box_ocean_heat_uptake.rs */
    RSCM_B200_OCEAN_SURFACE_PP = 11, /* .../ocean_carbon_cycle/ocean_surface_partial_pressure.rs */
    RSCM_B200_CO2_BUDGET = 12,      /* crates/rscm-magicc/src/parameters/co2_budget.rs */
    RSCM_B200_TERRESTRIAL_CARBON = 13, /* crates/rscm-magicc/src/parameters/terrestrial_carbon.rs */
    RSCM_B200_CH4_CHEMISTRY = 14,   /* crates/rscm-magicc/src/parameters/ch4_chemistry.rs */
    RSCM_B200_N2O_CHEMISTRY = 15,   /* crates/rscm-magicc/src/parameters/n2o_chemistry.rs */
    RSCM_B200_OCEAN_CARBON = 16,    /* crates/rscm-magicc/src/parameters/ocean_carbon.rs (IRF forms flattened, 60 values) */
    /* crates/rscm-magicc/src/parameters/halocarbon.rs with the reference's default species list (:203-262, 23 F-gases
     * then 18 Montreal gases): br_multiplier, cfc11_release_normalisation, eesc_delay, air_molar_mass,
     * atmospheric_mass_tg, mixing_box_fraction, then per species lifetime, radiative_efficiency, concentration_pi,
     * molecular_weight, n_cl, n_br, fractional_release (293 values; all per-graph, none bindable per member) */
    RSCM_B200_HALOCARBON_CHEMISTRY = 17
} rscm_b200_component_kind;

/* GridType — crates/rscm-core/src/component.rs:56-64 */
typedef enum { RSCM_B200_SCALAR = 0, RSCM_B200_FOUR_BOX = 1, RSCM_B200_HEMISPHERIC = 2 } rscm_b200_grid;
/* AggregateOp — crates/rscm-core/src/schema.rs:60-80 */
typedef enum { RSCM_B200_AGG_SUM = 0, RSCM_B200_AGG_MEAN = 1, RSCM_B200_AGG_WEIGHTED = 2 } rscm_b200_agg_op;
/* VariableSource — crates/rscm-core/src/state/mod.rs:157-170 */
typedef enum { RSCM_B200_SRC_EXOGENOUS = 0, RSCM_B200_SRC_OWN_STATE = 1, RSCM_B200_SRC_UPSTREAM = 2 } rscm_b200_source;

/* One component, in insertion order (the order decides variable-source
 * classification, builder.rs:465-485).  Parameter block order:
 *   TWO_LAYER    : lambda0, a, efficacy, eta, heat_capacity_surface, heat_capacity_deep
 *   CARBON_CYCLE : tau, conc_pi, alpha_temperature, step_size (SolverOptions, default 0.1)
 *   CO2_ERF      : erf_2xco2, conc_pi
 *   GHG_FORCING  : method(0 Ipcctar,1 Olbl), co2_pi, ch4_pi, n2o_pi, delq2xco2, ch4_radeff,
 *                  n2o_radeff, olbl_co2_a1,b1,c1,d1, olbl_ch4_a3,b3,d3, olbl_n2o_a2,b2,c2,d2,
 *                  adjust_co2, adjust_ch4, adjust_n2o
 *   OZONE_FORCING: eesc_reference, strat_o3_scale, strat_cl_exponent, trop_radeff, trop_oz_ch4,
 *                  trop_oz_nox, trop_oz_co, trop_oz_voc, ch4_pi, nox_pi, co_pi, nmvoc_pi,
 *                  temp_feedback_scale
 *   AEROSOL_DIRECT: sox/bc/oc/nitrate_coefficient, sox_regional[4], bc_regional[4], oc_regional[4],
 *                  nitrate_regional[4], sox_pi, bc_pi, oc_pi, nox_pi, harmonize, harmonize_year,
 *                  harmonize_target
 *   AEROSOL_INDIRECT: cloud_albedo_coefficient, reference_burden, sox_weight, oc_weight, sox_pi,
 *                  oc_pi, harmonize, harmonize_year, harmonize_target
 *   CLIMATE_UDEB : the fields of ClimateUDEBParameters in declaration order (rf_regions_co2
 *                  expanded to 4 values, booleans/enums as 0/1/2), see rscm_b200/magicc.py
 *   FOUR_BOX_OHU, OCEAN_SURFACE_PP, CO2_BUDGET, TERRESTRIAL_CARBON, CH4_CHEMISTRY, N2O_CHEMISTRY:
 *                  the fields of the reference's parameter struct in declaration order, arrays
 *                  expanded, booleans as 0/1 (rscm_b200/components.py, rscm_b200/magicc.py)
 *   OCEAN_CARBON : model id (0 3D-GFDL, 1 2D-BERN, 2 HILDA), co2_pi, pco2_pi, gas_exchange_scale,
 *                  gas_exchange_tau, temp_sensitivity, irf_scale, mixed_layer_depth, ocean_surface_area,
 *                  sst_pi, steps_per_year, max_history_months, irf_switch_time, then irf_early and
 *                  irf_late as {kind (0 Polynomial, 1 ExponentialSum), n terms, 8 coefficients,
 *                  8 timescales}, delta_ospp_offsets[5], delta_ospp_coefficients[5],
 *                  enable_temp_feedback (60 values)
 *   HALOCARBON_CHEMISTRY: see the enum above (293 values)
 */
typedef struct {
    int32_t kind;     /* rscm_b200_component_kind */
    int32_t n_params;
    const double *params;
} rscm_b200_component_desc;

/* VariableSchema::add_variable — crates/rscm-core/src/schema.rs */
typedef struct {
    const char *name;
    int32_t grid; /* rscm_b200_grid */
} rscm_b200_schema_variable;

/* VariableSchema::add_aggregate — schema.rs; contributors in declaration order;
 * list chained aggregates in dependency order. */
typedef struct {
    const char *name;
    int32_t op;   /* rscm_b200_agg_op */
    int32_t grid; /* rscm_b200_grid */
    int32_t n_contributors;
    const char *const *contributors;
    const double *weights; /* n_contributors, AGG_WEIGHTED only, else NULL */
} rscm_b200_aggregate_desc;

typedef struct {
    const char *name;
    double value;
} rscm_b200_initial_value;

/* unit conversion factor applied on read for one (variable, consuming component)
 * pair — model/runtime.rs:385-389; computed by the caller's unit registry. */
typedef struct {
    int32_t component; /* index into components[] */
    const char *variable;
    double factor;
} rscm_b200_unit_factor;

/* The description ModelBuilder holds when build() is called
 * (crates/rscm-core/src/model/builder.rs:29-41). */
typedef struct {
    int32_t abi_version; /* RSCM_B200_ABI_VERSION */
    int32_t n_components;
    const rscm_b200_component_desc *components;
    int32_t has_schema;
    int32_t n_schema_variables;
    const rscm_b200_schema_variable *schema_variables;
    int32_t n_aggregates;
    const rscm_b200_aggregate_desc *aggregates;
    int32_t n_initial_values;
    const rscm_b200_initial_value *initial_values;
    int32_t n_unit_factors;
    const rscm_b200_unit_factor *unit_factors;
    const double *four_box_weights;    /* 4 or NULL (default 0.25 each) */
    const double *hemispheric_weights; /* 2 or NULL (default 0.5 each) */
    int32_t n_times;                   /* T = time_axis.len() */
    const double *time_bounds;         /* T+1 (TimeAxis bounds, timeseries.rs:24-26) */
    int32_t compute_dtype;             /* 0 = fp64 (parity path), 1 = fp32 (1e-4 path) */
    int32_t device;                    /* CUDA device ordinal, -1 = current, -2 = host-only handle:
                                          graph compile + introspection only, every run call
                                          returns RSCM_B200_ENODEVICE (used by CPU-side tests) */
} rscm_b200_graph_desc;

typedef struct rscm_b200_ensemble rscm_b200_ensemble;

/* ---- lifetime ----------------------------------------------------------- */
/* Replaces ModelBuilder::build (builder.rs:418-860) once per ensemble instead
 * of once per member (model_runner.rs:233-235). */
int rscm_b200_ensemble_create(const rscm_b200_graph_desc *desc, rscm_b200_ensemble **out);
void rscm_b200_ensemble_destroy(rscm_b200_ensemble *h);
const char *rscm_b200_last_error(const rscm_b200_ensemble *h);
const char *rscm_b200_last_global_error(void);
int rscm_b200_abi_version(void);
int rscm_b200_device_count(void);

/* ---- introspection (Model::debug_info equivalents, model/debug.rs:315-319) -- */
int rscm_b200_n_variables(const rscm_b200_ensemble *h);
const char *rscm_b200_variable_name(const rscm_b200_ensemble *h, int v);
int rscm_b200_variable_grid(const rscm_b200_ensemble *h, int v);
int rscm_b200_variable_is_endogenous(const rscm_b200_ensemble *h, int v);
int rscm_b200_variable_index(const rscm_b200_ensemble *h, const char *name);
/* exogenous variables, in the order scenario arrays must list them */
int rscm_b200_n_exogenous(const rscm_b200_ensemble *h);
int rscm_b200_exogenous_variable(const rscm_b200_ensemble *h, int i);
/* nodes = components then aggregators; order[] receives node ids in BFS order
 * (model/runtime.rs:504-510); returns the count */
int rscm_b200_n_nodes(const rscm_b200_ensemble *h);
int rscm_b200_execution_order(const rscm_b200_ensemble *h, int *order, int capacity);
int rscm_b200_variable_source(const rscm_b200_ensemble *h, int component, const char *variable);
/* canonical text of the fused device program chosen for this graph; for a run-time compiled program this includes the
 * specialisation on the current parameter binding (unbound slots are literals), valid until the next call */
const char *rscm_b200_program_signature(const rscm_b200_ensemble *h);
/* 0: the program comes from the ahead-of-time registry; 1: compiled at run time (NVRTC, sm_100a) */
int rscm_b200_program_is_jit(const rscm_b200_ensemble *h);
/* index of the time point whose "{:.6}" key equals that of `time`
 * (crates/rscm-calibrate/src/likelihood.rs:40-42); -1 if none */
int rscm_b200_time_index(const rscm_b200_ensemble *h, double time);

/* ---- per-member parameter binding ---------------------------------------- */
/* Replaces the `factory(params) -> Model` closure of DefaultModelRunner
 * (model_runner.rs:116-129, 233-235): column j of the parameter matrix feeds
 * the named slot(s).  Slot names: "<ComponentType>.<field>" (first component
 * of that type), "<ComponentType>#<i>.<field>" (i-th component overall) or
 * "initial:<variable name>".  A column may appear several times (e.g. conc_pi of
 * CarbonCycle and CO2ERF).  Unbound slots keep the graph description's value.
 */
/* Rebinding with a different n_columns drops priors set for the previous binding
 * (they are per column): call rscm_b200_set_priors again. */
int rscm_b200_bind_parameters(rscm_b200_ensemble *h, int n_bindings, const char *const *slots,
                              const int32_t *columns, int n_columns);

/* ---- output selection ------------------------------------------------------ */
/* DefaultModelRunner::output_variables (model_runner.rs:124) plus a time
 * sub-range; default after create = every variable, every time point. */
int rscm_b200_select_outputs(rscm_b200_ensemble *h, int n_vars, const int32_t *vars,
                             int32_t t_start, int32_t t_stop, int32_t t_step);
/* rows of the output matrix = sum over selected vars of n_regions * n_selected_times */
int64_t rscm_b200_output_rows(const rscm_b200_ensemble *h);

/* ---- run: ModelRunner::run_batch (model_runner.rs:261-266) ---------------- */
/* params   : [n_columns][M] (layout 0, SoA) or [M][n_columns] (layout 1, the
 *            reference's &[Vec<f64>])
 * scenarios: [S][n_exogenous][T][R_v]   (R_v regions of that variable)
 * out      : [rows][S*M], run index = s*M + m; row order = selected variable,
 *            then time, then region: Timeseries storage [T][R]
 *            (timeseries.rs:261-275), NaN where the reference leaves NaN
 * status   : [S*M] bit0 = a component failed / get_last_step assertion would fire,
 *            bit1 = a non-finite value was produced; may be NULL
 * All pointers are DEVICE pointers; work is enqueued on `stream`
 * (a cudaStream_t, NULL = default stream) and is asynchronous.
 * Streams: the handle's staged scenario table, scratch and summary buffers are
 * per handle, so the engine orders launches itself: a launch on a stream other
 * than the previous launch's first waits (cudaStreamWaitEvent) for the previous
 * launch's kernel.  Calls on different streams, and device calls followed by the
 * host entry points (which use internal non-blocking streams), are therefore safe
 * but do not overlap each other; the CALLER's buffers (params, scenarios, out)
 * must still not be reused before the work that reads/writes them has finished.
 */
int rscm_b200_run_device(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                         const double *scenarios, int64_t S, double *out, uint8_t *status,
                         void *stream);
/* Same with HOST pointers: chunks members, overlaps H2D / kernel / D2H on two
 * streams, returns when `out` is complete.  Pinned host memory gives full PCIe rate. */
int rscm_b200_run_host(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                       const double *scenarios, int64_t S, double *out, uint8_t *status);

/* ---- calibration: log-posterior per member -------------------------------- */
/* Observation = crates/rscm-calibrate/src/target.rs:25; time already resolved
 * to an index with rscm_b200_time_index. */
typedef struct {
    int32_t variable;   /* model variable index (scalar variables only, model_runner.rs:175-190) */
    int32_t time_index;
    double value;
    double sigma;
} rscm_b200_obs;

typedef enum {
    RSCM_B200_PRIOR_NONE = 0,
    RSCM_B200_PRIOR_UNIFORM = 1,        /* a=low, b=high     distribution.rs:157-163 */
    RSCM_B200_PRIOR_NORMAL = 2,         /* a=mean, b=std     distribution.rs:256-259 */
    RSCM_B200_PRIOR_LOGNORMAL = 3,      /* a=mu, b=sigma     distribution.rs:353-360 */
    RSCM_B200_PRIOR_BOUND_NORMAL = 4,   /* Bound{Normal}     distribution.rs:490-497 */
    RSCM_B200_PRIOR_BOUND_LOGNORMAL = 5,
    RSCM_B200_PRIOR_BOUND_UNIFORM = 6
} rscm_b200_prior_kind;

typedef struct {
    int32_t kind;
    int32_t reserved;
    double a, b;
    double low, high;
} rscm_b200_prior;

/* Target + GaussianLikelihood{normalize} (likelihood.rs:99-253); observations
 * grouped by variable in Target order. */
int rscm_b200_set_target(rscm_b200_ensemble *h, const rscm_b200_obs *obs, int64_t n_obs, int normalize);
/* ParameterSet (parameter_set.rs:255-270): one prior per parameter column, or n=0 for none */
int rscm_b200_set_priors(rscm_b200_ensemble *h, const rscm_b200_prior *priors, int n_columns);

/* ensemble-level summary of a log-posterior evaluation (warp-shuffle + block reduction) */
typedef struct {
    double max_logpost;
    int64_t argmax;    /* run index of the maximum, -1 if none finite */
    double sum_finite; /* sum of finite log-posteriors */
    int64_t n_finite;
    int64_t n_runs;
} rscm_b200_logpost_summary;

/* EnsembleSampler::log_posterior_batch (sampler/ensemble.rs:143-178) fused into
 * the member loop: no timeseries is written, 8 B per run.  Device pointers;
 * `summary` (device, may be NULL) receives the block-reduced summary. */
int rscm_b200_logpost_device(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                             const double *scenarios, int64_t S, double *logpost,
                             rscm_b200_logpost_summary *summary, void *stream);
int rscm_b200_logpost_host(rscm_b200_ensemble *h, const double *params, int64_t M, int params_layout,
                           const double *scenarios, int64_t S, double *logpost,
                           rscm_b200_logpost_summary *summary);

/* ---- measurement helpers --------------------------------------------------- */
/* number of engine kernels launched through this handle since creation */
int64_t rscm_b200_launch_count(const rscm_b200_ensemble *h);
/* average device time (ms, CUDA events on the launch stream) of the main fused
 * kernel over the launches since the last reset; 0 if none */
double rscm_b200_kernel_ms(rscm_b200_ensemble *h, int reset);
/* dynamic shared memory (bytes) one CTA of the fused kernel asks for (log_posterior != 0: the variant that also stages
 * the observation tables): with the kernel's register count this fixes how many CTAs an SM holds */
int64_t rscm_b200_shared_bytes(const rscm_b200_ensemble *h, int log_posterior);
/* DFMA-saturating micro-benchmark: measured FP64 (or FP32 when dtype=1) FMA
 * throughput of `device` in TFLOP/s (2 flop per FMA) */
int rscm_b200_measure_fma_peak(int device, int dtype, double *tflops);

/* ---- scenario ingestion on the device -----------------------------------------
 * Timeseries::interpolate_into (crates/rscm-core/src/timeseries.rs:586-611; the step ModelBuilder::build applies to every
 * exogenous series, model/builder.rs:768) for n_series series that share one source axis: source values
 * [n_series][K][R] at times d_src_times[K] (time_axis.values()), result [n_series][T][R] at d_dst_times[T].
 * strategy 0 = Linear (interpolate/strategies/linear_spline.rs:33-95), 1 = Next, 2 = Previous, with the is_close!
 * boundary snap (strategies/mod.rs:38-41) and extrapolation beyond both ends, as interpolate_into allows.  With
 * n_series = S * n_exogenous this turns scenario files of any resolution into the [S][n_exo][T][R] block
 * rscm_b200_run_device takes, without a host pass.  Device pointers; asynchronous on `stream`. */
int rscm_b200_interpolate_device(const double *d_src_times, int64_t K, const double *d_src_values, int64_t n_series, int R,
                                 const double *d_dst_times, int64_t T, int strategy, double *d_out, void *stream);

/* Self-test hook: y[i] = exp(x[i]) (op 0), log(x[i]) (op 1) or pow(x[2i], x[2i+1]) (op 2; x holds 2n values) with the
 * engine's own fp64 device implementations (rscm_b200/csrc/components.cuh), so that their accuracy can be checked
 * from the host.  n = number of results. */
int rscm_b200_device_math(int op, const double *d_x, int64_t n, double *d_y, void *stream);

/* ---- ensemble sampler: the stretch move on the device ------------------------
 * Replaces the per-walker host loops of EnsembleSampler::update_group
 * (crates/rscm-calibrate/src/sampler/ensemble.rs:489-546) around the
 * log_posterior_batch call: StretchMove::propose / sample_z (moves.rs:55-59,
 * 110-125) and acceptance_probability + accept/reject (moves.rs:76-92,
 * ensemble.rs:520-543).  All pointers are DEVICE pointers on the current device;
 * positions are SoA: (column c, walker w) at positions[c*ld + w]; the active half
 * is [active_begin, active_begin + n_active), the complementary half
 * [comp_begin, comp_begin + n_comp).  Every random draw is a pure function of
 * (seed, walker index, step, purpose) through Philox4x32-10 (the reference uses a
 * non-reproducible thread_rng: statistical parity only), so ranks that hold the
 * same walker state take identical decisions.  `step` must differ between the two
 * half-updates of an iteration (e.g. 2*iteration + half).
 *   propose: proposals[c*ld_proposals + i] = x_j + z_i (x_i - x_j), z[i] = z_i, i in [0, n_active)
 *   accept : walker i takes its proposal (and logpost[w] = logpost_new[i]) with
 *            probability min(1, z^(n_cols-1) exp(new - old)), 0 when `new` is not finite;
 *            *n_accepted (may be NULL) is incremented by the number of accepted moves. */
int rscm_b200_stretch_propose(const double *d_positions, int64_t ld, int n_cols, int64_t active_begin, int64_t n_active,
                              int64_t comp_begin, int64_t n_comp, double a, uint64_t seed, uint32_t step, double *d_proposals,
                              int64_t ld_proposals, double *d_z, void *stream);
int rscm_b200_stretch_accept(double *d_positions, int64_t ld, int n_cols, int64_t active_begin, int64_t n_active,
                             const double *d_proposals, int64_t ld_proposals, const double *d_z, const double *d_logpost_new,
                             double *d_logpost, uint64_t seed, uint32_t step, unsigned long long *d_n_accepted, void *stream);

/* ---- multi-GPU: member sharding and the all-gather of per-member log-posteriors ----
 * The path shards by member: ModelRunner::run_batch maps `run` over independent
 * rows (crates/rscm-calibrate/src/model_runner.rs:261-266, rayon par_iter), and
 * EnsembleSampler::log_posterior_batch (sampler/ensemble.rs:143-178) needs every
 * member's log-posterior on the host that advances the walkers.  One process per
 * GPU; rank r owns members [r*M/G, (r+1)*M/G) (rscm_b200_comm_shard).  The only
 * exchange is 8 bytes per member.  A communicator wraps an NCCL communicator
 * (libnccl.so.2 is loaded at run time: RSCM_B200_NCCL_LIB, else the copy already
 * mapped into the process, else the system's) plus, when the GPUs can map each
 * other's memory (CUDA IPC over NVLink / NVSwitch), a table of peer pointers for
 * buffers obtained from rscm_b200_comm_symmetric_alloc.
 *
 * Set-up mirrors ncclGetUniqueId / ncclCommInitRank: rank 0 calls
 * rscm_b200_comm_unique_id, the host distributes the 128 bytes by its own means
 * (MPI, a file, torch.distributed's store), every rank calls rscm_b200_comm_init.
 * All of comm_init, comm_symmetric_alloc, allgather_f64, logpost_sharded_device and
 * sampler_iterate are COLLECTIVE: every rank must call them in the same order.
 */
#define RSCM_B200_UNIQUE_ID_BYTES 128
typedef struct rscm_b200_comm rscm_b200_comm;

int rscm_b200_comm_unique_id(void *unique_id /* RSCM_B200_UNIQUE_ID_BYTES */);
/* device: CUDA ordinal of this rank's GPU (-1 = current).  world = 1 needs no NCCL and no unique id. */
int rscm_b200_comm_init(const void *unique_id, int rank, int world, int device, rscm_b200_comm **out);
void rscm_b200_comm_destroy(rscm_b200_comm *c);
const char *rscm_b200_comm_last_error(const rscm_b200_comm *c);
int rscm_b200_comm_rank(const rscm_b200_comm *c);
int rscm_b200_comm_world(const rscm_b200_comm *c);
/* 1 when symmetric buffers are mapped into every peer (the fused all-gather is available), else 0 (NCCL only;
 * RSCM_B200_NO_P2P=1 forces this) */
int rscm_b200_comm_peer_access(const rscm_b200_comm *c);
/* member block of this rank: [begin, end) = [rank*M/world, (rank+1)*M/world) */
int rscm_b200_comm_shard(const rscm_b200_comm *c, int64_t M, int64_t *begin, int64_t *end);
/* `bytes` of zeroed device memory that every peer can store into (freed with the communicator) */
int rscm_b200_comm_symmetric_alloc(rscm_b200_comm *c, size_t bytes, void **d_ptr);
/* ncclAllGather of n_local doubles per rank: d_global[r*n_local + i] = rank r's d_local[i].  Device pointers;
 * in place when d_local == d_global + rank*n_local.  Replaces the rayon join at the end of run_batch. */
int rscm_b200_allgather_f64(rscm_b200_comm *c, const double *d_local, int64_t n_local, double *d_global, void *stream);
/* log_posterior_batch over G GPUs.  `params` is the GLOBAL matrix ([n_columns][M] layout 0 or [M][n_columns]
 * layout 1), identical on every rank; each rank evaluates its member block in place (no copy of the block: the kernel
 * takes the offset and the global leading dimension) and all ranks end up with all S*M log-posteriors in
 * logpost_global[s*M + m].  When logpost_global lies in symmetric memory and peers are mapped, ONE kernel does both:
 * it stores each member's 8 bytes into every peer's buffer over NVLink as the member finishes and its last block raises
 * a flag in every peer; a one-warp kernel then waits for all flags.  Otherwise the kernel is followed by ncclAllGather
 * (equal blocks, S = 1) or grouped ncclBroadcasts, in place.  Asynchronous on `stream`.
 * With the fused path consecutive calls must alternate between two destination buffers (a peer may still be reading
 * the previous result while this rank's next evaluation stores into it). */
int rscm_b200_logpost_sharded_device(rscm_b200_ensemble *h, rscm_b200_comm *c, const double *params, int64_t M, int params_layout,
                                     const double *scenarios, int64_t S, double *logpost_global, void *stream);
/* synchronous health check of the fused path: RSCM_B200_ECOMM if a peer missed a rendezvous (20 s timeout on the device) */
int rscm_b200_comm_check(rscm_b200_comm *c);

/* Walker state of the stretch-move sampler, all DEVICE pointers, replicated on every rank
 * (EnsembleSampler::run, sampler/ensemble.rs:412-487; SamplerState, sampler/state.rs). */
typedef struct {
    double *positions;       /* [n_cols][ld] SoA */
    int64_t ld;
    int32_t n_cols;
    int32_t reserved;
    int64_t n_walkers;       /* even */
    double *logpost;         /* [n_walkers] current log-posterior of every walker */
    double *proposals;       /* [n_cols][n_walkers/2] scratch */
    double *z;               /* [n_walkers/2] scratch */
    double *logpost_new[2];  /* [n_walkers/2] each, one per half-update; symmetric memory enables the fused all-gather */
    unsigned long long *n_accepted; /* running count of accepted moves, may be NULL */
    double a;                /* stretch parameter (> 1) */
    uint64_t seed;
    uint32_t first_iteration; /* iteration number of the first iteration of this call (keys the random stream) */
    uint32_t thin;           /* chain recording: after iteration i with i % thin == 0 ... (0 = no recording) */
    double *chain_positions; /* ... positions are copied to chain_positions[(i/thin)][n_cols][n_walkers] */
    double *chain_logpost;   /* ... and logpost to chain_logpost[(i/thin)][n_walkers] (Chain::push, sampler/chain.rs:63) */
    int64_t chain_capacity;  /* kept samples the two chain buffers can hold */
} rscm_b200_sampler_state;
/* n_iterations of EnsembleSampler::update (sampler/ensemble.rs:489-546): per half-ensemble propose -> sharded fused
 * log-posterior (+ all-gather) -> accept/reject, enqueued on `stream` without host synchronisation.  The first iteration
 * of a call runs eagerly; with use_graph != 0 the remaining ones replay ONE captured CUDA graph (the iteration number is a
 * device-side counter), so the host cost per iteration is a single graph launch.  Target and priors must be set. */
int rscm_b200_sampler_iterate(rscm_b200_ensemble *h, rscm_b200_comm *c, const rscm_b200_sampler_state *state, const double *scenarios,
                              int64_t S, int n_iterations, int use_graph, void *stream);

/* ---- ensemble summaries on the device -----------------------------------------
 * Quantiles across members of an output block d_out[rows][S*M] (the layout of
 * rscm_b200_run_device), per row and scenario: d_result[nq][rows][S].  This is the
 * across-member percentile band users of the reference compute with pandas /
 * numpy over looped Model::run results (docs/notebooks/scenario_pipeline.py:
 * 339-400; python/rscm/calibrate/pandas_helpers.py); on the device it replaces the
 * copy of the whole block to the host.  NaNs are ignored and interpolation is
 * numpy's "linear" method: results equal numpy.nanquantile bit for bit.  `q` is a
 * HOST array of 1..5 values in [0, 1]; d_out / d_result are device pointers. */
int rscm_b200_member_quantiles(const double *d_out, int64_t rows, int64_t S, int64_t M, const double *q, int nq, double *d_result,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RSCM_B200_H */
