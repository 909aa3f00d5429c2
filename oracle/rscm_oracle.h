/*
 * rscm_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, fp64, scalar restatement of the reference algorithm for the
 * ensemble hot path of lewisjared/rscm v0.5.0 (paths below are relative to
 * the reference checkout).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped CUDA path never calls into it.
 *
 * What is restated (each function in rscm_oracle.c cites file:line):
 *   - framework semantics: ModelBuilder::build variable-source
 *     classification + graph edges (crates/rscm-core/src/model/builder.rs:418-560,
 *     631-700), BFS execution order (model/runtime.rs:504-510), per-step
 *     read/write indexing (model/runtime.rs:368-497, state/windows.rs:155-234),
 *     NaN-initialised storage (model/builder.rs:735-830), schema aggregates
 *     (schema.rs:760-806, 874-951);
 *   - physics: TwoLayer (crates/rscm-two-layer/src/component.rs:160-251),
 *     CarbonCycle (crates/rscm-components/src/components/carbon_cycle.rs:102-158),
 *     CO2ERF (co2_erf.rs:57-81), MAGICC box components (magicc_*.c);
 *   - numerics: classical fixed-step RK4 as published by the third-party crate
 *     ode_solvers 0.6.1 (Cargo.lock:623-626; source NOT in the reference tree —
 *     restated from its published algorithm: n = ceil((x_end-x)/h) steps of
 *     constant h, no end clipping), get_last_step check
 *     (crates/rscm-core/src/ivp/mod.rs:73-102);
 *   - calibration: Gaussian likelihood (crates/rscm-calibrate/src/likelihood.rs:186-253),
 *     priors (distribution.rs:157-163,256-259,353-360,490-497),
 *     log-posterior assembly (sampler/ensemble.rs:143-178).
 *
 * PARITY PINNING STATUS:
 *   - The Rust reference cannot be built or imported here (no cargo/rustc, no
 *     wheel), so no reference-generated outputs exist.
 *   - CarbonCycle: pinned by the reference's own analytical test
 *     (crates/rscm-components/tests/coupled_models.rs:13-141, rel < 1e-2).
 *   - CO2ERF: pinned by co2_erf.rs:94-113 known answers (1e-10).
 *   - GhgForcing: pinned by the MAGICC7 golden CSVs the reference tests use
 *     (tests/regression/data/ghg_forcing/NN.csv, rtol 1e-5 / atol 1e-6).
 *   - Gaussian likelihood / priors / aggregate: pinned by the reference's
 *     doctest and unit-test known answers.
 *   - ClimateUDEB: pinned by the MAGICC7 ocean golden runs of the reference's
 *     regression suite (tests/regression/data/ocean_udeb, phased 1-5 %
 *     tolerances of tests/regression/test_ocean_udeb.py).
 *   - The other MAGICC boxes (ozone, aerosols, CH4 / N2O / halocarbon chemistry,
 *     terrestrial / ocean carbon, CO2 budget, LAMCALC): formulas pinned by the
 *     known answers of the reference's unit tests (tests/test_magicc_*.py,
 *     test_ocean_carbon.py, test_halocarbon.py); their multi-decade series:
 *     PARITY UNPINNED (no reference-generated numbers).
 *   - TwoLayer numeric values: PARITY UNPINNED — the reference's tests for it
 *     are qualitative only (component.rs:300-406).
 */
#ifndef RSCM_ORACLE_H
#define RSCM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* component kinds (oracle numbering is private to the oracle) */
enum {
    ORC_TWO_LAYER = 1,     /* params: lambda0, a, efficacy, eta, heat_capacity_surface, heat_capacity_deep */
    ORC_CARBON_CYCLE = 2,  /* params: tau, conc_pi, alpha_temperature, step_size */
    ORC_CO2_ERF = 3,       /* params: erf_2xco2, conc_pi */
    ORC_AGGREGATOR = 4,    /* internal */
    ORC_GHG_FORCING = 5,   /* params: see magicc_forcing.c */
    ORC_OZONE_FORCING = 6,
    ORC_AEROSOL_DIRECT = 7,
    ORC_AEROSOL_INDIRECT = 8,
    ORC_CLIMATE_UDEB = 9,  /* params: see magicc_climate.c */
    ORC_FOUR_BOX_OHU = 10, /* the following: see magicc_boxes.c */
    ORC_OCEAN_SURFACE_PP = 11,
    ORC_CO2_BUDGET = 12,
    ORC_TERRESTRIAL_CARBON = 13,
    ORC_CH4_CHEMISTRY = 14,
    ORC_N2O_CHEMISTRY = 15,
    ORC_OCEAN_CARBON = 16, /* see magicc_ocean.c */
    ORC_HALOCARBON_CHEMISTRY = 17, /* see magicc_halocarbon.c */
    ORC_KIND_MAX = 32
};

enum { ORC_GRID_SCALAR = 0, ORC_GRID_FOUR_BOX = 1, ORC_GRID_HEMISPHERIC = 2 };
enum { ORC_AGG_SUM = 0, ORC_AGG_MEAN = 1, ORC_AGG_WEIGHTED = 2 };
enum { ORC_SRC_EXOGENOUS = 0, ORC_SRC_OWN_STATE = 1, ORC_SRC_UPSTREAM = 2 };
enum { ORC_REQ_INPUT = 0, ORC_REQ_OUTPUT = 1, ORC_REQ_STATE = 2 };

typedef struct orc_model orc_model;

orc_model *orc_model_new(void);
void orc_model_free(orc_model *m);
const char *orc_last_error(const orc_model *m);

/* builder (call order of orc_add_component is the reference's insertion order) */
int orc_add_component(orc_model *m, int kind, const double *params, int n_params);
int orc_add_schema_variable(orc_model *m, const char *name, int grid);
int orc_add_aggregate(orc_model *m, const char *name, int op, int grid, int n_contrib,
                      const char *const *contributors, const double *weights);
int orc_set_initial_value(orc_model *m, const char *name, double v);
int orc_set_time_bounds(orc_model *m, const double *bounds, int n_times); /* bounds[n_times+1] */
int orc_set_exogenous(orc_model *m, const char *name, int grid, const double *values); /* [T][R] */
int orc_set_unit_factor(orc_model *m, int component, const char *variable, double factor);
int orc_set_grid_weights(orc_model *m, int grid, const double *weights);
int orc_build(orc_model *m);

/* introspection after build */
int orc_n_variables(const orc_model *m);
const char *orc_variable_name(const orc_model *m, int v);
int orc_variable_grid(const orc_model *m, int v);
int orc_variable_index(const orc_model *m, const char *name);
int orc_variable_is_endogenous(const orc_model *m, int v);
int orc_n_nodes(const orc_model *m);                  /* components + aggregators (no Null root) */
int orc_execution_order(const orc_model *m, int *order); /* node ids in BFS order; returns count */
int orc_node_kind(const orc_model *m, int node);
int orc_variable_source(const orc_model *m, int component, const char *variable);
int orc_n_times(const orc_model *m);

/* single run with the components' own parameters; out = [V][T][R_v] packed by
 * orc_variable_offset (units of doubles, T*R_v each) */
int64_t orc_variable_offset(const orc_model *m, int v);
int64_t orc_output_size(const orc_model *m);
int orc_run(orc_model *m, double *out);

/* ensemble: M parameter rows x S scenarios.
 *  bind_component[j] >= 0 : column j overrides params[bind_index[j]] of that component
 *  bind_component[j] == -1: column j overrides the initial value of variable bind_index[j]
 *  params: [M][n_cols] row-major (the reference's &[Vec<f64>]).
 *  scenarios: [S][n_exo][T*R] in the order of exo_vars; may be NULL when S==0
 *             (then the builder's exogenous data is used, S treated as 1).
 *  out: [n_out][T*R][S*M] (run index = s*M + m, fastest), status: [S*M]
 */
int orc_run_batch(const orc_model *m, int n_cols, const int *bind_component, const int *bind_index,
                  const double *params, int64_t M, int n_exo, const int *exo_vars,
                  const double *scenarios, int64_t S, int n_out, const int *out_vars, double *out,
                  uint8_t *status, int n_threads);

/* calibration */
typedef struct {
    int32_t variable;   /* index into out_vars-independent model variable table */
    int32_t time_index; /* resolved on the host with the "{:.6}" key rule */
    double value;
    double sigma;
} orc_obs;

enum { ORC_PRIOR_NONE = 0, ORC_PRIOR_UNIFORM = 1, ORC_PRIOR_NORMAL = 2, ORC_PRIOR_LOGNORMAL = 3,
       ORC_PRIOR_BOUND_NORMAL = 4, ORC_PRIOR_BOUND_LOGNORMAL = 5, ORC_PRIOR_BOUND_UNIFORM = 6 };
typedef struct {
    int32_t kind;
    int32_t pad;
    double a, b;     /* uniform: low, high; normal: mean, std; lognormal: mu, sigma */
    double low, high; /* Bound wrapper limits */
} orc_prior;

double orc_ln_pdf(const orc_prior *p, double x);
double orc_log_prior(const orc_prior *priors, int n, const double *x);
/* series: full single-run output as produced by orc_run */
double orc_ln_likelihood(const orc_model *m, const double *run_out, const orc_obs *obs, int64_t K,
                         int normalize);
int orc_time_index(const orc_model *m, double time); /* -1 when no key matches */
int orc_log_posterior_batch(const orc_model *m, int n_cols, const int *bind_component,
                            const int *bind_index, const double *params, int64_t M, int n_exo,
                            const int *exo_vars, const double *scenarios, int64_t S,
                            const orc_prior *priors, const orc_obs *obs, int64_t K, int normalize,
                            double *logpost, int n_threads);

/* small pure functions exposed for known-answer tests */
double orc_compute_aggregate(const double *vals, const double *weights, int n, int op);
double orc_co2_erf(double erf_2xco2, double conc_pi, double conc);
int orc_rk4_steps(double t0, double t1, double h);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
