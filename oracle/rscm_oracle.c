/*
 * rscm_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See rscm_oracle.h for scope and the parity-pinning statement.
 *
 * Compiled without fast-math and with -ffp-contract=off so that the operation
 * order written here (which follows the Rust reference expression by
 * expression) is the order executed: rustc never contracts a*b+c into an FMA.
 */
#include "orc_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* numerics: ode_solvers 0.6.1 Rk4 (third-party; restated from its published  */
/* algorithm) + rscm-core ivp::get_last_step                                  */
/* ------------------------------------------------------------------------- */

/* number of fixed steps: Rk4::integrate computes
 *   num_steps = ((x_end - x) / step_size).ceil() as usize
 * (call sites: crates/rscm-core/src/ivp/mod.rs:245-253). */
int orc_rk4_steps(double t0, double t1, double h)
{
    double n = ceil((t1 - t0) / h);
    if (!(n >= 0.0)) return 0;
    return (int)n;
}

/* Classical RK4, constant h for all n steps (no clipping of the last step):
 *   k0 = f(x, y); k1 = f(x+h/2, y + k0*(h/2)); k2 = f(x+h/2, y + k1*(h/2));
 *   k3 = f(x+h, y + k2*h);  y += (k0 + k1*2 + k2*2 + k3) * (h/6);  x += h
 * followed by get_last_step's assertion |x_last - t_next| < 5e-3
 * (crates/rscm-core/src/ivp/mod.rs:73,90-102).  All right-hand sides on this
 * path are autonomous, x only feeds that assertion. */
int orc_rk4(orc_rhs f, const void *self, int dim, double t0, double t1, double h, double *y)
{
    double k0[8], k1[8], k2[8], k3[8], tmp[8];
    const double half = h / 2.0;
    const double h6 = h / 6.0;
    const int n = orc_rk4_steps(t0, t1, h);
    double x = t0;
    for (int s = 0; s < n; ++s) {
        f(self, y, k0);
        for (int i = 0; i < dim; ++i) tmp[i] = y[i] + k0[i] * half;
        f(self, tmp, k1);
        for (int i = 0; i < dim; ++i) tmp[i] = y[i] + k1[i] * half;
        f(self, tmp, k2);
        for (int i = 0; i < dim; ++i) tmp[i] = y[i] + k2[i] * h;
        f(self, tmp, k3);
        for (int i = 0; i < dim; ++i)
            y[i] = y[i] + (((k0[i] + k1[i] * 2.0) + k2[i] * 2.0) + k3[i]) * h6;
        x = x + h;
    }
    /* results always holds the initial point, so y.len() > 1 iff n >= 1 */
    if (n < 1) return 1;
    if (!(fabs(x - t1) < 5e-3)) return 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* window accessors                                                           */
/* ------------------------------------------------------------------------- */

static inline const double *var_block(const orc_ctx *c, int v)
{
    return c->data + c->m->vars[v].offset;
}

/* Raw read of input i at absolute time index, with the read-side grid
 * transform (model/runtime.rs:399-408 -> state/aggregating.rs:162-177,611-618)
 * and the unit factor (state/windows.rs:158-159) applied. */
double orc_in_at(const orc_ctx *c, int i, int region, int index)
{
    const orc_model *m = c->m;
    const int v = c->node->in_var[i];
    const orc_var *var = &m->vars[v];
    if (index < 0 || index >= m->T) return NAN;
    const double *row = var_block(c, v) + (int64_t)index * var->n_regions;
    const int want = c->node->in_grid[i];
    const double k = c->node->in_factor[i];
    if (want == var->grid) return row[region] * k;
    if (var->grid == ORC_GRID_FOUR_BOX && want == ORC_GRID_SCALAR) {
        /* AggregatingFourBoxWindow::aggregate: with custom weights -> sum of
         * w_i v_i over non-NaN terms (no renormalisation); default grid ->
         * FourBoxGrid::aggregate_global (spatial/four_box.rs:146) */
        const double *w = m->w_fourbox;
        double s = 0.0;
        for (int r = 0; r < 4; ++r) {
            /* custom weights (with_grid_weights): NaN terms are skipped; default grid:
             * aggregate_global is a plain sum of v*w, NaN propagates */
            if (m->has_w_fourbox && isnan(row[r])) continue;
            s += row[r] * w[r];
        }
        return s * k;
    }
    if (var->grid == ORC_GRID_FOUR_BOX && want == ORC_GRID_HEMISPHERIC) {
        /* state/aggregating.rs:611-618 */
        double a = (region == 0) ? (row[0] + row[1]) / 2.0 : (row[2] + row[3]) / 2.0;
        return a * k;
    }
    if (var->grid == ORC_GRID_HEMISPHERIC && want == ORC_GRID_SCALAR) {
        const double *w = m->w_hemi;
        double s = 0.0;
        for (int r = 0; r < 2; ++r) {
            if (m->has_w_hemi && isnan(row[r])) continue;
            s += row[r] * w[r];
        }
        return s * k;
    }
    return NAN; /* broadcast (coarse->fine) is rejected at build */
}

double orc_in_start(const orc_ctx *c, int i, int region) { return orc_in_at(c, i, region, c->N); }

double orc_in_end(const orc_ctx *c, int i, int region, int *ok)
{
    const int next = c->N + 1;
    if (next >= c->m->T) { if (ok) *ok = 0; return NAN; }
    if (ok) *ok = 1;
    return orc_in_at(c, i, region, next);
}

/* TimeseriesWindow::get — state/windows.rs:229-234 */
double orc_in_get(const orc_ctx *c, int i, int region)
{
    if (c->node->in_src[i] == ORC_SRC_UPSTREAM) {
        int ok;
        double v = orc_in_end(c, i, region, &ok);
        if (ok) return v;
    }
    return orc_in_start(c, i, region);
}

double orc_in_offset(const orc_ctx *c, int i, int region, int off, int *ok)
{
    const int idx = c->N + off;
    if (idx < 0 || idx >= c->m->T) { if (ok) *ok = 0; return NAN; }
    if (ok) *ok = 1;
    return orc_in_at(c, i, region, idx);
}

/* latest non-NaN value at or before N+1 (Timeseries::latest_value semantics as
 * used by InputState::get_global for endogenous series, state/mod.rs:231-254) */
double orc_in_latest(const orc_ctx *c, int i, int region)
{
    int hi = c->N + 1;
    if (hi >= c->m->T) hi = c->m->T - 1;
    for (int idx = hi; idx >= 0; --idx) {
        double v = orc_in_at(c, i, region, idx);
        if (!isnan(v)) return v;
    }
    return NAN;
}

/* ------------------------------------------------------------------------- */
/* TwoLayer — crates/rscm-two-layer/src/component.rs                          */
/* ------------------------------------------------------------------------- */

typedef struct {
    double lambda0, a, efficacy, eta, cs, cd;
    double erf;
} two_layer_sys;

/* IVP::calculate_dy_dt — component.rs:160-188 (expression order preserved) */
static void two_layer_rhs(const void *self, const double *y, double *dy)
{
    const two_layer_sys *p = (const two_layer_sys *)self;
    const double ts = y[0];
    const double td = y[1];
    const double erf = p->erf;
    const double diff = ts - td;
    const double lambda_eff = p->lambda0 - p->a * ts;
    const double hx_surface = p->efficacy * p->eta * diff;
    const double dts = (erf - lambda_eff * ts - hx_surface) / p->cs;
    const double hx_deep = p->eta * diff;
    const double dtd = hx_deep / p->cd;
    dy[0] = dts;
    dy[1] = dtd;
    dy[2] = p->cs * dts + p->cd * dtd;
}

/* Component::solve — component.rs:223-251.  inputs: [erf, Ts(state), Td(state)];
 * outputs: [Ts, Td]. h = 0.1 hard-coded (:240). */
static int two_layer_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    two_layer_sys s = {p[0], p[1], p[2], p[3], p[4], p[5], 0.0};
    s.erf = orc_in_get(c, 0, 0);
    double y[3] = {orc_in_start(c, 1, 0), orc_in_start(c, 2, 0), 0.0};
    if (orc_rk4(two_layer_rhs, &s, 3, t0, t1, 0.1, y)) return 1;
    out[0] = y[0];
    out[1] = y[1];
    return 0;
}

static const orc_def two_layer_defs[] = {
    {"Effective Radiative Forcing", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Deep Ocean Temperature", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
static const orc_kind_info kind_two_layer = {ORC_TWO_LAYER, "TwoLayer", 3, two_layer_defs, 6,
                                             two_layer_solve, 0, NULL};

/* ------------------------------------------------------------------------- */
/* CarbonCycle — crates/rscm-components/src/components/carbon_cycle.rs        */
/* ------------------------------------------------------------------------- */

#define ORC_GTC_PER_PPM 2.13 /* crates/rscm-components/src/constants.rs:37 */

typedef struct {
    double tau, conc_pi, alpha;
    double emissions, temperature;
} carbon_sys;

/* IVP::calculate_dy_dt — carbon_cycle.rs:134-158 */
static void carbon_rhs(const void *self, const double *y, double *dy)
{
    const carbon_sys *p = (const carbon_sys *)self;
    const double conc = y[0];
    const double lifetime = p->tau * exp(p->alpha * p->temperature);
    const double uptake = (conc - p->conc_pi) / lifetime;
    dy[0] = p->emissions / ORC_GTC_PER_PPM - uptake;
    dy[1] = uptake * ORC_GTC_PER_PPM;
    dy[2] = p->emissions;
}

/* Component::solve — carbon_cycle.rs:102-131.
 * inputs (definition order: inputs then states): [emissions, temperature,
 * concentration, cumulative_emissions, cumulative_uptake]; y0 order is
 * (concentration, cumulative_uptake, cumulative_emissions) (:110-114);
 * outputs (states order): [concentration, cumulative_emissions, cumulative_uptake]. */
static int carbon_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    carbon_sys s = {p[0], p[1], p[2], 0.0, 0.0};
    s.emissions = orc_in_get(c, 0, 0);
    s.temperature = orc_in_get(c, 1, 0);
    double y[3] = {orc_in_start(c, 2, 0), orc_in_start(c, 4, 0), orc_in_start(c, 3, 0)};
    if (orc_rk4(carbon_rhs, &s, 3, t0, t1, p[3], y)) return 1;
    out[0] = y[0];
    out[1] = y[2];
    out[2] = y[1];
    return 0;
}

static const orc_def carbon_defs[] = {
    {"Emissions|CO2|Anthropogenic", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Surface Temperature", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Atmospheric Concentration|CO2", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Cumulative Emissions|CO2", ORC_REQ_STATE, ORC_GRID_SCALAR},
    {"Cumulative Land Uptake", ORC_REQ_STATE, ORC_GRID_SCALAR},
};
static const orc_kind_info kind_carbon = {ORC_CARBON_CYCLE, "CarbonCycle", 5, carbon_defs, 4,
                                          carbon_solve, 0, NULL};

/* ------------------------------------------------------------------------- */
/* CO2ERF — crates/rscm-components/src/components/co2_erf.rs:57-81            */
/* ------------------------------------------------------------------------- */

double orc_co2_erf(double erf_2xco2, double conc_pi, double conc)
{
    return erf_2xco2 / log(2.0) * log(1.0 + (conc - conc_pi) / conc_pi);
}

static int co2_erf_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st; (void)t0; (void)t1;
    out[0] = orc_co2_erf(p[0], p[1], orc_in_get(c, 0, 0));
    return 0;
}

static const orc_def co2_erf_defs[] = {
    {"Atmospheric Concentration|CO2", ORC_REQ_INPUT, ORC_GRID_SCALAR},
    {"Effective Radiative Forcing|CO2", ORC_REQ_OUTPUT, ORC_GRID_SCALAR},
};
static const orc_kind_info kind_co2_erf = {ORC_CO2_ERF, "CO2ERF", 2, co2_erf_defs, 2,
                                           co2_erf_solve, 0, NULL};

/* ------------------------------------------------------------------------- */
/* Aggregates — crates/rscm-core/src/schema.rs:760-806, 874-951               */
/* ------------------------------------------------------------------------- */

double orc_compute_aggregate(const double *vals, const double *weights, int n, int op)
{
    double sum = 0.0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (isnan(vals[i])) continue;
        sum += (op == ORC_AGG_WEIGHTED) ? vals[i] * weights[i] : vals[i];
        ++cnt;
    }
    if (cnt == 0) return NAN;
    if (op == ORC_AGG_MEAN) return sum / (double)cnt;
    return sum;
}

static const orc_kind_info kind_aggregator = {ORC_AGGREGATOR, "AggregatorComponent", 0, NULL, 0,
                                              NULL, 0, NULL};

static void aggregator_solve(const orc_node *n, orc_ctx *c, double *out)
{
    const int R = orc_grid_regions(n->agg_grid);
    double vals[ORC_MAX_CONTRIB];
    for (int r = 0; r < R; ++r) {
        for (int i = 0; i < n->agg_n; ++i) {
            int ok;
            double v = orc_in_end(c, i, r, &ok); /* at_end().unwrap_or(at_start()) :887-896 */
            if (!ok) v = orc_in_start(c, i, r);
            vals[i] = v;
        }
        out[r] = orc_compute_aggregate(vals, n->agg_w, n->agg_n, n->agg_op);
    }
}

/* ------------------------------------------------------------------------- */
/* kind registry                                                              */
/* ------------------------------------------------------------------------- */

const orc_kind_info *orc_kind_lookup(int kind)
{
    switch (kind) {
    case ORC_TWO_LAYER: return &kind_two_layer;
    case ORC_CARBON_CYCLE: return &kind_carbon;
    case ORC_CO2_ERF: return &kind_co2_erf;
    case ORC_AGGREGATOR: return &kind_aggregator;
    case ORC_GHG_FORCING: return &orc_kind_ghg_forcing;
    case ORC_OZONE_FORCING: return &orc_kind_ozone_forcing;
    case ORC_AEROSOL_DIRECT: return &orc_kind_aerosol_direct;
    case ORC_AEROSOL_INDIRECT: return &orc_kind_aerosol_indirect;
    case ORC_CLIMATE_UDEB: return &orc_kind_climate_udeb;
    case ORC_FOUR_BOX_OHU: return &orc_kind_fbohu;
    case ORC_OCEAN_SURFACE_PP: return &orc_kind_ospp;
    case ORC_CO2_BUDGET: return &orc_kind_co2_budget;
    case ORC_TERRESTRIAL_CARBON: return &orc_kind_terrestrial;
    case ORC_CH4_CHEMISTRY: return &orc_kind_ch4;
    case ORC_N2O_CHEMISTRY: return &orc_kind_n2o;
    case ORC_OCEAN_CARBON: return &orc_kind_ocean_carbon;
    case ORC_HALOCARBON_CHEMISTRY: return orc_kind_halocarbon();
    default: return NULL;
    }
}

/* ------------------------------------------------------------------------- */
/* builder                                                                    */
/* ------------------------------------------------------------------------- */

static int fail(orc_model *m, const char *msg)
{
    snprintf(m->err, sizeof m->err, "%s", msg);
    return -1;
}

orc_model *orc_model_new(void)
{
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    for (int i = 0; i < 4; ++i) m->w_fourbox[i] = 0.25; /* spatial/four_box.rs:70-73 */
    m->w_hemi[0] = m->w_hemi[1] = 0.5;                  /* spatial/hemispheric.rs */
    return m;
}

void orc_model_free(orc_model *m)
{
    if (!m) return;
    for (int i = 0; i < m->n_exo_in; ++i) free(m->exo_vals[i]);
    for (int i = 0; i < m->n_vars; ++i) free(m->vars[i].exo_data);
    free(m->bounds);
    free(m);
}

const char *orc_last_error(const orc_model *m) { return m->err; }

int orc_add_component(orc_model *m, int kind, const double *params, int n_params)
{
    const orc_kind_info *k = orc_kind_lookup(kind);
    if (!k || kind == ORC_AGGREGATOR) return fail(m, "unknown component kind");
    if (n_params != k->n_params) return fail(m, "wrong parameter count");
    if (m->n_user != m->n_nodes) return fail(m, "components must be added before aggregates");
    orc_node *n = &m->nodes[m->n_nodes++];
    memset(n, 0, sizeof *n);
    n->kind = kind;
    n->n_params = n_params;
    memcpy(n->params, params, sizeof(double) * (size_t)n_params);
    m->n_user++;
    return m->n_nodes - 1;
}

int orc_add_schema_variable(orc_model *m, const char *name, int grid)
{
    m->has_schema = 1;
    snprintf(m->schema_name[m->n_schema], ORC_MAX_NAME, "%s", name);
    m->schema_grid[m->n_schema++] = grid;
    return 0;
}

int orc_add_aggregate(orc_model *m, const char *name, int op, int grid, int n_contrib,
                      const char *const *contributors, const double *weights)
{
    m->has_schema = 1;
    orc_node *n = &m->nodes[m->n_nodes++];
    memset(n, 0, sizeof *n);
    n->kind = ORC_AGGREGATOR;
    n->agg_op = op;
    n->agg_grid = grid;
    n->agg_n = n_contrib;
    snprintf(n->agg_name, ORC_MAX_NAME, "%s", name);
    for (int i = 0; i < n_contrib; ++i) {
        snprintf(n->agg_contrib[i], ORC_MAX_NAME, "%s", contributors[i]);
        n->agg_w[i] = weights ? weights[i] : 1.0;
    }
    return m->n_nodes - 1;
}

int orc_set_initial_value(orc_model *m, const char *name, double v)
{
    for (int i = 0; i < m->n_init; ++i)
        if (!strcmp(m->init_name[i], name)) { m->init_val[i] = v; return 0; }
    snprintf(m->init_name[m->n_init], ORC_MAX_NAME, "%s", name);
    m->init_val[m->n_init++] = v;
    return 0;
}

int orc_set_time_bounds(orc_model *m, const double *bounds, int n_times)
{
    free(m->bounds);
    m->T = n_times;
    m->bounds = (double *)malloc(sizeof(double) * (size_t)(n_times + 1));
    memcpy(m->bounds, bounds, sizeof(double) * (size_t)(n_times + 1));
    return 0;
}

int orc_set_exogenous(orc_model *m, const char *name, int grid, const double *values)
{
    if (m->T <= 0) return fail(m, "set the time axis before exogenous data");
    const size_t n = (size_t)m->T * (size_t)orc_grid_regions(grid);
    int i = m->n_exo_in++;
    snprintf(m->exo_name[i], ORC_MAX_NAME, "%s", name);
    m->exo_grid[i] = grid;
    m->exo_vals[i] = (double *)malloc(n * sizeof(double));
    memcpy(m->exo_vals[i], values, n * sizeof(double));
    return 0;
}

int orc_set_unit_factor(orc_model *m, int component, const char *variable, double factor)
{
    int i = m->n_uf++;
    m->uf_comp[i] = component;
    snprintf(m->uf_var[i], ORC_MAX_NAME, "%s", variable);
    m->uf_val[i] = factor;
    return 0;
}

int orc_set_grid_weights(orc_model *m, int grid, const double *w)
{
    if (grid == ORC_GRID_FOUR_BOX) { memcpy(m->w_fourbox, w, 4 * sizeof(double)); m->has_w_fourbox = 1; }
    else if (grid == ORC_GRID_HEMISPHERIC) { memcpy(m->w_hemi, w, 2 * sizeof(double)); m->has_w_hemi = 1; }
    else return fail(m, "weights only apply to FourBox/Hemispheric");
    return 0;
}

static int find_var(const orc_model *m, const char *name)
{
    for (int i = 0; i < m->n_vars; ++i)
        if (!strcmp(m->vars[i].name, name)) return i;
    return -1;
}

static int add_var(orc_model *m, const char *name, int grid, int req)
{
    int v = find_var(m, name);
    if (v >= 0) return v; /* first definition wins: model/validation.rs:30-107 */
    v = m->n_vars++;
    orc_var *var = &m->vars[v];
    memset(var, 0, sizeof *var);
    snprintf(var->name, ORC_MAX_NAME, "%s", name);
    var->grid = grid;
    var->req = req;
    return v;
}

static int is_aggregate_name(const orc_model *m, const char *name)
{
    for (int i = m->n_user; i < m->n_nodes; ++i)
        if (!strcmp(m->nodes[i].agg_name, name)) return 1;
    return 0;
}

static void add_edge(orc_model *m, int from, int to)
{
    m->e_from[m->n_edges] = from;
    m->e_to[m->n_edges] = to;
    m->n_edges++;
}

/* ModelBuilder::build — crates/rscm-core/src/model/builder.rs:418-860.
 * Graph node ids: 0 = NullComponent root, user component i -> i+1,
 * aggregators after them in schema order. */
int orc_build(orc_model *m)
{
    if (m->T < 2 || !m->bounds) return fail(m, "time axis required");
    int producer[ORC_MAX_VARS]; /* `endogenous` map: var -> graph node id */
    for (int i = 0; i < ORC_MAX_VARS; ++i) producer[i] = -1;
    m->n_vars = 0;
    m->n_edges = 0;
    /* pending aggregate deps (builder.rs:441,504-511) */
    int pend_node[ORC_MAX_EDGES], pend_var[ORC_MAX_EDGES], n_pend = 0;

    for (int ci = 0; ci < m->n_user; ++ci) {
        orc_node *n = &m->nodes[ci];
        const orc_kind_info *k = orc_kind_lookup(n->kind);
        const int gnode = ci + 1;
        int has_dep = 0;
        n->n_in = n->n_out = 0;
        /* classification pass (builder.rs:465-488) + edges (:490-519), in inputs() order */
        for (int d = 0; d < k->n_defs; ++d) {
            const orc_def *def = &k->defs[d];
            if (def->req == ORC_REQ_OUTPUT) continue;
            int v = add_var(m, def->name, def->grid, def->req);
            int src;
            if (def->req == ORC_REQ_STATE) src = ORC_SRC_OWN_STATE;
            else if (producer[v] >= 0) src = ORC_SRC_UPSTREAM;
            else if (is_aggregate_name(m, def->name)) src = ORC_SRC_UPSTREAM;
            else src = ORC_SRC_EXOGENOUS;
            int i = n->n_in++;
            n->in_var[i] = v;
            n->in_src[i] = src;
            n->in_grid[i] = def->grid;
            n->in_factor[i] = 1.0;
            if (producer[v] >= 0) { add_edge(m, producer[v], gnode); has_dep = 1; }
            else if (is_aggregate_name(m, def->name)) {
                pend_node[n_pend] = gnode; pend_var[n_pend] = v; n_pend++; has_dep = 1;
            } else {
                m->vars[v].in_exogenous_list = 1;
            }
        }
        if (!has_dep) add_edge(m, 0, gnode); /* builder.rs:521-531 */
        /* provides (builder.rs:533-560), in outputs() order */
        for (int d = 0; d < k->n_defs; ++d) {
            const orc_def *def = &k->defs[d];
            if (def->req == ORC_REQ_INPUT) continue;
            int v = add_var(m, def->name, def->grid, def->req);
            int o = n->n_out++;
            n->out_var[o] = v;
            n->out_grid[o] = def->grid;
            if (producer[v] >= 0) add_edge(m, producer[v], gnode);
            producer[v] = gnode;
        }
    }
    for (int i = 0; i < m->n_uf; ++i) {
        if (m->uf_comp[i] < 0 || m->uf_comp[i] >= m->n_user) return fail(m, "unit factor: bad component");
        orc_node *n = &m->nodes[m->uf_comp[i]];
        int v = find_var(m, m->uf_var[i]);
        for (int j = 0; j < n->n_in; ++j)
            if (n->in_var[j] == v) n->in_factor[j] = m->uf_val[i];
    }

    if (m->has_schema) {
        /* schema variables: storage grid follows the schema (builder.rs:593-630) */
        for (int s = 0; s < m->n_schema; ++s) {
            int v = find_var(m, m->schema_name[s]);
            if (v < 0) {
                v = add_var(m, m->schema_name[s], m->schema_grid[s], ORC_REQ_INPUT);
                m->vars[v].in_exogenous_list = 1;
            } else if (m->vars[v].grid != m->schema_grid[s]) {
                m->vars[v].grid = m->schema_grid[s];
                if (producer[v] < 0) m->vars[v].in_exogenous_list = 1;
            }
        }
        /* aggregators (builder.rs:631-700); chained aggregates must be added in
         * dependency order by the caller (schema.topological_order_aggregates) */
        for (int ai = m->n_user; ai < m->n_nodes; ++ai) {
            orc_node *n = &m->nodes[ai];
            const int gnode = ai + 1;
            int has_dep = 0;
            n->n_in = n->agg_n;
            for (int i = 0; i < n->agg_n; ++i) {
                int v = find_var(m, n->agg_contrib[i]);
                if (v < 0) {
                    v = add_var(m, n->agg_contrib[i], n->agg_grid, ORC_REQ_INPUT);
                    m->vars[v].in_exogenous_list = 1;
                }
                n->in_var[i] = v;
                n->in_src[i] = ORC_SRC_EXOGENOUS; /* unused: aggregator reads at_end explicitly */
                n->in_grid[i] = n->agg_grid;
                n->in_factor[i] = 1.0;
                if (producer[v] >= 0) { add_edge(m, producer[v], gnode); has_dep = 1; }
            }
            if (!has_dep) add_edge(m, 0, gnode);
            int v = add_var(m, n->agg_name, n->agg_grid, ORC_REQ_OUTPUT);
            m->vars[v].grid = n->agg_grid;
            n->n_out = 1;
            n->out_var[0] = v;
            n->out_grid[0] = n->agg_grid;
            producer[v] = gnode;
        }
        for (int i = 0; i < n_pend; ++i)
            if (producer[pend_var[i]] >= 0) add_edge(m, producer[pend_var[i]], pend_node[i]);
    } else if (m->n_nodes != m->n_user) {
        return fail(m, "aggregates require a schema");
    }

    /* read-side transforms are only defined fine->coarse */
    for (int ni = 0; ni < m->n_nodes; ++ni) {
        const orc_node *n = &m->nodes[ni];
        for (int i = 0; i < n->n_in; ++i) {
            int sg = m->vars[n->in_var[i]].grid, wg = n->in_grid[i];
            if (sg == wg) continue;
            if (sg == ORC_GRID_FOUR_BOX) continue;
            if (sg == ORC_GRID_HEMISPHERIC && wg == ORC_GRID_SCALAR) continue;
            return fail(m, "unsupported read transform (coarse -> fine)");
        }
    }

    /* state variables need initial values (builder.rs:703-716) */
    for (int v = 0; v < m->n_vars; ++v) {
        orc_var *var = &m->vars[v];
        var->n_regions = orc_grid_regions(var->grid);
        var->endogenous = producer[v] >= 0;
        var->has_initial = 0;
        for (int i = 0; i < m->n_init; ++i)
            if (!strcmp(m->init_name[i], var->name)) { var->has_initial = 1; var->initial = m->init_val[i]; }
        if (var->req == ORC_REQ_STATE && !var->has_initial) {
            snprintf(m->err, sizeof m->err, "missing initial value for state variable '%s'", var->name);
            return -1;
        }
        /* exogenous data only if the name went through the `exogenous` list (builder.rs:751-755) */
        free(var->exo_data);
        var->exo_data = NULL;
        var->has_exo_data = 0;
        if (var->in_exogenous_list) {
            for (int i = 0; i < m->n_exo_in; ++i) {
                if (strcmp(m->exo_name[i], var->name)) continue;
                if (m->exo_grid[i] != var->grid) continue; /* wrong grid -> treated as absent */
                size_t n = (size_t)m->T * (size_t)var->n_regions;
                var->exo_data = (double *)malloc(n * sizeof(double));
                memcpy(var->exo_data, m->exo_vals[i], n * sizeof(double));
                var->has_exo_data = 1;
            }
        }
    }
    int64_t off = 0;
    for (int v = 0; v < m->n_vars; ++v) {
        m->vars[v].offset = off;
        off += (int64_t)m->T * m->vars[v].n_regions;
    }
    m->out_size = off;

    /* Execution order: petgraph Bfs from the Null root (model/runtime.rs:504-510).
     * Graph::neighbors() walks a node's outgoing edges most-recently-added first. */
    {
        int visited[ORC_MAX_NODES + 1] = {0};
        int queue[ORC_MAX_NODES + 1], qh = 0, qt = 0;
        queue[qt++] = 0;
        visited[0] = 1;
        m->n_order = 0;
        while (qh < qt) {
            int u = queue[qh++];
            if (u != 0) m->order[m->n_order++] = u - 1;
            for (int e = m->n_edges - 1; e >= 0; --e) {
                if (m->e_from[e] != u) continue;
                int w = m->e_to[e];
                if (!visited[w]) { visited[w] = 1; queue[qt++] = w; }
            }
        }
    }
    /* cycle check (builder.rs:563) via Kahn */
    {
        int indeg[ORC_MAX_NODES + 1] = {0}, done = 0, total = m->n_nodes + 1;
        int removed[ORC_MAX_NODES + 1] = {0};
        for (int e = 0; e < m->n_edges; ++e) indeg[m->e_to[e]]++;
        for (;;) {
            int progressed = 0;
            for (int u = 0; u < total; ++u) {
                if (removed[u] || indeg[u]) continue;
                removed[u] = 1; done++; progressed = 1;
                for (int e = 0; e < m->n_edges; ++e)
                    if (m->e_from[e] == u) indeg[m->e_to[e]]--;
            }
            if (!progressed) break;
        }
        if (done != total) return fail(m, "component graph contains a cycle");
    }
    m->built = 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* introspection                                                              */
/* ------------------------------------------------------------------------- */

int orc_n_variables(const orc_model *m) { return m->n_vars; }
const char *orc_variable_name(const orc_model *m, int v) { return m->vars[v].name; }
int orc_variable_grid(const orc_model *m, int v) { return m->vars[v].grid; }
int orc_variable_index(const orc_model *m, const char *name) { return find_var(m, name); }
int orc_variable_is_endogenous(const orc_model *m, int v) { return m->vars[v].endogenous; }
int orc_n_nodes(const orc_model *m) { return m->n_nodes; }
int orc_execution_order(const orc_model *m, int *order)
{
    for (int i = 0; i < m->n_order; ++i) order[i] = m->order[i];
    return m->n_order;
}
int orc_node_kind(const orc_model *m, int node) { return m->nodes[node].kind; }
int orc_variable_source(const orc_model *m, int component, const char *variable)
{
    const orc_node *n = &m->nodes[component];
    int v = find_var(m, variable);
    for (int i = 0; i < n->n_in; ++i)
        if (n->in_var[i] == v) return n->in_src[i];
    return -1;
}
int orc_n_times(const orc_model *m) { return m->T; }
int64_t orc_variable_offset(const orc_model *m, int v) { return m->vars[v].offset; }
int64_t orc_output_size(const orc_model *m) { return m->out_size; }
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* run                                                                        */
/* ------------------------------------------------------------------------- */

typedef struct {
    const double *params[ORC_MAX_NODES];  /* per node parameter block for this member */
    const double *initial;                /* per variable initial override or NULL */
    const uint8_t *has_initial;
    const double *const *exo;             /* per variable exogenous override [T*R] or NULL */
} run_inputs;

/* Storage initialisation — builder.rs:735-830: every series NaN-filled;
 * exogenous data copied (interpolate_into onto the same axis is the identity);
 * otherwise the initial value (if any) at index 0, broadcast to all regions. */
static void init_storage(const orc_model *m, const run_inputs *ri, double *data)
{
    for (int v = 0; v < m->n_vars; ++v) {
        const orc_var *var = &m->vars[v];
        double *blk = data + var->offset;
        const int64_t n = (int64_t)m->T * var->n_regions;
        const double *exo = (ri->exo && ri->exo[v]) ? ri->exo[v] : (var->has_exo_data ? var->exo_data : NULL);
        if (exo && var->in_exogenous_list) {
            memcpy(blk, exo, (size_t)n * sizeof(double));
            continue;
        }
        for (int64_t i = 0; i < n; ++i) blk[i] = NAN;
        int has = var->has_initial;
        double iv = var->initial;
        if (ri->has_initial && ri->has_initial[v]) { has = 1; iv = ri->initial[v]; }
        if (has)
            for (int r = 0; r < var->n_regions; ++r) blk[r] = iv;
    }
}

/* Model::run / step / step_model_component — model/runtime.rs:368-527 */
static int run_one(const orc_model *m, const run_inputs *ri, double *data, void **states)
{
    int failed = 0;
    init_storage(m, ri, data);
    for (int ni = 0; ni < m->n_nodes; ++ni) {
        const orc_kind_info *k = orc_kind_lookup(m->nodes[ni].kind);
        if (k->state_size && k->init_state) k->init_state(ri->params[ni], states[ni]);
    }
    double out[4 * ORC_MAX_DEFS];
    for (int N = 0; N < m->T - 1; ++N) {
        const double t0 = m->bounds[N], t1 = m->bounds[N + 1];
        for (int oi = 0; oi < m->n_order; ++oi) {
            const int ni = m->order[oi];
            const orc_node *n = &m->nodes[ni];
            orc_ctx c = {m, n, data, N};
            if (n->kind == ORC_AGGREGATOR) {
                aggregator_solve(n, &c, out);
            } else {
                const orc_kind_info *k = orc_kind_lookup(n->kind);
                if (k->solve(ri->params[ni], &c, t0, t1, out, states ? states[ni] : NULL)) {
                    failed = 1; /* "Solving failed": outputs stay NaN, run continues (:493-495) */
                    continue;
                }
            }
            /* outputs are packed region-major per output definition */
            int pos = 0;
            for (int o = 0; o < n->n_out; ++o) {
                const orc_var *var = &m->vars[n->out_var[o]];
                const int Rc = orc_grid_regions(n->out_grid[o]);
                double *row = data + var->offset + (int64_t)(N + 1) * var->n_regions;
                if (Rc == var->n_regions) {
                    for (int r = 0; r < Rc; ++r) row[r] = out[pos + r];
                } else if (n->out_grid[o] == ORC_GRID_FOUR_BOX && var->grid == ORC_GRID_SCALAR) {
                    /* write-side aggregation: model/transformations.rs:31-128 */
                    double s = 0.0;
                    for (int r = 0; r < 4; ++r) s += out[pos + r] * m->w_fourbox[r];
                    row[0] = s;
                } else if (n->out_grid[o] == ORC_GRID_FOUR_BOX && var->grid == ORC_GRID_HEMISPHERIC) {
                    const double *w = m->w_fourbox;
                    double wn = w[0] + w[1], ws = w[2] + w[3];
                    row[0] = (out[pos + 0] * w[0] + out[pos + 1] * w[1]) / wn;
                    row[1] = (out[pos + 2] * w[2] + out[pos + 3] * w[3]) / ws;
                } else if (n->out_grid[o] == ORC_GRID_HEMISPHERIC && var->grid == ORC_GRID_SCALAR) {
                    row[0] = out[pos] * m->w_hemi[0] + out[pos + 1] * m->w_hemi[1];
                }
                pos += Rc;
            }
        }
    }
    return failed;
}

static void **alloc_states(const orc_model *m)
{
    void **st = (void **)calloc((size_t)m->n_nodes, sizeof(void *));
    for (int ni = 0; ni < m->n_nodes; ++ni) {
        const orc_kind_info *k = orc_kind_lookup(m->nodes[ni].kind);
        if (k->state_size) st[ni] = calloc(1, k->state_size);
    }
    return st;
}
static void free_states(const orc_model *m, void **st)
{
    for (int ni = 0; ni < m->n_nodes; ++ni) free(st[ni]);
    free(st);
}

int orc_run(orc_model *m, double *out)
{
    if (!m->built) return fail(m, "model not built");
    run_inputs ri;
    memset(&ri, 0, sizeof ri);
    for (int ni = 0; ni < m->n_nodes; ++ni) ri.params[ni] = m->nodes[ni].params;
    void **st = alloc_states(m);
    run_one(m, &ri, out, st);
    free_states(m, st);
    return 0;
}

typedef struct {
    double *data;
    double *pblock; /* [n_nodes][ORC_MAX_PARAMS] */
    double *initial;
    uint8_t *has_initial;
    const double **exo;
    void **states;
} worker;

static worker worker_new(const orc_model *m)
{
    worker w;
    w.data = (double *)malloc((size_t)m->out_size * sizeof(double));
    w.pblock = (double *)malloc((size_t)m->n_nodes * ORC_MAX_PARAMS * sizeof(double));
    w.initial = (double *)calloc((size_t)m->n_vars, sizeof(double));
    w.has_initial = (uint8_t *)calloc((size_t)m->n_vars, 1);
    w.exo = (const double **)calloc((size_t)m->n_vars, sizeof(double *));
    w.states = alloc_states(m);
    return w;
}
static void worker_free(const orc_model *m, worker *w)
{
    free(w->data); free(w->pblock); free(w->initial); free(w->has_initial); free((void *)w->exo);
    free_states(m, w->states);
}

/* one (member, scenario) run into the worker's storage */
static int worker_run(const orc_model *m, worker *w, int n_cols, const int *bc, const int *bi,
                      const double *prow, int n_exo, const int *exo_vars, const double *scen)
{
    run_inputs ri;
    memset(&ri, 0, sizeof ri);
    for (int ni = 0; ni < m->n_nodes; ++ni) {
        double *pb = w->pblock + (size_t)ni * ORC_MAX_PARAMS;
        memcpy(pb, m->nodes[ni].params, sizeof(double) * ORC_MAX_PARAMS);
        ri.params[ni] = pb;
    }
    memset(w->has_initial, 0, (size_t)m->n_vars);
    for (int j = 0; j < n_cols; ++j) {
        if (bc[j] >= 0) w->pblock[(size_t)bc[j] * ORC_MAX_PARAMS + bi[j]] = prow[j];
        else if (bc[j] == -1) { w->initial[bi[j]] = prow[j]; w->has_initial[bi[j]] = 1; }
    }
    ri.initial = w->initial;
    ri.has_initial = w->has_initial;
    for (int v = 0; v < m->n_vars; ++v) w->exo[v] = NULL;
    if (scen) {
        int64_t off = 0;
        for (int e = 0; e < n_exo; ++e) {
            w->exo[exo_vars[e]] = scen + off;
            off += (int64_t)m->T * m->vars[exo_vars[e]].n_regions;
        }
    }
    ri.exo = w->exo;
    return run_one(m, &ri, w->data, w->states);
}

static int64_t scenario_stride(const orc_model *m, int n_exo, const int *exo_vars)
{
    int64_t s = 0;
    for (int e = 0; e < n_exo; ++e) s += (int64_t)m->T * m->vars[exo_vars[e]].n_regions;
    return s;
}

/* DefaultModelRunner::run_batch — crates/rscm-calibrate/src/model_runner.rs:223-266:
 * members are independent; OpenMP static schedule stands in for rayon par_iter. */
int orc_run_batch(const orc_model *m, int n_cols, const int *bind_component, const int *bind_index,
                  const double *params, int64_t M, int n_exo, const int *exo_vars,
                  const double *scenarios, int64_t S, int n_out, const int *out_vars, double *out,
                  uint8_t *status, int n_threads)
{
    if (!m->built) return -1;
    const int64_t Seff = (scenarios && S > 0) ? S : 1;
    const int64_t runs = Seff * M;
    const int64_t sstride = scenario_stride(m, n_exo, exo_vars);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        worker w = worker_new(m);
#pragma omp for schedule(static)
        for (int64_t run = 0; run < runs; ++run) {
            const int64_t s = run / M, mi = run % M;
            const double *scen = (scenarios && S > 0) ? scenarios + s * sstride : NULL;
            int failed = worker_run(m, &w, n_cols, bind_component, bind_index,
                                    params ? params + mi * n_cols : NULL, n_exo, exo_vars, scen);
            if (status) status[run] = (uint8_t)failed;
            int64_t row = 0;
            for (int o = 0; o < n_out; ++o) {
                const orc_var *var = &m->vars[out_vars[o]];
                const int64_t n = (int64_t)m->T * var->n_regions;
                const double *src = w.data + var->offset;
                for (int64_t i = 0; i < n; ++i) out[(row + i) * runs + run] = src[i];
                row += n;
            }
        }
        worker_free(m, &w);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* calibration                                                                */
/* ------------------------------------------------------------------------- */

#define ORC_LN_2PI_HALF (0.5 * log(2.0 * 3.14159265358979323846))

/* Distribution::ln_pdf — crates/rscm-calibrate/src/distribution.rs */
double orc_ln_pdf(const orc_prior *p, double x)
{
    switch (p->kind) {
    case ORC_PRIOR_NONE: return 0.0;
    case ORC_PRIOR_UNIFORM: /* :157-163 */
        if (x < p->a || x > p->b) return -INFINITY;
        return -log(p->b - p->a);
    case ORC_PRIOR_NORMAL: { /* :256-259 */
        double z = (x - p->a) / p->b;
        return -0.5 * z * z - log(p->b) - 0.5 * log(2.0 * 3.14159265358979323846);
    }
    case ORC_PRIOR_LOGNORMAL: { /* :353-360 */
        if (x <= 0.0) return -INFINITY;
        double ln_x = log(x);
        double z = (ln_x - p->a) / p->b;
        return -0.5 * z * z - ln_x - log(p->b) - 0.5 * log(2.0 * 3.14159265358979323846);
    }
    case ORC_PRIOR_BOUND_NORMAL:
    case ORC_PRIOR_BOUND_LOGNORMAL:
    case ORC_PRIOR_BOUND_UNIFORM: { /* Bound :490-497 (unnormalised inner pdf) */
        if (x < p->low || x > p->high) return -INFINITY;
        orc_prior inner = *p;
        inner.kind = p->kind == ORC_PRIOR_BOUND_NORMAL ? ORC_PRIOR_NORMAL
                   : p->kind == ORC_PRIOR_BOUND_LOGNORMAL ? ORC_PRIOR_LOGNORMAL : ORC_PRIOR_UNIFORM;
        return orc_ln_pdf(&inner, x);
    }
    default: return NAN;
    }
}

/* ParameterSet::log_prior — parameter_set.rs:255-270 */
double orc_log_prior(const orc_prior *priors, int n, const double *x)
{
    double lp = 0.0;
    for (int i = 0; i < n; ++i) lp += orc_ln_pdf(&priors[i], x[i]);
    return lp;
}

/* time_key = format!("{:.6}", time) — likelihood.rs:40-42 */
int orc_time_index(const orc_model *m, double time)
{
    char key[64], k2[64];
    snprintf(key, sizeof key, "%.6f", time);
    int found = -1;
    /* HashMap insert: a later equal key overwrites an earlier one */
    for (int i = 0; i < m->T; ++i) {
        snprintf(k2, sizeof k2, "%.6f", m->bounds[i]);
        if (!strcmp(key, k2)) found = i;
    }
    return found;
}

/* GaussianLikelihood::ln_likelihood — likelihood.rs:186-253, fed by
 * DefaultModelRunner::extract_outputs (model_runner.rs:161-216): NaN model
 * values are never inserted, so a NaN at an observed time is "missing time"
 * => Err; +-inf is "non-finite" => Err.  Err is reported as NaN here. */
double orc_ln_likelihood(const orc_model *m, const double *run_out, const orc_obs *obs, int64_t K,
                         int normalize)
{
    double total = 0.0;
    /* per-variable partial sums in target order, then summed (likelihood.rs:237-253) */
    int64_t i = 0;
    while (i < K) {
        const int v = obs[i].variable;
        double ln_l = 0.0;
        for (; i < K && obs[i].variable == v; ++i) {
            if (obs[i].time_index < 0 || obs[i].time_index >= m->T) return NAN;
            const orc_var *var = &m->vars[v];
            const double model = run_out[var->offset + (int64_t)obs[i].time_index * var->n_regions];
            if (!isfinite(model)) return NAN;
            const double residual = obs[i].value - model;
            const double chi2 = (residual * residual) / (obs[i].sigma * obs[i].sigma);
            double l = -0.5 * chi2;
            if (normalize) {
                l -= 0.5 * log(2.0 * 3.14159265358979323846);
                l -= log(obs[i].sigma);
            }
            ln_l += l;
        }
        total += ln_l;
    }
    return total;
}

/* EnsembleSampler::log_posterior_batch — sampler/ensemble.rs:143-178 */
int orc_log_posterior_batch(const orc_model *m, int n_cols, const int *bind_component,
                            const int *bind_index, const double *params, int64_t M, int n_exo,
                            const int *exo_vars, const double *scenarios, int64_t S,
                            const orc_prior *priors, const orc_obs *obs, int64_t K, int normalize,
                            double *logpost, int n_threads)
{
    if (!m->built) return -1;
    const int64_t Seff = (scenarios && S > 0) ? S : 1;
    const int64_t runs = Seff * M;
    const int64_t sstride = scenario_stride(m, n_exo, exo_vars);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        worker w = worker_new(m);
#pragma omp for schedule(static)
        for (int64_t run = 0; run < runs; ++run) {
            const int64_t s = run / M, mi = run % M;
            const double *prow = params + mi * n_cols;
            const double *scen = (scenarios && S > 0) ? scenarios + s * sstride : NULL;
            /* the reference runs the model first, then the prior; the result is the same */
            double lp = priors ? orc_log_prior(priors, n_cols, prow) : 0.0;
            if (!isfinite(lp)) { logpost[run] = -INFINITY; continue; }
            worker_run(m, &w, n_cols, bind_component, bind_index, prow, n_exo, exo_vars, scen);
            double ll = orc_ln_likelihood(m, w.data, obs, K, normalize);
            logpost[run] = isnan(ll) ? -INFINITY : lp + ll;
        }
        worker_free(m, &w);
    }
    return 0;
}
