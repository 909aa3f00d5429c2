/* orc_internal.h — private structures of the CPU oracle (test infrastructure, not product). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include "rscm_oracle.h"
#include <math.h>
#include <stddef.h>

#define ORC_MAX_NAME 96
#define ORC_MAX_DEFS 96
#define ORC_MAX_PARAMS 320
#define ORC_MAX_NODES 48
#define ORC_MAX_VARS 224
#define ORC_MAX_EDGES 1024
#define ORC_MAX_CONTRIB 16

typedef struct {
    const char *name;
    int req;  /* ORC_REQ_* */
    int grid; /* ORC_GRID_* */
} orc_def;

struct orc_ctx;

/* A component kind: static definition table + solve().  `in` / `out` are
 * indexed in inputs() / outputs() order = definitions() filtered by
 * (Input|State) / (Output|State), crates/rscm-core/src/component.rs:357-383,
 * with definitions() ordered inputs, outputs, states
 * (crates/rscm-macros/src/lib.rs:630-636). */
typedef struct {
    int kind;
    const char *type_name; /* Rust struct name = Debug prefix used for lookups */
    int n_defs;
    const orc_def *defs;
    int n_params;
    /* returns 0 ok, 1 = component Err (outputs not written, run continues:
     * model/runtime.rs:493-495) */
    int (*solve)(const double *p, struct orc_ctx *c, double t0, double t1, double *out, void *state);
    size_t state_size;
    void (*init_state)(const double *p, void *state);
} orc_kind_info;

const orc_kind_info *orc_kind_lookup(int kind);

typedef struct {
    int kind;
    double params[ORC_MAX_PARAMS];
    int n_params;
    /* aggregator only */
    int agg_op, agg_grid, agg_n;
    char agg_name[ORC_MAX_NAME];
    char agg_contrib[ORC_MAX_CONTRIB][ORC_MAX_NAME];
    double agg_w[ORC_MAX_CONTRIB];
    /* resolved at build */
    int n_in, n_out;
    int in_var[ORC_MAX_DEFS], in_src[ORC_MAX_DEFS], in_grid[ORC_MAX_DEFS];
    double in_factor[ORC_MAX_DEFS];
    int out_var[ORC_MAX_DEFS], out_grid[ORC_MAX_DEFS];
} orc_node;

typedef struct {
    char name[ORC_MAX_NAME];
    int grid;
    int n_regions;
    int req; /* requirement type of first definition */
    int endogenous;
    int in_exogenous_list;
    int has_initial;
    double initial;
    int has_exo_data;
    double *exo_data; /* [T][R] */
    int64_t offset;   /* into a run's output block, doubles */
} orc_var;

struct orc_model {
    char err[256];
    int built;
    int n_nodes; /* user components first, then aggregators */
    int n_user;
    orc_node nodes[ORC_MAX_NODES];
    int n_vars;
    orc_var vars[ORC_MAX_VARS];
    int n_schema;
    char schema_name[ORC_MAX_VARS][ORC_MAX_NAME];
    int schema_grid[ORC_MAX_VARS];
    int has_schema;
    int n_init;
    char init_name[ORC_MAX_VARS][ORC_MAX_NAME];
    double init_val[ORC_MAX_VARS];
    int n_exo_in;
    char exo_name[ORC_MAX_VARS][ORC_MAX_NAME];
    int exo_grid[ORC_MAX_VARS];
    double *exo_vals[ORC_MAX_VARS];
    int n_uf;
    int uf_comp[ORC_MAX_VARS];
    char uf_var[ORC_MAX_VARS][ORC_MAX_NAME];
    double uf_val[ORC_MAX_VARS];
    double w_fourbox[4], w_hemi[2];
    int has_w_fourbox, has_w_hemi;
    int T;
    double *bounds; /* [T+1] */
    /* graph: node id 0 = Null root, user node i = i+1 */
    int n_edges;
    int e_from[ORC_MAX_EDGES], e_to[ORC_MAX_EDGES];
    int order[ORC_MAX_NODES];
    int n_order;
    int64_t out_size;
};

/* per-run execution context handed to solve() */
typedef struct orc_ctx {
    const struct orc_model *m;
    const orc_node *node;
    double *data; /* run storage [V] blocks of [T][R] */
    int N;        /* current time index */
} orc_ctx;

static inline int orc_grid_regions(int grid)
{
    return grid == ORC_GRID_FOUR_BOX ? 4 : (grid == ORC_GRID_HEMISPHERIC ? 2 : 1);
}

/* window accessors (state/windows.rs:155-270 and grid equivalents) */
double orc_in_at(const orc_ctx *c, int i, int region, int index); /* raw index read incl. transforms; NaN if OOB */
double orc_in_start(const orc_ctx *c, int i, int region);
double orc_in_end(const orc_ctx *c, int i, int region, int *ok);
double orc_in_get(const orc_ctx *c, int i, int region);
double orc_in_offset(const orc_ctx *c, int i, int region, int off, int *ok);
double orc_in_latest(const orc_ctx *c, int i, int region);

/* numerics */
typedef void (*orc_rhs)(const void *self, const double *y, double *dy);
/* returns 0 ok; 1 if get_last_step's |t_last - t_next| < 5e-3 assertion would fire */
int orc_rk4(orc_rhs f, const void *self, int dim, double t0, double t1, double h, double *y);

/* magicc kinds, registered from magicc_*.c */
extern const orc_kind_info orc_kind_ghg_forcing;
extern const orc_kind_info orc_kind_ozone_forcing;
extern const orc_kind_info orc_kind_aerosol_direct;
extern const orc_kind_info orc_kind_aerosol_indirect;
extern const orc_kind_info orc_kind_climate_udeb;
extern const orc_kind_info orc_kind_fbohu, orc_kind_ospp, orc_kind_co2_budget, orc_kind_terrestrial, orc_kind_ch4, orc_kind_n2o;
extern const orc_kind_info orc_kind_ocean_carbon;
const orc_kind_info *orc_kind_halocarbon(void); /* defs are generated from the species list: magicc_halocarbon.c */

#endif
