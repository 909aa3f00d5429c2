// graph.hpp — host-side graph compiler of the ensemble engine.
//
// Reproduces, once per ensemble, what the reference does once per member:
// ModelBuilder::build (crates/rscm-core/src/model/builder.rs:418-860) — variable
// table, insertion-order-dependent VariableSource classification, dependency
// edges, schema aggregators — and the petgraph BFS execution order
// (model/runtime.rs:504-510).  The result is lowered to (a) flat tables the
// kernel arguments are filled from and (b) the source text of a `Prog` struct:
// the component graph as straight-line device code (see kernel.cuh).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rscm_b200.h"

namespace rscm {

enum Req { REQ_INPUT = 0, REQ_OUTPUT = 1, REQ_STATE = 2 };
constexpr int KIND_AGGREGATOR = 4;

struct VarDef {
    const char *name;
    int req;
    int grid;
};

struct KindInfo {
    int kind;
    const char *type_name; // Rust struct name (Debug prefix used by the reference for lookups)
    const char *dev_name;  // prefix of <dev_name>_prepare / <dev_name>_solve in components.cuh
    std::vector<VarDef> defs;
    std::vector<const char *> param_names;
    int n_derived;
    int rk_step_param; // index of the RK4 step-size parameter, -1: fixed 0.1, -2: no RK4
    std::vector<int> bindable; // 1 if the parameter may be bound to a member column
    int reg_weight;            // rough register appetite; decides the kernel's min-CTAs-per-SM hint
    // stateful kinds (ClimateUDEB, ...): sizes of the per-thread state the fused kernel provides
    int n_state = 0;       // register/local values (R S[])
    int n_smem = 0;        // per-thread shared-memory scratch, in 8-byte words (both compute dtypes); it holds nothing
                           // between two solves, so the nodes of a program share the same words (the program takes the largest)
    int scratch_per_T = 0; // global scratch rows per time point (member-interleaved)
    bool needs_time = false; // solve uses the time bounds
    // input access override: (input index, mode) per `in[]` entry group; mode 0 = get(), 1 = at_start, 2 = at_end.
    // empty = one get() per input (state inputs at_start)
    std::vector<std::pair<int, int>> in_access = {};
    // per-graph constant table computed on the host from the (non-bindable) geometry parameters
    std::vector<double> (*const_table)(const std::vector<double> &params, std::string &err) = nullptr;
    // large per-graph table kept in global memory (n_times is passed for tables sized by the run length)
    // (n_times and the time bounds [n_times + 1] are passed for tables sized by, or derived from, the time axis)
    std::vector<double> (*global_table)(const std::vector<double> &params, int n_times, const double *bounds, std::string &err) = nullptr;
    int aux_param = -1; // index of an integer parameter handed to the device code as a compile-time literal
    int scratch_fixed = 0; // global scratch rows that do not scale with the run length (come first in the node's rows)
    bool no_slots = false; // all parameters are per-graph and only feed const_table: they take no kernel parameter slots
    int lanes = 1; // lanes of a warp that work on one member (ClimateUDEB: 4); the program takes the largest of its kinds
    bool aux_template = false; // the aux literal is also a template argument of <dev_name>_solve / _init_state
    int n_smem_lanes = 0; // extra per-thread shared-memory words when the PROGRAM runs lane groups (Prog::LANES > 1)
    // In a lane-group program only role 0 (warp 0 of the CTA) runs the component graph; a kind with lanes > 1 or
    // lane_aware is entered by all four warps: the emitter broadcasts its inputs through the exchange area, the kind
    // may use n_xch further exchange slots (doubles per member) of its own.
    bool lane_aware = false;
    int n_xch = 0;
    // the kind's solve is a template on its output argument: the emitter hands it a view that writes each output value into
    // its cell of the next time level, instead of a local array it copies out afterwards (kinds with dozens of outputs)
    bool scatter_out = false;
};

const KindInfo *kind_info(int kind);

struct Variable {
    std::string name;
    int grid = 0;
    int n_regions = 1;
    int req = REQ_INPUT;
    bool endogenous = false;
    bool in_exogenous_list = false;
    bool has_initial = false;
    double initial = 0.0;
    int cell0 = 0;   // first storage cell
    int exo_index = -1; // position among exogenous variables, -1 if endogenous
    int exo_row0 = -1;  // first staged row
};

struct Node {
    int kind = 0;
    std::vector<double> params;
    int param_base = 0;   // first parameter slot
    int derived_base = 0; // first derived-constant slot
    int rk_table = -1;    // row of the sub-step table
    int state_base = 0, smem_base = 0, scratch_base = 0, ctab_base = 0; // offsets of this node's stateful storage
    int xch_base = 0, xch_user = 0; // exchange slots of a lane node: broadcast inputs from xch_base, the kind's own from xch_user
    bool lane_node = false; // entered by all roles of a lane-group program
    int gtab_base = 0, aux = 0;
    std::vector<int> in_var, in_src, in_grid;
    std::vector<double> in_factor;
    std::vector<int> out_var, out_grid;
    // aggregator
    std::string agg_name;
    int agg_op = 0, agg_grid = 0;
    std::vector<std::string> agg_contrib;
    std::vector<double> agg_w;
};

struct Graph {
    std::vector<Variable> vars;
    std::vector<Node> nodes; // components in insertion order, then aggregators
    int n_user = 0;
    std::vector<int> order;     // node ids in execution order
    std::vector<int> exo_vars;  // variable ids, scenario order
    int n_cells = 0, n_slots = 0, n_derived = 0, n_exo_rows = 0, n_rk = 0;
    int lanes = 1;         // Prog::LANES: threads per member
    int n_xch = 0;         // Prog::NXCH: exchange slots (doubles per member) of a lane-group program
    bool stage_exo = true; // Prog::STAGE_EXO: exogenous rows staged in shared memory (else read from global)
    int n_state = 0, n_smem = 0, n_scratch_rows = 0; // stateful components: totals (scratch rows already x T)
    bool needs_time = false;
    std::vector<double> ctab; // concatenated per-graph constant tables (even length)
    std::vector<double> gtab; // concatenated large tables (global memory)
    std::vector<double> slot_default;
    std::vector<int> slot_bindable;
    std::vector<int> cell_var, cell_region;
    int T = 0;
    std::vector<double> bounds;
    double w_fourbox[4] = {0.25, 0.25, 0.25, 0.25};
    double w_hemi[2] = {0.5, 0.5};
    bool custom_w_fourbox = false, custom_w_hemi = false; // set through ModelBuilder::with_grid_weights
    std::vector<std::vector<int>> rk_nsub; // [n_rk][T]
    std::string program_source;            // emitted Prog body (without the struct name)
    std::string signature;                 // canonical key = hash-free text of the program

    int find_var(const std::string &name) const;
    // slot id for "<Type>.<field>", "<Type>#i.<field>"; cell id for "initial:<var>" (returned as -(cell)-1); -1e9 if unknown
    int resolve_slot(const std::string &slot, std::string &err) const;
};

// Builds the graph; returns false and fills err on failure.
bool compile_graph(const rscm_b200_graph_desc &desc, Graph &g, std::string &err);

// RK4 sub-step count per time step and get_last_step check
// (ode_solvers 0.6.1 Rk4::integrate; crates/rscm-core/src/ivp/mod.rs:90-102):
// returns n >= 1, or -1 when the reference would panic.
int rk4_substeps(double t0, double t1, double h);

// "{:.6}" key equality (crates/rscm-calibrate/src/likelihood.rs:40-42)
int time_index_for(const Graph &g, double time);

} // namespace rscm
