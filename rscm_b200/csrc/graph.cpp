// graph.cpp — host graph compiler (see graph.hpp).
#include "graph.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <sstream>

namespace rscm {

// ---------------------------------------------------------------------------
// Static descriptor tables: the output of the reference's #[derive(ComponentIO)]
// (crates/rscm-macros/src/lib.rs:356-678) restated per kind.  Order of defs =
// inputs, outputs, states (lib.rs:630-636).
// ---------------------------------------------------------------------------
// ClimateUDEB per-graph tables: af_top[n], af_bottom[n], af_diff[n] (compute_area_factors /
// ocean_area_at_depth, crates/rscm-magicc/src/parameters/climate_udeb.rs) and the initial ocean profile of both
// hemispheres (initial_ocean_profile, same file; CMIP5 defaults or the analytical exponential profile).
#include "cmip5_profiles.inc"
static std::vector<double> udeb_const_table(const std::vector<double> &p, std::string &err)
{
    const int n = static_cast<int>(p[0]);
    if (n < 2 || n > 50 || p[0] != n) { err = "ClimateUDEB: n_layers must be an integer in [2, 50]"; return {}; }
    const int steps = static_cast<int>(p[35]);
    if (steps < 1 || p[35] != steps) { err = "ClimateUDEB: steps_per_year must be a positive integer"; return {}; }
    const double mld = p[1], dz = p[2], dda = p[21];
    auto area_at = [&](double depth) {
        static const double D[12] = {0.0, 200.0, 500.0, 1000.0, 1500.0, 2000.0, 2500.0, 3000.0, 3500.0, 4000.0, 4500.0, 5000.0};
        static const double A[12] = {1.0, 0.975, 0.95, 0.92, 0.91, 0.87, 0.81, 0.72, 0.55, 0.38, 0.18, 0.05};
        double hydro;
        if (depth <= D[0]) hydro = A[0];
        else if (depth >= D[11]) hydro = A[11];
        else {
            hydro = A[0];
            for (int i = 1; i < 12; ++i)
                if (depth <= D[i]) { hydro = A[i - 1] + (depth - D[i - 1]) / (D[i] - D[i - 1]) * (A[i] - A[i - 1]); break; }
        }
        return 1.0 + dda * (hydro - 1.0);
    };
    std::vector<double> t(5 * n, 0.0);
    for (int l = 0; l < n; ++l) {
        double zt, zb;
        if (l == 0) { zt = 0.0; zb = mld; }
        else { zt = mld + (static_cast<double>(l) - 1.0) * dz; zb = zt + dz; }
        const double at = area_at(zt), ab = area_at(zb), avg = (at + ab) / 2.0;
        t[l] = at / avg;
        t[n + l] = ab / avg;
        t[2 * n + l] = (at - ab) / avg;
    }
    // The initial profile only enters the variable-upwelling entrainment terms of step_hemisphere
    // (ocean_column.rs:150-205), always in the same geometry-only combinations, so the table holds those:
    //   g[0] = (init[1] - Tp) * af_bottom[0];  g[i] = init[i+1]*af_bottom[i] - init[i]*af_top[i] + Tp*af_diff[i];
    //   g[n-1] = (Tp - init[n-1]) * af_top[n-1]          (Tp = polar sinking temperature = 1, state.rs default)
    // followed by omr[l] = 1 - depth_l / total_depth, the relative-depth factor of layer_diffusivities (:23-52).
    t.resize(6 * n, 0.0);
    const double tp = 1.0;
    for (int h = 0; h < 2; ++h) {
        std::vector<double> init(n);
        for (int l = 0; l < n; ++l) {
            double v;
            if (p[34] == 2.0) v = (h == 0 ? CMIP5_NH : CMIP5_SH)[l < 50 ? l : 49];
            else {
                const double kap = p[3] * 3155.76;
                v = (l == 0) ? 17.2 : 1.0 + (17.2 - 1.0) * std::exp(-p[6] * ((static_cast<double>(l) - 1.0) * dz + 0.5 * dz) / kap);
            }
            init[l] = v;
        }
        double *g = &t[(3 + h) * n];
        g[0] = (init[1] - tp) * t[n];
        for (int i = 1; i < n - 1; ++i) g[i] = init[i + 1] * t[n + i] - init[i] * t[i] + tp * t[2 * n + i];
        g[n - 1] = (tp - init[n - 1]) * t[n - 1];
    }
    const double total_depth = mld + (static_cast<double>(n) - 1.0) * dz;
    for (int l = 0; l < n; ++l) t[5 * n + l] = 1.0 - (mld + static_cast<double>(l) * dz) / total_depth;
    // device layout (climate_udeb.cuh): one table per lane role q = 2*end + hemisphere, UDEB_MAXR = 25 rows of
    // UDEB_CT = 6 doubles in SWEEP order (top sweep: row j = layer j; bottom sweep: row j = layer n-1-j):
    //   {near area, far area, af_diff, g_h, omr, af_top}
    // near / far = area factor of the coupling towards the previous / next row of the sweep (top: af_top x dz/dz_up and
    // af_bottom; bottom: the other way round; dz/dz_up = 2 for layer 1 — the half layer below the mixed layer — when it is
    // an interior layer); omr = relative-depth factor of the diffusivity the row has to compute (top: boundary below the
    // layer; bottom: boundary above it).  Row 0 of a role is the end row of the column: it uses af_top, af_bottom (far), g, omr.
    const int k = (n - 2) >> 1;
    std::vector<double> r(4 * 25 * 6, 0.0);
    for (int q = 0; q < 4; ++q) {
        const int h = q & 1;
        const bool bottom = (q & 2) != 0;
        const int nrows = bottom ? n - k - 1 : k + 1;
        for (int j = 0; j < nrows; ++j) {
            const int l = bottom ? n - 1 - j : j;
            double *row = &r[(q * 25 + j) * 6];
            const double at = t[l], ab = t[n + l], ad = t[2 * n + l], g = t[(3 + h) * n + l];
            const double at_up = at * ((l == 1 && l != n - 1) ? 2.0 : 1.0);
            row[0] = bottom ? ab : at_up;
            row[1] = bottom ? at_up : ab;
            row[2] = ad;
            row[3] = g;
            row[4] = bottom ? (l > 0 ? t[5 * n + l - 1] : 0.0) : t[5 * n + l];
            row[5] = at;
        }
    }
    return r;
}

// ClimateUDEB: the window of the cumulative-temperature feedback (adjusted_ecs, udeb/mod.rs:302-330).  At step N the
// history holds one entry per completed step; walking it backwards with years_remaining = feedback_cumt_period, entries
// [first, N) count fully and entry first-1 with weight `partial` when the window ends inside it.  Which entries those are
// follows from the time axis and the period alone, in exactly this arithmetic, so the host states it once per graph:
// {period, then per step N: first, partial}.  A member whose period differs (bound per member) walks the axis itself.
static std::vector<double> udeb_window_table(const std::vector<double> &p, int n_times, const double *bounds, std::string &)
{
    const double period = p[15];
    std::vector<double> t(2 + 2 * static_cast<size_t>(n_times), 0.0);
    t[0] = period;
    for (int n = 0; n < n_times; ++n) {
        double rem = period, partial = 0.0;
        int first = n;
        for (int i = n - 1; i >= 0; --i) {
            if (rem <= 0.0) break;
            const double dt = bounds[i + 1] - bounds[i];
            if (dt <= rem) { first = i; rem -= dt; }
            else { partial = rem / dt; rem = 0.0; }
        }
        t[2 + 2 * n] = first;
        t[3 + 2 * n] = partial;
    }
    return t;
}

// OceanCarbon: scaled impulse-response function by lag in months, irf(k/12) for k = 0 .. steps*(T-1)
// (OceanCarbonParameters::irf / scale_irf, IrfForm::evaluate — crates/rscm-magicc/src/parameters/ocean_carbon.rs:99-130,378-397)
static std::vector<double> ocean_irf_table(const std::vector<double> &p, int n_times, const double *, std::string &err)
{
    const int steps = static_cast<int>(p[10]);
    if (steps < 1 || steps > 16 || p[10] != steps) { err = "OceanCarbon: steps_per_year must be an integer in [1, 16]"; return {}; }
    const long long months = static_cast<long long>(steps) * (n_times - 1);
    // flux_history is a deque bounded by max_history_months (ocean.rs:224-228): a flux older than that has left the
    // convolution.  The table states this as zero weights from lag max_history_months on (and carries a zero margin for
    // the sliding windows the device code loads).
    const double max_hist = p[11];
    if (!(max_hist >= 1.0)) { err = "OceanCarbon: max_history_months must be at least 1"; return {}; }
    auto form = [](const double *f, double t) {
        const int n = static_cast<int>(f[1]);
        if (f[0] == 0.0) {
            double r = 0.0;
            for (int i = n - 1; i >= 0; --i) r = r * t + f[2 + i];
            return r;
        }
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += f[2 + i] * std::exp(-t / f[10 + i]);
        return s;
    };
    if (p[14] < 1 || p[14] > 8 || p[32] < 1 || p[32] > 8) { err = "OceanCarbon: IRF forms take 1..8 terms"; return {}; }
    std::vector<double> tab(static_cast<size_t>(months) + 2 + 96, 0.0);
    for (size_t k = 0; k < static_cast<size_t>(months) + 2; ++k) {
        if (static_cast<double>(k) >= max_hist) break; // lag k means k + 1 entries of history
        const double t = static_cast<double>(k) * (1.0 / 12.0);
        const double raw = (t < p[12]) ? form(&p[13], t) : form(&p[31], t);
        tab[k] = (raw * p[6]) / (raw * p[6] + 1.0 - raw);
    }
    return tab;
}

// HalocarbonChemistry (crates/rscm-magicc/src/chemistry/halocarbon.rs): the reference's default species list
// (parameters/halocarbon.rs:203-262: 23 F-gases, then 18 Montreal gases).  All parameters are per-graph; the device
// code reads a per-species table {lifetime, conv*lifetime, radiative_efficiency/1000, concentration_pi,
// eesc weight = (n_cl + br_multiplier*n_br) * fractional_release/cfc11_release_normalisation (0 when the species
// releases nothing), is_fgas}.
static const char *const kHaloSpecies[41] = {
    "CF4", "C2F6", "C3F8", "C4F10", "C5F12", "C6F14", "C7F16", "C8F18", "c-C4F8", "HFC-23", "HFC-32", "HFC-43-10mee", "HFC-125",
    "HFC-134a", "HFC-143a", "HFC-152a", "HFC-227ea", "HFC-236fa", "HFC-245fa", "HFC-365mfc", "NF3", "SF6", "SO2F2",
    "CFC-11", "CFC-12", "CFC-113", "CFC-114", "CFC-115", "HCFC-22", "HCFC-141b", "HCFC-142b", "CH3CCl3", "CCl4", "CH3Cl", "CH2Cl2",
    "CHCl3", "CH3Br", "Halon-1211", "Halon-1301", "Halon-2402", "Halon-1202"};
constexpr int kHaloNS = 41, kHaloNF = 23, kHaloNG = 6;

static std::vector<double> halocarbon_const_table(const std::vector<double> &p, std::string &err)
{
    std::vector<double> t(6 * kHaloNS, 0.0);
    const double atm_mass_g = p[4] * 1e12;
    for (int s = 0; s < kHaloNS; ++s) {
        const double *sp = &p[kHaloNG + 7 * s];
        if (!(sp[0] > 0.0) || !(sp[3] > 0.0)) { err = std::string("HalocarbonChemistry: lifetime and molecular_weight of ") + kHaloSpecies[s] + " must be positive"; return {}; }
        const double conv = (p[3] / sp[3]) * (1e9 / atm_mass_g) * 1e12 / p[5]; // emission_to_concentration_factor :162-172
        t[6 * s + 0] = sp[0];
        t[6 * s + 1] = conv;
        t[6 * s + 2] = sp[1];
        t[6 * s + 3] = sp[2];
        t[6 * s + 4] = sp[6] > 0.0 ? sp[4] + p[0] * sp[5] : 0.0; // halogen loading, only for species that release (:211)
        t[6 * s + 5] = sp[6] > 0.0 ? sp[6] / p[1] : 0.0;           // normalised release
    }
    return t;
}

static KindInfo halocarbon_kind()
{
    static std::vector<std::string> names; // storage behind the const char* of the descriptor
    names.reserve(2 * kHaloNS + 7 * kHaloNS);
    KindInfo k{};
    k.kind = RSCM_B200_HALOCARBON_CHEMISTRY;
    k.type_name = "HalocarbonChemistry";
    k.dev_name = "halocarbon_chemistry";
    // definitions() is hand-written in the reference (halocarbon.rs:259-293): emissions input + concentration state
    // per species, then the four outputs
    for (int s = 0; s < kHaloNS; ++s) {
        names.push_back(std::string("Emissions|") + kHaloSpecies[s]);
        k.defs.push_back({names.back().c_str(), REQ_INPUT, RSCM_B200_SCALAR});
        names.push_back(std::string("Atmospheric Concentration|") + kHaloSpecies[s]);
        k.defs.push_back({names.back().c_str(), REQ_STATE, RSCM_B200_SCALAR});
    }
    k.defs.push_back({"Forcing|Halocarbons", REQ_OUTPUT, RSCM_B200_SCALAR});
    k.defs.push_back({"Forcing|F-gases", REQ_OUTPUT, RSCM_B200_SCALAR});
    k.defs.push_back({"Forcing|Montreal Gases", REQ_OUTPUT, RSCM_B200_SCALAR});
    k.defs.push_back({"EESC", REQ_OUTPUT, RSCM_B200_SCALAR});
    k.param_names = {"br_multiplier", "cfc11_release_normalisation", "eesc_delay", "air_molar_mass", "atmospheric_mass_tg", "mixing_box_fraction"};
    static const char *const fields[7] = {"lifetime", "radiative_efficiency", "concentration_pi", "molecular_weight", "n_cl", "n_br", "fractional_release"};
    for (int s = 0; s < kHaloNS; ++s)
        for (int f = 0; f < 7; ++f) {
            names.push_back(std::string(fields[f]) + "[" + kHaloSpecies[s] + "]");
            k.param_names.push_back(names.back().c_str());
        }
    k.n_derived = 0;
    k.rk_step_param = -2;
    k.bindable.assign(k.param_names.size(), 0);
    k.reg_weight = 48;
    k.n_state = kHaloNS + 1; // last non-NaN concentration per species (Timeseries::latest_value); [41]: species shared among lanes
    k.needs_time = true;
    // every input through InputState::get_global (mode 3)
    for (int i = 0; i < 2 * kHaloNS; ++i) k.in_access.push_back({i, 3});
    k.const_table = &halocarbon_const_table;
    k.no_slots = true;
    k.scatter_out = true;
    return k;
}

static const std::vector<KindInfo> &kinds()
{
    static std::vector<KindInfo> k = {
        {RSCM_B200_TWO_LAYER, "TwoLayer", "two_layer",
         // crates/rscm-two-layer/src/component.rs:147-154
         {{"Effective Radiative Forcing", REQ_INPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_STATE, RSCM_B200_SCALAR},
          {"Deep Ocean Temperature", REQ_STATE, RSCM_B200_SCALAR}},
         {"lambda0", "a", "efficacy", "eta", "heat_capacity_surface", "heat_capacity_deep"},
         5, -1, {1, 1, 1, 1, 1, 1}, 10},
        {RSCM_B200_CARBON_CYCLE, "CarbonCycle", "carbon_cycle",
         // crates/rscm-components/src/components/carbon_cycle.rs:62-72
         {{"Emissions|CO2|Anthropogenic", REQ_INPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|CO2", REQ_STATE, RSCM_B200_SCALAR},
          {"Cumulative Emissions|CO2", REQ_STATE, RSCM_B200_SCALAR},
          {"Cumulative Land Uptake", REQ_STATE, RSCM_B200_SCALAR}},
         {"tau", "conc_pi", "alpha_temperature", "step_size"},
         2, 3, {1, 1, 1, 0}, 10},
        {RSCM_B200_CO2_ERF, "CO2ERF", "co2_erf",
         // crates/rscm-components/src/components/co2_erf.rs:36-44
         {{"Atmospheric Concentration|CO2", REQ_INPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|CO2", REQ_OUTPUT, RSCM_B200_SCALAR}},
         {"erf_2xco2", "conc_pi"},
         2, -2, {1, 1}, 4},
        {RSCM_B200_GHG_FORCING, "GhgForcing", "ghg_forcing",
         // crates/rscm-magicc/src/forcing/ghg.rs (derive block at top of file)
         {{"Atmospheric Concentration|CO2", REQ_INPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|CH4", REQ_INPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|N2O", REQ_INPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|CO2", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|CH4", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|N2O", REQ_OUTPUT, RSCM_B200_SCALAR}},
         {"method", "co2_pi", "ch4_pi", "n2o_pi", "delq2xco2", "ch4_radeff", "n2o_radeff",
          "olbl_co2_a1", "olbl_co2_b1", "olbl_co2_c1", "olbl_co2_d1", "olbl_ch4_a3", "olbl_ch4_b3",
          "olbl_ch4_d3", "olbl_n2o_a2", "olbl_n2o_b2", "olbl_n2o_c2", "olbl_n2o_d2", "adjust_co2",
          "adjust_ch4", "adjust_n2o"},
         3, -2, {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, 40},
        {RSCM_B200_OZONE_FORCING, "OzoneForcing", "ozone_forcing",
         // crates/rscm-magicc/src/forcing/ozone.rs (derive block)
         {{"EESC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|CH4", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|NOx", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|CO", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|NMVOC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|O3|Stratospheric", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|O3|Tropospheric", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|O3|Temperature Feedback", REQ_OUTPUT, RSCM_B200_SCALAR}},
         {"eesc_reference", "strat_o3_scale", "strat_cl_exponent", "trop_radeff", "trop_oz_ch4", "trop_oz_nox", "trop_oz_co",
          "trop_oz_voc", "ch4_pi", "nox_pi", "co_pi", "nmvoc_pi", "temp_feedback_scale"},
         1, -2, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, 16},
        {RSCM_B200_AEROSOL_DIRECT, "AerosolDirect", "aerosol_direct",
         // crates/rscm-magicc/src/forcing/aerosol_direct.rs (derive block); FourBox output
         {{"Emissions|SOx", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|BC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|OC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|NOx", REQ_INPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|Aerosol|Direct", REQ_OUTPUT, RSCM_B200_FOUR_BOX}},
         {"sox_coefficient", "bc_coefficient", "oc_coefficient", "nitrate_coefficient",
          "sox_regional_0", "sox_regional_1", "sox_regional_2", "sox_regional_3",
          "bc_regional_0", "bc_regional_1", "bc_regional_2", "bc_regional_3",
          "oc_regional_0", "oc_regional_1", "oc_regional_2", "oc_regional_3",
          "nitrate_regional_0", "nitrate_regional_1", "nitrate_regional_2", "nitrate_regional_3",
          "sox_pi", "bc_pi", "oc_pi", "nox_pi", "harmonize", "harmonize_year", "harmonize_target"},
         1, -2, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0}, 24},
        {RSCM_B200_AEROSOL_INDIRECT, "AerosolIndirect", "aerosol_indirect",
         // crates/rscm-magicc/src/forcing/aerosol_indirect.rs (derive block)
         {{"Emissions|SOx", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|OC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Effective Radiative Forcing|Aerosol|Indirect", REQ_OUTPUT, RSCM_B200_SCALAR}},
         {"cloud_albedo_coefficient", "reference_burden", "sox_weight", "oc_weight", "sox_pi", "oc_pi", "harmonize",
          "harmonize_year", "harmonize_target"},
         1, -2, {1, 1, 1, 1, 1, 1, 0, 0, 0}, 8},
        {RSCM_B200_CLIMATE_UDEB, "ClimateUDEB", "climate_udeb",
         // crates/rscm-magicc/src/climate/udeb/mod.rs (derive block): inputs, outputs, states
         {{"Effective Radiative Forcing", REQ_INPUT, RSCM_B200_SCALAR},
          {"Heat Uptake", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Ocean Heat Content", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Sea Surface Temperature", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_STATE, RSCM_B200_FOUR_BOX}},
         {"n_layers", "mixed_layer_depth", "layer_thickness", "kappa", "kappa_min", "kappa_dkdt", "w_initial", "w_variable_fraction",
          "w_threshold_temp_nh", "w_threshold_temp_sh", "ecs", "rf_2xco2", "rlo", "feedback_q_sensitivity", "feedback_cumt_sensitivity",
          "feedback_cumt_period", "k_lo", "k_ns", "amplify_ocean_to_land", "nh_land_fraction", "sh_land_fraction", "depth_dependent_area",
          "temp_adjust_alpha", "temp_adjust_gamma", "polar_sinking_ratio", "land_heat_capacity_enabled", "k_lg", "land_hc_eff_thickness",
          "rf_regions_co2_0", "rf_regions_co2_1", "rf_regions_co2_2", "rf_regions_co2_3", "efficacy_apply", "prescribed_efficacy_co2",
          "ocean_temp_profile", "steps_per_year", "max_temperature"},
         1, -2,
         // geometry / switches are per-graph (they size the shared-memory layout and the host-computed tables)
         {0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 1},
         96, /*n_state: 19 scalars + this lane's 25 rows of the ocean column*/ 44, /*n_smem: eliminated off-diagonal of those rows*/ 25, /*scratch_per_T*/ 1, /*needs_time*/ true,
         /*in_access: erf at_start, erf at_end, surface temperature at_start*/ {{0, 1}, {0, 2}, {1, 1}}, &udeb_const_table, &udeb_window_table,
         /*aux_param: n_layers sizes the register rows*/ 0, /*scratch_fixed*/ 0, /*no_slots*/ false, /*lanes: 2 hemispheres x 2 sweep ends*/ 4,
         /*aux_template*/ true, /*n_smem_lanes*/ 0, /*lane_aware*/ true, /*n_xch*/ 28},
        {RSCM_B200_FOUR_BOX_OHU, "FourBoxOceanHeatUptake", "four_box_ohu",
         // crates/rscm-components/src/components/four_box_ocean_heat_uptake.rs
         {{"Effective Radiative Forcing|Aggregated", REQ_INPUT, RSCM_B200_SCALAR}, {"Heat Uptake|Ocean", REQ_OUTPUT, RSCM_B200_FOUR_BOX}},
         {"northern_ocean_ratio", "northern_land_ratio", "southern_ocean_ratio", "southern_land_ratio"},
         1, -2, {1, 1, 1, 1}, 4},
        {RSCM_B200_OCEAN_SURFACE_PP, "OceanSurfacePartialPressure", "ocean_surface_pp",
         // crates/rscm-components/src/components/ocean_carbon_cycle/ocean_surface_partial_pressure.rs
         {{"Sea Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Dissolved Inorganic Carbon", REQ_INPUT, RSCM_B200_SCALAR},
          {"Ocean Surface Partial Pressure|CO2", REQ_OUTPUT, RSCM_B200_SCALAR}},
         {"ospp_preindustrial", "sensitivity_ospp_to_temperature", "sea_surface_temperature_preindustrial", "delta_ospp_offsets_0",
          "delta_ospp_offsets_1", "delta_ospp_offsets_2", "delta_ospp_offsets_3", "delta_ospp_offsets_4", "delta_ospp_coefficients_0",
          "delta_ospp_coefficients_1", "delta_ospp_coefficients_2", "delta_ospp_coefficients_3", "delta_ospp_coefficients_4"},
         1, -2, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, 8},
        {RSCM_B200_CO2_BUDGET, "CO2Budget", "co2_budget",
         // crates/rscm-magicc/src/carbon/budget.rs (derive block)
         {{"Emissions|CO2|Fossil", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|CO2|Land Use", REQ_INPUT, RSCM_B200_SCALAR},
          {"Carbon Flux|Terrestrial", REQ_INPUT, RSCM_B200_SCALAR},
          {"Carbon Flux|Ocean", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|CO2|Net", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Airborne Fraction|CO2", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|CO2", REQ_STATE, RSCM_B200_SCALAR}},
         {"gtc_per_ppm", "co2_pi"},
         1, -2, {1, 1}, 4, 0, 0, 0, /*needs_time*/ true},
        {RSCM_B200_TERRESTRIAL_CARBON, "TerrestrialCarbon", "terrestrial_carbon",
         // crates/rscm-magicc/src/carbon/terrestrial.rs (derive block)
         {{"Atmospheric Concentration|CO2", REQ_INPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|CO2|Land Use", REQ_INPUT, RSCM_B200_SCALAR},
          {"Carbon Flux|Terrestrial", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Carbon Pool|Plant", REQ_STATE, RSCM_B200_SCALAR},
          {"Carbon Pool|Detritus", REQ_STATE, RSCM_B200_SCALAR},
          {"Carbon Pool|Soil", REQ_STATE, RSCM_B200_SCALAR},
          {"Carbon Pool|Humus", REQ_STATE, RSCM_B200_SCALAR}},
         {"npp_pi", "co2_pi", "beta", "npp_temp_sensitivity", "resp_temp_sensitivity", "detritus_temp_sensitivity", "soil_temp_sensitivity",
          "humus_temp_sensitivity", "plant_pool_pi", "detritus_pool_pi", "soil_pool_pi", "humus_pool_pi", "respiration_pi",
          "frac_npp_to_plant", "frac_npp_to_detritus", "frac_plant_to_detritus", "frac_detritus_to_soil", "frac_soil_to_humus",
          "enable_fertilization", "enable_temp_feedback"},
         5, -2, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0}, 24, 0, 0, 0, /*needs_time*/ true},
        {RSCM_B200_CH4_CHEMISTRY, "CH4Chemistry", "ch4_chemistry",
         // crates/rscm-magicc/src/chemistry/ch4.rs (derive block)
         {{"Emissions|CH4", REQ_INPUT, RSCM_B200_SCALAR},
          {"Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|NOx", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|CO", REQ_INPUT, RSCM_B200_SCALAR},
          {"Emissions|NMVOC", REQ_INPUT, RSCM_B200_SCALAR},
          {"Lifetime|CH4", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|CH4", REQ_STATE, RSCM_B200_SCALAR}},
         {"ch4_pi", "natural_emissions", "tau_oh", "tau_soil", "tau_strat", "tau_trop_cl", "ch4_self_feedback", "oh_sensitivity_scale",
          "oh_nox_sensitivity", "oh_co_sensitivity", "oh_nmvoc_sensitivity", "temp_sensitivity", "include_temp_feedback",
          "include_emissions_feedback", "ppb_to_tg", "nox_reference", "co_reference", "nmvoc_reference"},
         1, -2, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1}, 24, /*n_state*/ 1},
        {RSCM_B200_N2O_CHEMISTRY, "N2OChemistry", "n2o_chemistry",
         // crates/rscm-magicc/src/chemistry/n2o.rs (derive block)
         {{"Emissions|N2O", REQ_INPUT, RSCM_B200_SCALAR},
          {"Lifetime|N2O", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Atmospheric Concentration|N2O", REQ_STATE, RSCM_B200_SCALAR}},
         {"n2o_pi", "natural_emissions", "tau_n2o", "lifetime_feedback", "strat_delay", "ppb_to_tg"},
         1, -2, {1, 1, 1, 1, 0, 1}, 16, /*n_state*/ 8, 0, 0, /*needs_time*/ true},
        {RSCM_B200_OCEAN_CARBON, "OceanCarbon", "ocean_carbon",
         // crates/rscm-magicc/src/carbon/ocean.rs (derive block)
         {{"Atmospheric Concentration|CO2", REQ_INPUT, RSCM_B200_SCALAR},
          {"Sea Surface Temperature", REQ_INPUT, RSCM_B200_SCALAR},
          {"Carbon Flux|Ocean", REQ_OUTPUT, RSCM_B200_SCALAR},
          {"Ocean Surface pCO2", REQ_STATE, RSCM_B200_SCALAR},
          {"Cumulative Ocean Uptake", REQ_STATE, RSCM_B200_SCALAR}},
         {"model", "co2_pi", "pco2_pi", "gas_exchange_scale", "gas_exchange_tau", "temp_sensitivity", "irf_scale", "mixed_layer_depth",
          "ocean_surface_area", "sst_pi", "steps_per_year", "max_history_months", "irf_switch_time",
          "irf_early_kind", "irf_early_n", "irf_early_c0", "irf_early_c1", "irf_early_c2", "irf_early_c3", "irf_early_c4", "irf_early_c5",
          "irf_early_c6", "irf_early_c7", "irf_early_t0", "irf_early_t1", "irf_early_t2", "irf_early_t3", "irf_early_t4", "irf_early_t5",
          "irf_early_t6", "irf_early_t7",
          "irf_late_kind", "irf_late_n", "irf_late_c0", "irf_late_c1", "irf_late_c2", "irf_late_c3", "irf_late_c4", "irf_late_c5",
          "irf_late_c6", "irf_late_c7", "irf_late_t0", "irf_late_t1", "irf_late_t2", "irf_late_t3", "irf_late_t4", "irf_late_t5",
          "irf_late_t6", "irf_late_t7",
          "delta_ospp_offsets_0", "delta_ospp_offsets_1", "delta_ospp_offsets_2", "delta_ospp_offsets_3", "delta_ospp_offsets_4",
          "delta_ospp_coefficients_0", "delta_ospp_coefficients_1", "delta_ospp_coefficients_2", "delta_ospp_coefficients_3",
          "delta_ospp_coefficients_4", "enable_temp_feedback"},
         2, -2,
         // the IRF (scale, switch time, both forms), step count and history bound are per-graph: they define the lag table
         {0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0, 0, 0,
          0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0},
         48, /*n_state*/ 2, /*n_smem*/ 0, /*scratch_per_T*/ 16, /*needs_time*/ true, {}, nullptr, &ocean_irf_table, /*aux_param*/ 10,
         /*scratch_fixed: 4 x 16 rows of block-prefix sums before the history (OCEAN_HIST0)*/ 64, /*no_slots*/ false, /*lanes*/ 1, /*aux_template*/ false,
         /*n_smem_lanes: as one CTA-wide region, two staged history tiles (16 words x 128 threads = 2 x 32 months x 32
           members, OCEAN_KT in magicc_boxes.cuh) and their IRF windows (2 words = 2 x 128 lags)*/ 18, /*lane_aware*/ true, /*n_xch: the tiles' two mbarriers*/ 1},
    };
    static const bool extended = (k.push_back(halocarbon_kind()), true);
    (void)extended;
    return k;
}

const KindInfo *kind_info(int kind)
{
    for (const auto &k : kinds())
        if (k.kind == kind) return &k;
    return nullptr;
}

static int grid_regions(int grid)
{
    return grid == RSCM_B200_FOUR_BOX ? 4 : (grid == RSCM_B200_HEMISPHERIC ? 2 : 1);
}

int Graph::find_var(const std::string &name) const
{
    for (size_t i = 0; i < vars.size(); ++i)
        if (vars[i].name == name) return static_cast<int>(i);
    return -1;
}

// ode_solvers Rk4::integrate: n = ceil((x_end - x)/h) steps of constant h, x += h
// each step; then get_last_step asserts y.len() > 1 and |x_last - t_next| < 5e-3.
int rk4_substeps(double t0, double t1, double h)
{
    const double nd = std::ceil((t1 - t0) / h);
    if (!(nd >= 1.0) || nd > 32000.0) return -1;
    const int n = static_cast<int>(nd);
    double x = t0;
    for (int i = 0; i < n; ++i) x = x + h;
    if (!(std::fabs(x - t1) < 5e-3)) return -1;
    return n;
}

int time_index_for(const Graph &g, double time)
{
    char key[64], k2[64];
    std::snprintf(key, sizeof key, "%.6f", time);
    int found = -1;
    for (int i = 0; i < g.T; ++i) {
        std::snprintf(k2, sizeof k2, "%.6f", g.bounds[i]);
        if (!std::strcmp(key, k2)) found = i; // HashMap insert: the later duplicate wins
    }
    return found;
}

static std::string lit(double v)
{
    char b[64];
    std::snprintf(b, sizeof b, "R(%.17g)", v);
    return b;
}

// ---------------------------------------------------------------------------
// emit: the component graph as a device `Prog` body
// ---------------------------------------------------------------------------
static std::string cell_ref(bool at_end, int cell)
{
    std::ostringstream s;
    s << (at_end ? "nxt[" : "cur[") << cell << "]";
    return s.str();
}

// expression reading input `i` (region r of the component's view) of node n
static std::string input_expr(const Graph &g, const Node &n, int i, int r, bool at_end, std::ostringstream &pre,
                              int &tmp_id)
{
    const Variable &var = g.vars[n.in_var[i]];
    const int want = n.in_grid[i];
    std::string e;
    if (want == var.grid) {
        e = cell_ref(at_end, var.cell0 + r);
    } else if (var.grid == RSCM_B200_FOUR_BOX && want == RSCM_B200_SCALAR) {
        // AggregatingFourBoxWindow::aggregate — state/aggregating.rs:162-177
        const int id = tmp_id++;
        pre << "        const R rv" << id << "[4] = {" << cell_ref(at_end, var.cell0) << ", " << cell_ref(at_end, var.cell0 + 1)
            << ", " << cell_ref(at_end, var.cell0 + 2) << ", " << cell_ref(at_end, var.cell0 + 3) << "};\n";
        pre << "        const R rw" << id << "[4] = {" << lit(g.w_fourbox[0]) << ", " << lit(g.w_fourbox[1]) << ", "
            << lit(g.w_fourbox[2]) << ", " << lit(g.w_fourbox[3]) << "};\n";
        std::ostringstream s;
        s << "rscm_dev::" << (g.custom_w_fourbox ? "read_weighted" : "read_plain") << "<R, 4>(rv" << id << ", rw" << id << ")";
        e = s.str();
    } else if (var.grid == RSCM_B200_FOUR_BOX && want == RSCM_B200_HEMISPHERIC) {
        // state/aggregating.rs:611-618
        std::ostringstream s;
        s << "((" << cell_ref(at_end, var.cell0 + 2 * r) << " + " << cell_ref(at_end, var.cell0 + 2 * r + 1) << ") / R(2))";
        e = s.str();
    } else { // Hemispheric -> Scalar
        const int id = tmp_id++;
        pre << "        const R rv" << id << "[2] = {" << cell_ref(at_end, var.cell0) << ", " << cell_ref(at_end, var.cell0 + 1) << "};\n";
        pre << "        const R rw" << id << "[2] = {" << lit(g.w_hemi[0]) << ", " << lit(g.w_hemi[1]) << "};\n";
        std::ostringstream s;
        s << "rscm_dev::" << (g.custom_w_hemi ? "read_weighted" : "read_plain") << "<R, 2>(rv" << id << ", rw" << id << ")";
        e = s.str();
    }
    if (n.in_factor[i] != 1.0) e = "(" + e + " * " + lit(n.in_factor[i]) + ")";
    return e;
}

static void emit_program(Graph &g)
{
    std::ostringstream o;
    o << "    static constexpr int NC = " << g.n_cells << ";\n";
    o << "    static constexpr int NP = " << g.n_slots << ";\n";
    o << "    static constexpr int ND = " << g.n_derived << ";\n";
    int weight = 0;
    for (const Node &n : g.nodes)
        if (n.kind != KIND_AGGREGATOR) weight += kind_info(n.kind)->reg_weight;
    // 8 CTAs of 128 threads per SM = 64 registers per thread; register-hungry programs get 4 (128 registers)
    // programs with per-thread shared-memory scratch are limited by shared memory, not registers (2 = no register cap;
    // two CTAs fit one SM in fp32)
    // Programs whose members span several lanes (ClimateUDEB) are bound by latency — dependent chains, two rendezvous per
    // sub-step — with a few warps per scheduler, so the CTAs per SM matter more than a spill-free register allocation:
    // 4 CTAs (128 registers, about 200 B of spills per thread) where the shared memory of four fits one SM, measured
    // 3.06e8 against 2.66e8 member-years/s with 3 CTAs (168 registers) on config 4.  The exogenous rows then stay in
    // global memory (L2; 28 KB per CTA at config 4 — staging them is worth 3 % only when it costs no CTA).
    int lanes_blocks = 3;
    const int tpad = (g.T + 3) / 4 * 4;
    const long lanes_smem = 16 + 1024 /*static + reserved per CTA*/ + (g.needs_time ? (tpad + 4) * 8L : 0) + static_cast<long>(g.ctab.size()) * 8 +
                            static_cast<long>(g.n_rk) * tpad * 4 + g.n_smem * 1024L + g.n_xch * 256L;
    const long exo_smem = static_cast<long>(g.n_exo_rows) * tpad * 8;
    if (g.lanes > 1 && lanes_smem <= 232448 / 4) lanes_blocks = 4;
    if (const char *e = std::getenv("RSCM_B200_LANES_MIN_BLOCKS")) lanes_blocks = std::max(1, std::min(8, std::atoi(e))); // tuning knob
    o << "    static constexpr int MIN_BLOCKS = " << (g.lanes > 1 ? lanes_blocks : (g.n_smem > 0 ? 2 : (weight <= 32 ? 8 : 4))) << ";\n";
    o << "    static constexpr int NS = " << g.n_state << ";\n";
    o << "    static constexpr int NSM = " << g.n_smem << ";\n";
    o << "    static constexpr int LANES = " << g.lanes << ";\n";
    o << "    static constexpr int NXCH = " << g.n_xch << ";\n";
    // more cells than a thread can keep in registers next to its working set: they live in local memory, and the kernel
    // swaps the two time levels instead of copying them (kernel.cuh)
    o << "    static constexpr bool SWAP_CELLS = " << (g.n_cells > 64 ? "true" : "false") << ";\n";
    o << "    static constexpr bool SYNC_STEPS = " << (g.n_cells > 64 ? "true" : "false") << ";\n";
    o << "    static constexpr bool NEEDS_TIME = " << (g.needs_time ? "true" : "false") << ";\n";
    // exogenous rows are staged into shared memory unless per-thread scratch or a long row list needs the space
    g.stage_exo = (g.n_smem == 0 || g.lanes > 1) && g.n_exo_rows <= 24;
    if (g.lanes > 1 && lanes_smem + exo_smem > 232448 / lanes_blocks) g.stage_exo = false; // not at the price of a CTA per SM
    if (g.lanes > 1 && std::getenv("RSCM_B200_LANES_STAGE_EXO")) g.stage_exo = std::atoi(std::getenv("RSCM_B200_LANES_STAGE_EXO")) != 0; // tuning knob
    o << "    static constexpr bool STAGE_EXO = " << (g.stage_exo ? "true" : "false") << ";\n";
    o << "    __host__ __device__ static constexpr int exo_row(int c) { return ";
    for (int c = 0; c < g.n_cells; ++c) {
        const Variable &v = g.vars[g.cell_var[c]];
        if (v.exo_row0 >= 0) o << "c == " << c << " ? " << (v.exo_row0 + g.cell_region[c]) << " : ";
    }
    o << "-1; }\n";
    o << "    __host__ __device__ static constexpr bool endogenous(int c) { return ";
    for (int c = 0; c < g.n_cells; ++c)
        if (g.vars[g.cell_var[c]].endogenous) o << "c == " << c << " || ";
    o << "false; }\n";
    // Cells start every step as NaN ("not written yet", timeseries.rs:334-345).  A cell whose single producer runs
    // before every node that reads it is overwritten on all paths (a failing component writes NaN explicitly below), so
    // its blanket NaN store can be dropped; everything else keeps it.
    std::vector<int> n_writers(g.vars.size(), 0), writer_pos(g.vars.size(), -1), first_reader_pos(g.vars.size(), 1 << 30);
    for (size_t pos = 0; pos < g.order.size(); ++pos) {
        const Node &n = g.nodes[g.order[pos]];
        for (int v : n.in_var) first_reader_pos[v] = std::min(first_reader_pos[v], static_cast<int>(pos));
        for (int v : n.out_var) { ++n_writers[v]; writer_pos[v] = static_cast<int>(pos); }
    }
    o << "    __host__ __device__ static constexpr bool nan_init(int c) { return ";
    for (int c = 0; c < g.n_cells; ++c) {
        const int v = g.cell_var[c];
        const bool safe = n_writers[v] == 1 && writer_pos[v] <= first_reader_pos[v];
        if (!safe) o << "c == " << c << " || ";
    }
    o << "false; }\n";
    o << "    __host__ __device__ static constexpr int regions(int c) { return ";
    for (int c = 0; c < g.n_cells; ++c) {
        const int r = g.vars[g.cell_var[c]].n_regions;
        if (r != 1) o << "c == " << c << " ? " << r << " : ";
    }
    o << "1; }\n";
    o << "    template <class R> __device__ __forceinline__ static void prepare(const R *P, R *D) {\n";
    o << "        (void)P; (void)D;\n";
    for (size_t ni = 0; ni < g.nodes.size(); ++ni) {
        const Node &n = g.nodes[ni];
        if (n.kind == KIND_AGGREGATOR) continue;
        const KindInfo *k = kind_info(n.kind);
        o << "        rscm_dev::" << k->dev_name << "_prepare<R>(P + " << n.param_base << ", D + " << n.derived_base << ");\n";
    }
    o << "    }\n";
    auto node_ref = [](const Node &n) {
        std::ostringstream s;
        s << "rscm_dev::NodeRef{" << (n.rk_table >= 0 ? n.rk_table : 0) << ", " << n.ctab_base << ", " << n.smem_base << ", "
          << n.scratch_base << ", " << n.gtab_base << ", " << n.aux << ", " << n.xch_user << "}";
        return s.str();
    };
    o << "    template <class R> __device__ __forceinline__ static void init_state(const R *P, const R *D, R *S,\n"
         "                                                                        const rscm_dev::StepCtx<R> &cx) {\n";
    o << "        (void)P; (void)D; (void)S; (void)cx;\n";
    for (const Node &n : g.nodes) {
        if (n.kind == KIND_AGGREGATOR) continue;
        const KindInfo *k = kind_info(n.kind);
        if (k->n_state == 0 && k->n_smem == 0) continue;
        o << "        rscm_dev::" << k->dev_name << "_init_state<R" << (k->aux_template ? ", " + std::to_string(n.aux) : std::string()) << ">(P + " << n.param_base << ", D + " << n.derived_base << ", S + "
          << n.state_base << ", cx, " << node_ref(n) << ");\n";
    }
    o << "    }\n";
    o << "    template <class R> __device__ __forceinline__ static void step(const R *P, const R *D, const R *cur, R *nxt, R *S,\n"
         "                                                                  const rscm_dev::StepCtx<R> &cx, unsigned &fail) {\n";
    o << "        (void)P; (void)D; (void)cur; (void)S; (void)cx; (void)fail;\n";
    int tmp_id = 0;
    // lane-group programs: role 0 (warp 0) runs the graph; lane nodes are entered by every role
    const bool lanes = g.lanes > 1;
    // RSCM_B200_NODE_CLOCKS=1 (profiling aid, tools/node_clocks.py): thread 0 of CTA 0 accumulates the cycles it spends in each
    // node and prints them after the last step.  The program text changes, so these builds have their own cache entries.
    const bool node_clocks = std::getenv("RSCM_B200_NODE_CLOCKS") != nullptr;
    if (node_clocks)
        o << "        __shared__ long long node_clk[" << g.nodes.size() << "];\n        const bool clk_on = threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;\n"
             "        if (clk_on && cx.N == 0) for (int i = 0; i < " << g.nodes.size() << "; ++i) node_clk[i] = 0;\n        long long clk_t = 0;\n";
    for (int ni : g.order) {
        const Node &n = g.nodes[ni];
        const bool lane_node = lanes && n.kind != KIND_AGGREGATOR && n.lane_node;
        if (node_clocks) o << "      clk_t = clock64();\n";
        o << "      " << (lanes && !lane_node ? "if (cx.role == 0) " : "") << "{ // node " << ni << "\n";
        std::ostringstream pre;
        if (n.kind == KIND_AGGREGATOR) {
            // AggregatorComponent::solve — schema.rs:874-951: contributors read at_end
            const int R = grid_regions(n.agg_grid);
            const int nc = static_cast<int>(n.in_var.size());
            std::ostringstream body;
            for (int r = 0; r < R; ++r) {
                body << "        { const R av[" << nc << "] = {";
                for (int i = 0; i < nc; ++i) body << (i ? ", " : "") << input_expr(g, n, i, r, true, pre, tmp_id);
                body << "};\n";
                const int oc = g.vars[n.out_var[0]].cell0 + r;
                if (n.agg_op == RSCM_B200_AGG_SUM) body << "          nxt[" << oc << "] = rscm_dev::agg_sum<R, " << nc << ">(av); }\n";
                else if (n.agg_op == RSCM_B200_AGG_MEAN) body << "          nxt[" << oc << "] = rscm_dev::agg_mean<R, " << nc << ">(av); }\n";
                else {
                    body << "          const R aw[" << nc << "] = {";
                    for (int i = 0; i < nc; ++i) body << (i ? ", " : "") << lit(n.agg_w[i]);
                    body << "};\n          nxt[" << oc << "] = rscm_dev::agg_weighted<R, " << nc << ">(av, aw); }\n";
                }
            }
            o << pre.str() << body.str();
        } else {
            const KindInfo *k = kind_info(n.kind);
            std::vector<std::string> in_exprs;
            // default access: one get() per input — UpstreamOutput -> at_end (always in range during run), else
            // at_start (state/windows.rs:229-234); state inputs are read with at_start.  Kinds that call at_start /
            // at_end explicitly (e.g. ClimateUDEB's erf_start / erf_end) list their accesses in in_access.
            std::vector<std::pair<int, int>> access = k->in_access;
            if (access.empty())
                for (size_t i = 0; i < n.in_var.size(); ++i) access.push_back({static_cast<int>(i), 0});
            for (const auto &ac : access) {
                const int i = ac.first;
                const int Rc = grid_regions(n.in_grid[i]);
                const bool at_end = ac.second == 2 || (ac.second == 0 && n.in_src[i] == RSCM_B200_SRC_UPSTREAM);
                if (ac.second == 3 && n.in_src[i] == RSCM_B200_SRC_UPSTREAM) {
                    // InputState::get_global of an endogenous series = Timeseries::latest_value (state/mod.rs:231-254):
                    // the value written this step if it is not NaN, else the one at the current index.  (Own states
                    // keep their last non-NaN value in the component's S[]; exogenous series are read at the
                    // current time, NaN included.)
                    for (int r = 0; r < Rc; ++r) {
                        const std::string e = input_expr(g, n, i, r, true, pre, tmp_id), b = input_expr(g, n, i, r, false, pre, tmp_id);
                        const int id = tmp_id++;
                        pre << "        const R lv" << id << " = " << e << ";\n";
                        in_exprs.push_back("(lv" + std::to_string(id) + " == lv" + std::to_string(id) + " ? lv" + std::to_string(id) + " : " + b + ")");
                    }
                    continue;
                }
                for (int r = 0; r < Rc; ++r) in_exprs.push_back(input_expr(g, n, i, r, at_end, pre, tmp_id));
            }
            int n_out_vals = 0;
            for (size_t i = 0; i < n.out_var.size(); ++i) n_out_vals += grid_regions(n.out_grid[i]);
            const std::string solve_call = std::string("rscm_dev::") + k->dev_name + "_solve<R" + (k->aux_template ? ", " + std::to_string(n.aux) : std::string()) +
                                           ">(P + " + std::to_string(n.param_base) + ", D + " + std::to_string(n.derived_base) + ", in, out, cx, S + " +
                                           std::to_string(n.state_base) + ", " + node_ref(n) + ")";
            if (lane_node) {
                // role 0 evaluates the inputs from its cells and publishes them; the other roles pick them up
                const size_t nin = in_exprs.empty() ? 1 : in_exprs.size();
                o << "        R in[" << nin << "];\n";
                o << "        if (cx.role == 0) {\n" << pre.str();
                for (size_t i = 0; i < in_exprs.size(); ++i) {
                    o << "          in[" << i << "] = " << in_exprs[i] << ";\n";
                    o << "          cx.xch[" << (n.xch_base + static_cast<int>(i)) * 32 << "] = static_cast<double>(in[" << i << "]);\n";
                }
                if (in_exprs.empty()) o << "          in[0] = R(0);\n";
                o << "        }\n        __syncthreads();\n";
                o << "        if (cx.role != 0) {\n";
                for (size_t i = 0; i < in_exprs.size(); ++i)
                    o << "          in[" << i << "] = R(cx.xch[" << (n.xch_base + static_cast<int>(i)) * 32 << "]);\n";
                if (in_exprs.empty()) o << "          in[0] = R(0);\n";
                o << "        }\n";
                o << "        R out[" << n_out_vals << "];\n";
                o << "        const bool ok = " << solve_call << ";\n";
                o << "        __syncthreads(); // every role has left the node before role 0 reuses the exchange slots\n";
                o << "        if (cx.role == 0) {\n        if (ok) {\n";
            } else {
            o << pre.str();
            // inputs that are a run of consecutive current-level cells (HalocarbonChemistry's 82 in a schema that lists a
            // species' emissions and concentration together) are read in place: no copy of cells that live in local memory
            int run0 = -1;
            if (in_exprs.size() >= 8 && in_exprs[0].rfind("cur[", 0) == 0) {
                run0 = std::atoi(in_exprs[0].c_str() + 4);
                for (size_t i = 0; i < in_exprs.size() && run0 >= 0; ++i)
                    if (in_exprs[i] != "cur[" + std::to_string(run0 + static_cast<int>(i)) + "]") run0 = -1;
            }
            if (run0 >= 0) {
                o << "        const R *in = cur + " << run0 << ";\n";
            } else {
                o << "        const R in[" << (in_exprs.empty() ? 1 : in_exprs.size()) << "] = {";
                for (size_t i = 0; i < in_exprs.size(); ++i) o << (i ? ", " : "") << in_exprs[i];
                if (in_exprs.empty()) o << "R(0)";
                o << "};\n";
            }
            bool scatter = k->scatter_out;
            for (size_t i = 0; i < n.out_var.size(); ++i) scatter = scatter && n.out_grid[i] == g.vars[n.out_var[i]].grid;
            if (scatter) {
                std::vector<int> oc;
                for (size_t i = 0; i < n.out_var.size(); ++i)
                    for (int r = 0; r < grid_regions(n.out_grid[i]); ++r) oc.push_back(g.vars[n.out_var[i]].cell0 + r);
                // the longest arithmetic run of cells from the first output on
                int lin_n = 1;
                const int lin_b = oc.size() > 1 ? oc[1] - oc[0] : 0;
                while (lin_n < static_cast<int>(oc.size()) && oc[lin_n] == oc[0] + lin_b * lin_n) ++lin_n;
                o << "        constexpr short out_cell[" << n_out_vals << "] = {";
                for (size_t i = 0; i < oc.size(); ++i) o << (i ? ", " : "") << oc[i];
                o << "};\n        const rscm_dev::ScatterOut<R, " << lin_n << ", " << oc[0] << ", " << lin_b << "> out{nxt, out_cell};\n";
                o << "        if (!" << solve_call << ") {\n            fail |= 1u;\n";
                for (size_t i = 0; i < n.out_var.size(); ++i) {
                    const Variable &var = g.vars[n.out_var[i]];
                    for (int r = 0; r < var.n_regions; ++r) o << "            nxt[" << (var.cell0 + r) << "] = rscm_dev::r_nan<R>();\n";
                }
                o << "        }\n      }\n";
                if (node_clocks) o << "      if (clk_on) node_clk[" << ni << "] += clock64() - clk_t;\n";
                continue;
            }
            o << "        R out[" << n_out_vals << "];\n";
            o << "        if (" << solve_call << ") {\n";
            }
            int pos = 0;
            for (size_t i = 0; i < n.out_var.size(); ++i) {
                const Variable &var = g.vars[n.out_var[i]];
                const int Rc = grid_regions(n.out_grid[i]);
                if (n.out_grid[i] == var.grid) {
                    for (int r = 0; r < Rc; ++r) o << "            nxt[" << (var.cell0 + r) << "] = out[" << (pos + r) << "];\n";
                } else if (n.out_grid[i] == RSCM_B200_FOUR_BOX && var.grid == RSCM_B200_SCALAR) {
                    // write-side aggregation — model/transformations.rs:31-128
                    o << "            nxt[" << var.cell0 << "] = out[" << pos << "] * " << lit(g.w_fourbox[0]) << " + out[" << pos + 1
                      << "] * " << lit(g.w_fourbox[1]) << " + out[" << pos + 2 << "] * " << lit(g.w_fourbox[2]) << " + out["
                      << pos + 3 << "] * " << lit(g.w_fourbox[3]) << ";\n";
                } else if (n.out_grid[i] == RSCM_B200_FOUR_BOX && var.grid == RSCM_B200_HEMISPHERIC) {
                    const double wn = g.w_fourbox[0] + g.w_fourbox[1], ws = g.w_fourbox[2] + g.w_fourbox[3];
                    o << "            nxt[" << var.cell0 << "] = (out[" << pos << "] * " << lit(g.w_fourbox[0]) << " + out[" << pos + 1
                      << "] * " << lit(g.w_fourbox[1]) << ") / " << lit(wn) << ";\n";
                    o << "            nxt[" << var.cell0 + 1 << "] = (out[" << pos + 2 << "] * " << lit(g.w_fourbox[2]) << " + out["
                      << pos + 3 << "] * " << lit(g.w_fourbox[3]) << ") / " << lit(ws) << ";\n";
                } else { // Hemispheric -> Scalar
                    o << "            nxt[" << var.cell0 << "] = out[" << pos << "] * " << lit(g.w_hemi[0]) << " + out[" << pos + 1
                      << "] * " << lit(g.w_hemi[1]) << ";\n";
                }
                pos += Rc;
            }
            o << "        } else {\n            fail |= 1u;\n";
            for (size_t i = 0; i < n.out_var.size(); ++i) {
                const Variable &var = g.vars[n.out_var[i]];
                for (int r = 0; r < var.n_regions; ++r) o << "            nxt[" << (var.cell0 + r) << "] = rscm_dev::r_nan<R>();\n";
            }
            o << "        }\n";
            if (lane_node) o << "        }\n";
        }
        o << "      }\n";
        if (node_clocks) o << "      if (clk_on) node_clk[" << ni << "] += clock64() - clk_t;\n";
    }
    if (node_clocks) {
        o << "      if (clk_on && cx.N == " << g.T - 2 << ") {\n";
        for (int ni : g.order)
            o << "        printf(\"node_clocks " << ni << " " << (g.nodes[ni].kind == KIND_AGGREGATOR ? "Aggregator" : kind_info(g.nodes[ni].kind)->dev_name)
              << " %lld\\n\", node_clk[" << ni << "] / " << g.T - 1 << ");\n";
        o << "      }\n";
    }
    o << "    }\n";
    g.program_source = o.str();
    g.signature = g.program_source;
}

// ---------------------------------------------------------------------------
// compile_graph — ModelBuilder::build restated (builder.rs:418-860)
// ---------------------------------------------------------------------------
bool compile_graph(const rscm_b200_graph_desc &d, Graph &g, std::string &err)
{
    if (d.abi_version != RSCM_B200_ABI_VERSION) { err = "ABI version mismatch"; return false; }
    if (d.n_times < 2 || !d.time_bounds) { err = "time axis with at least 2 points required"; return false; }
    if (d.n_components < 1 || !d.components) { err = "at least one component required"; return false; }
    g = Graph();
    g.T = d.n_times;
    g.bounds.assign(d.time_bounds, d.time_bounds + d.n_times + 1);
    if (d.four_box_weights) { std::memcpy(g.w_fourbox, d.four_box_weights, sizeof g.w_fourbox); g.custom_w_fourbox = true; }
    if (d.hemispheric_weights) { std::memcpy(g.w_hemi, d.hemispheric_weights, sizeof g.w_hemi); g.custom_w_hemi = true; }

    std::vector<std::string> agg_names;
    for (int i = 0; i < d.n_aggregates; ++i) agg_names.push_back(d.aggregates[i].name);
    auto is_agg = [&](const std::string &n) {
        for (auto &a : agg_names) if (a == n) return true;
        return false;
    };
    if (d.n_aggregates > 0 && !d.has_schema) { err = "aggregates require a schema"; return false; }

    bool grid_mismatch = false;
    std::string mismatch_msg;
    auto add_var = [&](const char *name, int grid, int req) {
        int v = g.find_var(name);
        if (v >= 0) {
            // The first definition wins; without a schema a second definition on another grid is an error
            // (verify_definition, model/validation.rs:30-107: GridTypeMismatch unless has_schema — the schema's relaxed
            // rules are what allow the read/write aggregation transforms).
            if (!d.has_schema && g.vars[v].grid != grid && !grid_mismatch) {
                static const char *const GN[] = {"Scalar", "FourBox", "Hemispheric"};
                grid_mismatch = true;
                mismatch_msg = std::string("grid type mismatch for variable '") + name + "': defined on " + GN[g.vars[v].grid] +
                               ", redefined on " + GN[grid] + " (a VariableSchema is required for grid aggregation)";
            }
            return v;
        }
        Variable var;
        var.name = name;
        var.grid = grid;
        var.req = req;
        g.vars.push_back(var);
        return static_cast<int>(g.vars.size()) - 1;
    };

    // graph node ids: 0 = NullComponent root, node i -> i + 1
    std::vector<std::pair<int, int>> edges;
    std::map<int, int> producer; // `endogenous`: var -> graph node id
    std::vector<std::pair<int, int>> pending; // (graph node, var) aggregate deps (builder.rs:441)

    for (int ci = 0; ci < d.n_components; ++ci) {
        const rscm_b200_component_desc &cd = d.components[ci];
        const KindInfo *k = kind_info(cd.kind);
        if (!k) {
            err = "component kind " + std::to_string(cd.kind) + " has no device implementation (no CPU fallback)";
            return false;
        }
        if (k->bindable.size() != k->param_names.size()) {
            err = std::string("internal: descriptor table of ") + k->type_name + " is inconsistent";
            return false;
        }
        if (cd.n_params != static_cast<int>(k->param_names.size()) || !cd.params) {
            err = std::string("wrong parameter count for ") + k->type_name;
            return false;
        }
        Node n;
        n.kind = cd.kind;
        n.params.assign(cd.params, cd.params + cd.n_params);
        const int gnode = ci + 1;
        bool has_dep = false;
        // requires: classification (builder.rs:465-488) and edges (:490-519), inputs() order
        for (const VarDef &def : k->defs) {
            if (def.req == REQ_OUTPUT) continue;
            const int v = add_var(def.name, def.grid, def.req);
            int src;
            if (def.req == REQ_STATE) src = RSCM_B200_SRC_OWN_STATE;
            else if (producer.count(v)) src = RSCM_B200_SRC_UPSTREAM;
            else if (is_agg(def.name)) src = RSCM_B200_SRC_UPSTREAM;
            else src = RSCM_B200_SRC_EXOGENOUS;
            n.in_var.push_back(v);
            n.in_src.push_back(src);
            n.in_grid.push_back(def.grid);
            n.in_factor.push_back(1.0);
            if (producer.count(v)) { edges.push_back({producer[v], gnode}); has_dep = true; }
            else if (is_agg(def.name)) { pending.push_back({gnode, v}); has_dep = true; }
            else g.vars[v].in_exogenous_list = true;
        }
        if (!has_dep) edges.push_back({0, gnode}); // builder.rs:521-531
        // provides (builder.rs:533-560), outputs() order
        for (const VarDef &def : k->defs) {
            if (def.req == REQ_INPUT) continue;
            const int v = add_var(def.name, def.grid, def.req);
            n.out_var.push_back(v);
            n.out_grid.push_back(def.grid);
            if (producer.count(v)) edges.push_back({producer[v], gnode});
            producer[v] = gnode;
        }
        g.nodes.push_back(n);
    }
    g.n_user = d.n_components;
    if (grid_mismatch) { err = mismatch_msg; return false; }

    for (int i = 0; i < d.n_unit_factors; ++i) {
        const rscm_b200_unit_factor &uf = d.unit_factors[i];
        if (uf.component < 0 || uf.component >= g.n_user) { err = "unit factor: bad component index"; return false; }
        const int v = g.find_var(uf.variable);
        Node &n = g.nodes[uf.component];
        for (size_t j = 0; j < n.in_var.size(); ++j)
            if (n.in_var[j] == v) n.in_factor[j] = uf.factor;
    }

    if (d.has_schema) {
        // schema is the source of truth for the storage grid (builder.rs:593-630)
        for (int s = 0; s < d.n_schema_variables; ++s) {
            const rscm_b200_schema_variable &sv = d.schema_variables[s];
            int v = g.find_var(sv.name);
            if (v < 0) {
                v = add_var(sv.name, sv.grid, REQ_INPUT);
                g.vars[v].in_exogenous_list = true;
            } else if (g.vars[v].grid != sv.grid) {
                g.vars[v].grid = sv.grid;
                if (!producer.count(v)) g.vars[v].in_exogenous_list = true;
            }
        }
        // aggregators (builder.rs:631-700)
        for (int ai = 0; ai < d.n_aggregates; ++ai) {
            const rscm_b200_aggregate_desc &ad = d.aggregates[ai];
            Node n;
            n.kind = KIND_AGGREGATOR;
            n.agg_name = ad.name;
            n.agg_op = ad.op;
            n.agg_grid = ad.grid;
            const int gnode = static_cast<int>(g.nodes.size()) + 1;
            bool has_dep = false;
            if (ad.n_contributors < 1) { err = "aggregate without contributors"; return false; }
            if (ad.op == RSCM_B200_AGG_WEIGHTED && !ad.weights) { err = "weighted aggregate without weights"; return false; }
            for (int i = 0; i < ad.n_contributors; ++i) {
                int v = g.find_var(ad.contributors[i]);
                if (v < 0) {
                    v = add_var(ad.contributors[i], ad.grid, REQ_INPUT);
                    g.vars[v].in_exogenous_list = true;
                }
                n.agg_contrib.push_back(ad.contributors[i]);
                n.agg_w.push_back(ad.weights ? ad.weights[i] : 1.0);
                n.in_var.push_back(v);
                n.in_src.push_back(RSCM_B200_SRC_EXOGENOUS); // unused: contributors are read at_end explicitly
                n.in_grid.push_back(ad.grid);
                n.in_factor.push_back(1.0);
                if (producer.count(v)) { edges.push_back({producer[v], gnode}); has_dep = true; }
            }
            if (!has_dep) edges.push_back({0, gnode});
            const int v = add_var(ad.name, ad.grid, REQ_OUTPUT);
            g.vars[v].grid = ad.grid;
            n.out_var.push_back(v);
            n.out_grid.push_back(ad.grid);
            producer[v] = gnode;
            g.nodes.push_back(n);
        }
        for (auto &p : pending)
            if (producer.count(p.second)) edges.push_back({producer[p.second], p.first});
    }

    // supported transforms only run fine -> coarse
    for (const Node &n : g.nodes) {
        for (size_t i = 0; i < n.in_var.size(); ++i) {
            const int sg = g.vars[n.in_var[i]].grid, wg = n.in_grid[i];
            if (sg == wg || sg == RSCM_B200_FOUR_BOX || (sg == RSCM_B200_HEMISPHERIC && wg == RSCM_B200_SCALAR)) continue;
            err = "variable '" + g.vars[n.in_var[i]].name + "': read transform coarse -> fine is not defined";
            return false;
        }
        for (size_t i = 0; i < n.out_var.size(); ++i) {
            const int sg = g.vars[n.out_var[i]].grid, cg = n.out_grid[i];
            if (sg == cg || cg == RSCM_B200_FOUR_BOX || (cg == RSCM_B200_HEMISPHERIC && sg == RSCM_B200_SCALAR)) continue;
            err = "variable '" + g.vars[n.out_var[i]].name + "': write transform coarse -> fine is not defined";
            return false;
        }
    }

    // variables: initial values, endogenous flag, cells, staged rows
    for (size_t v = 0; v < g.vars.size(); ++v) {
        Variable &var = g.vars[v];
        var.n_regions = grid_regions(var.grid);
        var.endogenous = producer.count(static_cast<int>(v)) > 0;
        for (int i = 0; i < d.n_initial_values; ++i)
            if (var.name == d.initial_values[i].name) { var.has_initial = true; var.initial = d.initial_values[i].value; }
        if (var.req == REQ_STATE && !var.has_initial) { // builder.rs:703-716
            err = "missing initial value for state variable '" + var.name + "'";
            return false;
        }
        var.cell0 = g.n_cells;
        for (int r = 0; r < var.n_regions; ++r) { g.cell_var.push_back(static_cast<int>(v)); g.cell_region.push_back(r); }
        g.n_cells += var.n_regions;
        if (!var.endogenous) {
            var.exo_index = static_cast<int>(g.exo_vars.size());
            var.exo_row0 = g.n_exo_rows;
            g.exo_vars.push_back(static_cast<int>(v));
            g.n_exo_rows += var.n_regions;
        }
    }

    // execution order: petgraph Bfs from the root; Graph::neighbors walks outgoing
    // edges most-recently-added first (model/runtime.rs:504-510)
    {
        const int total = static_cast<int>(g.nodes.size()) + 1;
        std::vector<char> seen(total, 0);
        std::deque<int> q;
        q.push_back(0);
        seen[0] = 1;
        while (!q.empty()) {
            const int u = q.front();
            q.pop_front();
            if (u != 0) g.order.push_back(u - 1);
            for (int e = static_cast<int>(edges.size()) - 1; e >= 0; --e) {
                if (edges[e].first != u) continue;
                const int w = edges[e].second;
                if (!seen[w]) { seen[w] = 1; q.push_back(w); }
            }
        }
        // cycle check (builder.rs:563)
        std::vector<int> indeg(total, 0);
        std::vector<char> removed(total, 0);
        for (auto &e : edges) indeg[e.second]++;
        int done = 0;
        for (bool progressed = true; progressed;) {
            progressed = false;
            for (int u = 0; u < total; ++u) {
                if (removed[u] || indeg[u]) continue;
                removed[u] = 1; ++done; progressed = true;
                for (auto &e : edges) if (e.first == u) indeg[e.second]--;
            }
        }
        if (done != total) { err = "component graph contains a cycle"; return false; }
        // The device program executes nodes in exactly this BFS order with `nxt` cells
        // NaN until written, so a consumer that the BFS schedules before its producer
        // reads NaN at N+1 just as it does in the reference (no topological re-sort).
    }

    // lanes per member of the program = the largest of its kinds (decides which kinds get lane-group scratch below)
    for (const Node &n : g.nodes)
        if (n.kind != KIND_AGGREGATOR) g.lanes = std::max(g.lanes, kind_info(n.kind)->lanes);

    // parameter / derived slots, RK4 tables
    for (size_t ni = 0; ni < g.nodes.size(); ++ni) {
        Node &n = g.nodes[ni];
        if (n.kind == KIND_AGGREGATOR) continue;
        const KindInfo *k = kind_info(n.kind);
        n.param_base = g.n_slots;
        n.derived_base = g.n_derived;
        if (!k->no_slots) {
            for (size_t p = 0; p < n.params.size(); ++p) {
                g.slot_default.push_back(n.params[p]);
                g.slot_bindable.push_back(k->bindable[p]);
            }
            g.n_slots += static_cast<int>(n.params.size());
        }
        g.n_derived += k->n_derived;
        if (k->aux_param >= 0) n.aux = static_cast<int>(n.params[k->aux_param]);
        // stateful kinds: per-thread state, shared-memory scratch, global scratch, per-graph constant tables
        n.state_base = g.n_state;
        n.smem_base = 0; // per-thread shared-memory words hold nothing between two solves: every node uses the same ones
        n.scratch_base = g.n_scratch_rows;
        n.ctab_base = static_cast<int>(g.ctab.size());
        g.n_state += k->n_state;
        g.n_smem = std::max(g.n_smem, k->n_smem + (g.lanes > 1 ? k->n_smem_lanes : 0));
        if (g.lanes > 1 && (k->lanes > 1 || k->lane_aware)) {
            // a lane node: its input values travel from role 0 to the other roles through the first exchange slots
            n.lane_node = true;
            int n_in_vals = 0;
            std::vector<std::pair<int, int>> access = k->in_access;
            if (access.empty())
                for (size_t i = 0; i < n.in_var.size(); ++i) access.push_back({static_cast<int>(i), 0});
            for (const auto &ac : access) n_in_vals += grid_regions(n.in_grid[ac.first]);
            n.xch_base = g.n_xch;
            n.xch_user = g.n_xch + n_in_vals;
            g.n_xch += n_in_vals + k->n_xch;
        }
        g.n_scratch_rows += k->scratch_fixed + k->scratch_per_T * g.T;
        g.needs_time = g.needs_time || k->needs_time;
        if (k->const_table) {
            std::string terr;
            const std::vector<double> tab = k->const_table(n.params, terr);
            if (!terr.empty()) { err = terr; return false; }
            g.ctab.insert(g.ctab.end(), tab.begin(), tab.end());
            if (n.kind == RSCM_B200_HALOCARBON_CHEMISTRY) {
                // decay factors exp(-dt / lifetime) per species (halocarbon.rs:115-134) for the step length of a uniform
                // time axis: {dt_ref (NaN when the steps differ), 41 factors}.  Lifetimes are per-graph, so on a uniform
                // axis the 41 exponentials per member-year are constants of the graph.
                double dt_ref = g.bounds[1] - g.bounds[0];
                for (int t = 1; t < g.T; ++t)
                    if (g.bounds[t + 1] - g.bounds[t] != dt_ref) dt_ref = std::nan("");
                g.ctab.push_back(dt_ref);
                for (int sp = 0; sp < kHaloNS; ++sp) g.ctab.push_back(std::exp(-dt_ref / tab[6 * sp]));
            }
        }
        if (g.gtab.size() % 2) g.gtab.push_back(0.0); // every node's table starts on a 16-byte boundary (bulk copies of table windows)
        n.gtab_base = static_cast<int>(g.gtab.size());
        if (k->global_table) {
            std::string terr;
            const std::vector<double> tab = k->global_table(n.params, g.T, g.bounds.data(), terr);
            if (!terr.empty()) { err = terr; return false; }
            g.gtab.insert(g.gtab.end(), tab.begin(), tab.end());
        }
        if (n.kind == RSCM_B200_N2O_CHEMISTRY && !(n.params[4] >= 0.0 && n.params[4] <= 6.0)) {
            err = "N2OChemistry: strat_delay must be in [0, 6] (history ring of the device kernel)";
            return false;
        }
        if (n.kind == RSCM_B200_CLIMATE_UDEB && n.params[34] != 2.0) {
            // analytical initial profile depends on kappa and w_initial: they feed the host-computed table
            g.slot_bindable[n.param_base + 3] = 0;
            g.slot_bindable[n.param_base + 6] = 0;
        }
        if (k->rk_step_param != -2) {
            const double h = k->rk_step_param >= 0 ? n.params[k->rk_step_param] : 0.1;
            n.rk_table = g.n_rk++;
            std::vector<int> tab(g.T, 0);
            for (int N = 0; N + 1 < g.T; ++N) tab[N] = rk4_substeps(g.bounds[N], g.bounds[N + 1], h);
            g.rk_nsub.push_back(tab);
        }
    }
    if (g.ctab.size() % 2) g.ctab.push_back(0.0); // keep every staged section a 16-byte multiple
    // HalocarbonChemistry: aux = 1 when every species' emissions are exogenous — then the members of a warp (one scenario)
    // see the same emissions, and the device code may share the 41 species among the lanes (magicc_boxes.cuh)
    for (Node &n : g.nodes) {
        if (n.kind != RSCM_B200_HALOCARBON_CHEMISTRY) continue;
        bool exo = true;
        for (size_t i = 0; i < n.in_src.size(); i += 2) exo = exo && n.in_src[i] == RSCM_B200_SRC_EXOGENOUS;
        n.aux = exo ? 1 : 0;
    }
    emit_program(g);
    return true;
}

int Graph::resolve_slot(const std::string &slot, std::string &err) const
{
    if (slot.rfind("initial:", 0) == 0) {
        const int v = find_var(slot.substr(8));
        if (v < 0) { err = "unknown variable in '" + slot + "'"; return -1000000000; }
        return -(vars[v].cell0) - 1;
    }
    const size_t dot = slot.rfind('.');
    if (dot == std::string::npos) { err = "bad slot '" + slot + "'"; return -1000000000; }
    std::string type = slot.substr(0, dot), field = slot.substr(dot + 1);
    int want_index = -1;
    const size_t hash = type.find('#');
    if (hash != std::string::npos) {
        want_index = std::atoi(type.c_str() + hash + 1);
        type = type.substr(0, hash);
    }
    for (int ni = 0; ni < n_user; ++ni) {
        const Node &n = nodes[ni];
        const KindInfo *k = kind_info(n.kind);
        if (type != k->type_name) continue;
        if (want_index >= 0 && want_index != ni) continue;
        if (k->no_slots) { err = "parameters of " + type + " are per-graph and cannot vary per member"; return -1000000000; }
        for (size_t p = 0; p < k->param_names.size(); ++p) {
            if (field != k->param_names[p]) continue;
            if (!slot_bindable[n.param_base + p]) { err = "parameter '" + slot + "' cannot vary per member"; return -1000000000; }
            return n.param_base + static_cast<int>(p);
        }
        err = "unknown field in '" + slot + "'";
        return -1000000000;
    }
    err = "no component matches '" + slot + "'";
    return -1000000000;
}

} // namespace rscm
