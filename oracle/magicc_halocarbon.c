/* magicc_halocarbon.c — CPU ORACLE (test infrastructure, not product) for the reference's HalocarbonChemistry.
 *
 * Restates crates/rscm-magicc/src/chemistry/halocarbon.rs:
 *   decay_species :115-134, species_forcing :137-145, calculate_{total,fgas,montreal}_forcing :148-196,
 *   calculate_eesc :204-225, definitions :259-293 (hand-written, NOT the derive macro: per species an
 *   `Emissions|<s>` input followed by an `Atmospheric Concentration|<s>` state, then the four outputs),
 *   solve :295-352 (inputs through InputState::get_global, crates/rscm-core/src/state/mod.rs:231-254:
 *   exogenous -> value at the current time, endogenous -> Timeseries::latest_value, i.e. the last non-NaN value).
 * and emission_to_concentration_factor, crates/rscm-magicc/src/parameters/halocarbon.rs:162-172.
 *
 * The species LIST is the reference's default one (parameters/halocarbon.rs:203-262: 23 F-gases then 18 Montreal
 * gases); the per-species NUMBERS are parameters.
 * Parameter layout: br_multiplier, cfc11_release_normalisation, eesc_delay (not used by solve), air_molar_mass,
 * atmospheric_mass_tg, mixing_box_fraction, then per species: lifetime, radiative_efficiency, concentration_pi,
 * molecular_weight, n_cl, n_br, fractional_release.
 *
 * Parity pin: the reference's own unit tests of this file (halocarbon.rs:358-741) restated in
 * tests/test_halocarbon.py. */
#include "orc_internal.h"
#include <stdio.h>

#define HALO_NS 41
#define HALO_NF 23
#define HALO_NG 6

static const char *const halo_names[HALO_NS] = {
    "CF4", "C2F6", "C3F8", "C4F10", "C5F12", "C6F14", "C7F16", "C8F18", "c-C4F8", "HFC-23", "HFC-32", "HFC-43-10mee", "HFC-125",
    "HFC-134a", "HFC-143a", "HFC-152a", "HFC-227ea", "HFC-236fa", "HFC-245fa", "HFC-365mfc", "NF3", "SF6", "SO2F2",
    "CFC-11", "CFC-12", "CFC-113", "CFC-114", "CFC-115", "HCFC-22", "HCFC-141b", "HCFC-142b", "CH3CCl3", "CCl4", "CH3Cl", "CH2Cl2",
    "CHCl3", "CH3Br", "Halon-1211", "Halon-1301", "Halon-2402", "Halon-1202"};

/* InputState::get_global */
static double in_global(const orc_ctx *c, int i)
{
    if (c->node->in_src[i] == ORC_SRC_EXOGENOUS) return orc_in_start(c, i, 0);
    return orc_in_latest(c, i, 0);
}

static int halo_solve(const double *p, orc_ctx *c, double t0, double t1, double *out, void *st)
{
    (void)st;
    const double dt = t1 - t0;
    const double atm_mass_g = p[4] * 1e12;
    double total = 0.0, fgas = 0.0, montreal = 0.0, eesc = 0.0;
    for (int s = 0; s < HALO_NS; ++s) {
        const double *sp = p + HALO_NG + 7 * s;
        const double lifetime = sp[0], rad_eff = sp[1], conc_pi = sp[2], mw = sp[3], n_cl = sp[4], n_br = sp[5], frac = sp[6];
        const double emission = in_global(c, 2 * s), conc = in_global(c, 2 * s + 1);
        const double decay = exp(-dt / lifetime);
        const double conv = (p[3] / mw) * (1e9 / atm_mass_g) * 1e12 / p[5];
        const double emissions_ppt = emission * conv;
        const double new_conc = conc * decay + emissions_ppt * lifetime * (1.0 - decay);
        out[s] = new_conc;
        const double forcing = (new_conc - conc_pi) * rad_eff / 1000.0;
        total += forcing;
        if (s < HALO_NF) fgas += forcing;
        else montreal += forcing;
        if (frac > 0.0) {
            const double halogen_loading = n_cl + p[0] * n_br;
            const double normalised_release = frac / p[1];
            eesc += new_conc * halogen_loading * normalised_release;
        }
    }
    out[HALO_NS] = total;
    out[HALO_NS + 1] = fgas;
    out[HALO_NS + 2] = montreal;
    out[HALO_NS + 3] = eesc;
    return 0;
}

static char halo_name_buf[2 * HALO_NS][ORC_MAX_NAME];
static orc_def halo_defs[2 * HALO_NS + 4];
static orc_kind_info halo_kind;

const orc_kind_info *orc_kind_halocarbon(void)
{
    if (halo_kind.kind == 0) {
        for (int s = 0; s < HALO_NS; ++s) {
            snprintf(halo_name_buf[2 * s], ORC_MAX_NAME, "Emissions|%s", halo_names[s]);
            snprintf(halo_name_buf[2 * s + 1], ORC_MAX_NAME, "Atmospheric Concentration|%s", halo_names[s]);
            halo_defs[2 * s] = (orc_def){halo_name_buf[2 * s], ORC_REQ_INPUT, ORC_GRID_SCALAR};
            halo_defs[2 * s + 1] = (orc_def){halo_name_buf[2 * s + 1], ORC_REQ_STATE, ORC_GRID_SCALAR};
        }
        halo_defs[2 * HALO_NS] = (orc_def){"Forcing|Halocarbons", ORC_REQ_OUTPUT, ORC_GRID_SCALAR};
        halo_defs[2 * HALO_NS + 1] = (orc_def){"Forcing|F-gases", ORC_REQ_OUTPUT, ORC_GRID_SCALAR};
        halo_defs[2 * HALO_NS + 2] = (orc_def){"Forcing|Montreal Gases", ORC_REQ_OUTPUT, ORC_GRID_SCALAR};
        halo_defs[2 * HALO_NS + 3] = (orc_def){"EESC", ORC_REQ_OUTPUT, ORC_GRID_SCALAR};
        halo_kind.type_name = "HalocarbonChemistry";
        halo_kind.n_defs = 2 * HALO_NS + 4;
        halo_kind.defs = halo_defs;
        halo_kind.n_params = HALO_NG + 7 * HALO_NS;
        halo_kind.solve = halo_solve;
        halo_kind.state_size = 0;
        halo_kind.init_state = NULL;
        halo_kind.kind = ORC_HALOCARBON_CHEMISTRY;
    }
    return &halo_kind;
}
