// Reference-side binding for the B200 packet-propagation library.
//
// This translation unit is compiled INSIDE the reference's source environment (its headers on the include
// path) and provides the C++ symbol
//     void update_packets(int nts, std::span<Packet> packets)            (update_packets.h:10)
// in place of the reference's update_packets.cc.  It flattens the reference's global state into the named
// arrays of include/artis_b200.h, forwards the call through the C ABI (loaded with dlopen so that the
// host binary has no link-time CUDA dependency), and folds the returned estimators and counters back
// into the reference's globals so that the rest of do_timestep() (sn3d.cc:790-838) runs unchanged.
//
// Build variants (integration/Makefile, oracle/Makefile):
//   default                       the drop-in: GPU path only, reference update_packets.cc NOT linked.
//   -DARTISB200_WITH_REFERENCE    oracle build: additionally embeds the reference's own update_packets.cc
//                                 (renamed) so that it can be run and its inputs/outputs dumped as
//                                 snapshots for the parity tests.  Test infrastructure, never shipped.
//
// Environment variables
//   ARTISB200_MODE       gpu (default in the drop-in) | ref | ref_perpacket  (oracle build only)
//   ARTISB200_LIB        path of libartis_b200_<preset>.so (gpu mode)
//   ARTISB200_DUMP_DIR   directory for snapshots; ARTISB200_DUMP_TS = comma list of timesteps or "all"
//   ARTISB200_RNG        philox (default) | xoshiro (needs a -DGPU_ON build: Packet carries rngstate)
//   ARTISB200_MAXSTEPS   max packet steps per kernel launch (0 = whole history per launch)
//   ARTISB200_DEVICE     CUDA device ordinal (default: LOCAL_RANK or 0)

#ifdef ARTISB200_WITH_REFERENCE
#define update_packets update_packets_reference_impl
#include "update_packets.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>
#undef update_packets
#endif

#include <dlfcn.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <bit>
#include <cmath>
#include <cstring>
#include <limits>
#include <span>
#include <string>
#include <type_traits>
#include <vector>

#include "artis_b200.h"
#define ARTISB200_REFERENCE_OPTIONS 1
#include "artis_b200_options.h"
#include "artisoptions.h"
#include "atomic.h"
#include "b200_access.h"
#include "b200_snapshot.h"
#include "decay.h"
#include "exspec.h"
#include "gammapkt.h"
#include "globals.h"
#include "grid.h"
#include "ltepop.h"
#include "kpkt.h"
#include "mpi_logging.h"
#include "nltepop.h"
#include "nonthermal.h"
#include "packet.h"
#include "radfield.h"
#include "ratecoeff.h"
#include "rpkt.h"
#include "stats.h"
#include "update_packets.h"
#include "vectors.h"

namespace {

// ---------------------------------------------------------------------------------------------------
// flattening of the reference globals into named arrays (generic over the sink)
// ---------------------------------------------------------------------------------------------------

template <class Sink>
void emit_static(Sink& s) {
  const auto ncoord = grid::b200_ncoordgrid();
  const int64_t ncoord64[3] = {ncoord[0], ncoord[1], ncoord[2]};
  s.i64("scalar.grid_type", static_cast<int64_t>(grid::b200_propgridtype()));
  s.arr("scalar.ncoordgrid", ncoord64, 3);
  s.f64("scalar.tmin", globals::tmin);
  s.f64("scalar.rmax", globals::rmax);
  s.f64("scalar.vmax", globals::vmax);
  s.i64("scalar.nphixspoints", globals::NPHIXSPOINTS);
  s.f64("scalar.nphixsnuincrement", globals::NPHIXSNUINCREMENT);
  s.f64("scalar.last_phixs_nuovernuedge", last_phixs_nuovernuedge);
  s.i64("scalar.tablesize", TABLESIZE);

  for (int axis = 0; axis < 3; axis++) {
    const auto coords = grid::b200_coord_pos_min_tmin(axis);
    const std::string name = "grid.coord_pos_min_tmin" + std::to_string(axis);
    s.arr(name.c_str(), coords.data(), static_cast<int64_t>(coords.size()));
  }
  const auto propcell_nonemptymgi = grid::b200_propcell_nonemptymgi();
  s.arr("grid.propcell_nonemptymgi", propcell_nonemptymgi.data(), static_cast<int64_t>(propcell_nonemptymgi.size()));

  const int nc = grid::get_nonempty_npts_model();
  std::vector<float> ffegrp(nc);
  for (int nonemptymgi = 0; nonemptymgi < nc; nonemptymgi++) {
    ffegrp[nonemptymgi] = grid::get_ffegrp(grid::get_mgi_of_nonemptymgi(nonemptymgi));
  }
  s.arr("cell.ffegrp", ffegrp.data(), nc);
  // inputs of the parameterised thermalisation schemes (gammapkt.cc:777-858, update_packets.cc:68-76)
  std::vector<float> rho_tmin(nc);
  for (int nonemptymgi = 0; nonemptymgi < nc; nonemptymgi++) {
    rho_tmin[nonemptymgi] = grid::get_rho_tmin(grid::get_mgi_of_nonemptymgi(nonemptymgi));
  }
  s.arr("cell.rho_tmin", rho_tmin.data(), nc);
  s.f64("scalar.ejecta_kinetic_energy", grid::get_ejecta_kinetic_energy());
  s.f64("scalar.mtot_input", grid::mtot_input);

  // elements and ions
  const int nelements = get_nelements();
  std::vector<int> e_anumber(nelements);
  std::vector<int> e_nions(nelements);
  std::vector<int> e_lowest(nelements);
  std::vector<int> e_uniqueionstart(nelements);
  const int nions = get_includedions();
  std::vector<int> i_nlevels(nions);
  std::vector<int> i_nlevels_ionising(nions);
  std::vector<int> i_maxrecomb(nions);
  std::vector<int> i_coolingoffset(nions);
  std::vector<int> i_ncoolingterms(nions);
  std::vector<int> i_levelstart(nions);
  std::vector<int> i_groundcontindex(nions);
  std::vector<int> i_nlevels_nlte(nions);
  std::vector<int> i_nltestart(nions);
  std::vector<int> i_nlevels_autoion(nions);
  std::vector<double> i_ionpot(nions);
  for (int element = 0; element < nelements; element++) {
    const auto& el = globals::elements[element];
    e_anumber[element] = el.anumber;
    e_nions[element] = get_nions(element);
    e_lowest[element] = el.lowest_ionstage;
    e_uniqueionstart[element] = el.uniqueionindexstart;
    for (int ion = 0; ion < get_nions(element); ion++) {
      const auto& io = el.ions[ion];
      const int u = get_uniqueionindex(element, ion);
      i_nlevels[u] = io.nlevels;
      i_nlevels_ionising[u] = io.nlevels_ionising;
      i_maxrecomb[u] = io.maxrecombininglevel;
      i_coolingoffset[u] = io.coolingoffset;
      i_ncoolingterms[u] = io.ncoolingterms;
      i_levelstart[u] = io.uniquelevelindexstart;
      i_groundcontindex[u] = io.groundcontindex;
      i_nlevels_nlte[u] = io.nlevels_excited_nlte;
      i_nltestart[u] = io.allnltelevelsindexstart;
      i_nlevels_autoion[u] = io.nlevels_autoion;
      i_ionpot[u] = io.ionpot;
    }
  }
  s.arr("elem.anumber", e_anumber.data(), nelements);
  s.arr("elem.nions", e_nions.data(), nelements);
  s.arr("elem.lowest_ionstage", e_lowest.data(), nelements);
  s.arr("elem.uniqueionindexstart", e_uniqueionstart.data(), nelements);
  s.arr("ion.nlevels", i_nlevels.data(), nions);
  s.arr("ion.nlevels_ionising", i_nlevels_ionising.data(), nions);
  s.arr("ion.maxrecombininglevel", i_maxrecomb.data(), nions);
  s.arr("ion.coolingoffset", i_coolingoffset.data(), nions);
  s.arr("ion.ncoolingterms", i_ncoolingterms.data(), nions);
  s.arr("ion.uniquelevelindexstart", i_levelstart.data(), nions);
  s.arr("ion.groundcontindex", i_groundcontindex.data(), nions);
  s.arr("ion.nlevels_excited_nlte", i_nlevels_nlte.data(), nions);
  s.arr("ion.allnltelevelsindexstart", i_nltestart.data(), nions);
  s.arr("ion.nlevels_autoion", i_nlevels_autoion.data(), nions);
  s.arr("ion.ionpot", i_ionpot.data(), nions);

  // levels
  const auto& lv = globals::alllevels;
  const int64_t nlev = get_includedlevels();
  s.arr("level.epsilon", lv.epsilon.data(), nlev);
  s.arr("level.statweight", lv.statweight.data(), nlev);
  s.arr("level.alltrans_startdown", lv.alltrans_startdown.data(), nlev);
  s.arr("level.ndowntrans", lv.ndowntrans.data(), nlev);
  s.arr("level.nuptrans", lv.nuptrans.data(), nlev);
  s.arr("level.closestgroundlevelcont", lv.closestgroundlevelcont.data(), nlev);
  s.arr("level.phixsstart", lv.phixsstart.data(), nlev);
  s.arr("level.nphixstargets", lv.nphixstargets.data(), nlev);
  s.arr("level.phixstargetstart", lv.phixstargetstart.data(), nlev);
  s.arr("level.bflist_start", lv.bflist_start.data(), nlev);
  s.arr("level.matransblock_start", lv.matransblock_start.data(), nlev);

  // transitions (per level: [down...][up...])
  const auto& tr = globals::alltrans;
  const auto ntrans = static_cast<int64_t>(tr.lineindex.size());
  s.arr("trans.lineindex", tr.lineindex.data(), ntrans);
  s.arr("trans.targetlevelindex", tr.targetlevelindex.data(), ntrans);
  s.arr("trans.einstein_A", tr.einstein_A.data(), ntrans);
  s.arr("trans.coll_str", tr.coll_str.data(), ntrans);
  s.arr("trans.osc_strength", tr.osc_strength.data(), ntrans);
  s.arr("trans.forbidden", tr.forbidden.data(), ntrans);

  // line list (descending nu)
  const auto& ll = globals::linelist;
  const int64_t nlines = globals::nlines;
  s.arr("line.nu", ll.nu.data(), nlines);
  s.arr("line.elementindex", ll.elementindex.data(), nlines);
  s.arr("line.ionindex", ll.ionindex.data(), nlines);
  s.arr("line.lower", ll.uniquelevelindex_lower.data(), nlines);
  s.arr("line.upper", ll.uniquelevelindex_upper.data(), nlines);
  s.arr("line.B_ul", ll.B_ul.data(), nlines);
  s.arr("line.B_lu", ll.B_lu.data(), nlines);

  // bound-free continua (ascending nu_edge)
  const auto& ac = globals::allcont;
  const int64_t nbf = globals::nbfcontinua;
  s.arr("cont.nu_edge", ac.nu_edge.data(), nbf);
  s.arr("cont.element", ac.element.data(), nbf);
  s.arr("cont.ion", ac.ion.data(), nbf);
  s.arr("cont.level", ac.level.data(), nbf);
  s.arr("cont.phixstargetindex", ac.phixstargetindex.data(), nbf);
  s.arr("cont.upperlevel", ac.upperlevel.data(), nbf);
  s.arr("cont.uniquelevelindex", ac.uniquelevelindex.data(), nbf);
  s.arr("cont.probability", ac.probability.data(), nbf);
  s.arr("cont.groundcontestimindex", ac.groundcontestimindex.data(), nbf);
  s.arr("cont.bfestimindex", ac.bfestimindex.data(), nbf);

  s.arr("phixs.table", globals::allphixs.data(), static_cast<int64_t>(globals::allphixs.size()));
  s.arr("phixstarget.levelindex", globals::allphixstargets_levelindex.data(),
        static_cast<int64_t>(globals::allphixstargets_levelindex.size()));
  s.arr("phixstarget.probability", globals::allphixstargets_probability.data(),
        static_cast<int64_t>(globals::allphixstargets_probability.size()));
  s.arr("groundcont.nu_edge", globals::groundcont_nu_edge.data(), static_cast<int64_t>(globals::groundcont_nu_edge.size()));
  s.arr("bfestim.nu_edge", globals::bfestim_nu_edge.data(), static_cast<int64_t>(globals::bfestim_nu_edge.size()));

  const auto lut_sp = b200_lut_spontrecombcoeffs();
  const auto lut_cp = b200_lut_corrphotoioncoeffs();
  const auto lut_bc = b200_lut_bfcooling_coeffs();
  const auto lut_tg = b200_lut_temperature_grid();
  s.arr("lut.spontrecomb", lut_sp.data(), static_cast<int64_t>(lut_sp.size()));
  s.arr("lut.corrphotoion", lut_cp.data(), static_cast<int64_t>(lut_cp.size()));
  s.arr("lut.bfcooling", lut_bc.data(), static_cast<int64_t>(lut_bc.size()));
  s.arr("lut.temperature_grid", lut_tg.data(), static_cast<int64_t>(lut_tg.size()));
  if constexpr (USE_XCOM_GAMMAPHOTOION) {
    // XCOM photoionisation tables (gammapkt.cc:52-58, read by init_xcom_photoion_data 244-262): rows of Z = 1..100
    std::vector<int> zstart;
    std::vector<double> energy;
    std::vector<double> sigma;
    gammapkt::b200_xcom_tables(zstart, energy, sigma);
    s.arr("xcom.zstart", zstart.data(), static_cast<int64_t>(zstart.size()));
    s.arr("xcom.energy", energy.data(), static_cast<int64_t>(energy.size()));
    s.arr("xcom.sigma", sigma.data(), static_cast<int64_t>(sigma.size()));
  }

  const int ncool = kpkt::ncoolingterms;
  std::vector<unsigned char> c_type(ncool);
  std::vector<int> c_level(ncool);
  std::vector<int> c_target(ncool);
  for (int i = 0; i < ncool; i++) {
    c_type[i] = static_cast<unsigned char>(kpkt::b200_coolinglist_type(i));
    c_level[i] = kpkt::b200_coolinglist_level(i);
    c_target[i] = kpkt::b200_coolinglist_phixstargetindex(i);
  }
  s.arr("cooling.type", c_type.data(), ncool);
  s.arr("cooling.level", c_level.data(), ncool);
  s.arr("cooling.phixstargetindex", c_target.data(), ncool);

  const auto nts_total = static_cast<int64_t>(globals::timesteps.size());
  std::vector<double> t_start(nts_total);
  std::vector<double> t_width(nts_total);
  std::vector<double> t_mid(nts_total);
  for (int64_t i = 0; i < nts_total; i++) {
    t_start[i] = globals::timesteps[i].start;
    t_width[i] = globals::timesteps[i].width;
    t_mid[i] = globals::timesteps[i].mid;
  }
  s.arr("timesteps.start", t_start.data(), nts_total);
  s.arr("timesteps.width", t_width.data(), nts_total);
  s.arr("timesteps.mid", t_mid.data(), nts_total);
}

template <class Sink>
void emit_timestep_state(Sink& s, const int nts) {
  const int64_t nc = grid::get_nonempty_npts_model();
  s.i64("scalar.nts", nts);
  s.i64("scalar.globals_timestep", globals::timestep);
  s.f64("scalar.max_path_step", globals::max_path_step);
  s.arr("cell.rho", grid::rho_allcells.data(), nc);
  s.arr("cell.Te", grid::Te_allcells.data(), nc);
  s.arr("cell.TJ", grid::TJ_allcells.data(), nc);
  s.arr("cell.TR", grid::TR_allcells.data(), nc);
  s.arr("cell.W", grid::W_allcells.data(), nc);
  s.arr("cell.nne", grid::nne_allcells.data(), nc);
  s.arr("cell.nnetot", grid::nnetot_allcells.data(), nc);
  s.arr("cell.kappagrey", grid::kappagrey_allcells.data(), nc);
  static_assert(sizeof(grid::CellThickness) == sizeof(int));
  s.arr("cell.thick", reinterpret_cast<const int*>(grid::thick_allcells.data()), nc);
  std::vector<float> clump(nc);
  for (int64_t i = 0; i < nc; i++) {
    clump[i] = grid::get_clumpfactor(static_cast<int>(i));
  }
  s.arr("cell.clumpfactor", clump.data(), nc);
  s.arr("cell.elem_massfracs", grid::elem_massfracs_allcells.data(),
        static_cast<int64_t>(grid::elem_massfracs_allcells.size()));
  s.arr("cell.ion_groundlevelpops", grid::ion_groundlevelpops_allcells.data(),
        static_cast<int64_t>(grid::ion_groundlevelpops_allcells.size()));
  s.arr("cell.ion_partfuncts", grid::ion_partfuncts_allcells.data(),
        static_cast<int64_t>(grid::ion_partfuncts_allcells.size()));
  s.arr("cell.ion_cooling_contribs", kpkt::ion_cooling_contribs_allcells.data(),
        static_cast<int64_t>(kpkt::ion_cooling_contribs_allcells.size()));
  s.arr("cell.corrphotoionrenorm", globals::corrphotoionrenorm.data(),
        static_cast<int64_t>(globals::corrphotoionrenorm.size()));
  if constexpr (!USE_LUT_PHOTOION && DETAILED_BF_ESTIMATORS_ON) {
    // the previous timestep's normalised bound-free rate estimators (radfield.cc:95, 923): from
    // DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP on they are the photoionisation coefficients (ratecoeff.cc:848-851)
    const auto prev = radfield::b200_prev_bfrate_normed();
    s.arr("radfield.prev_bfrate_normed", prev.data(), static_cast<int64_t>(prev.size()));
  }
  if constexpr (!USE_LUT_PHOTOION && std::is_same_v<Sink, b200::SnapshotWriter>) {
    // Oracle snapshots only (known-answer vectors; the library evaluates these itself on the device): the corrected
    // photoionisation rate coefficients without the LUT as the reference's own get_corrphotoioncoeff gives them
    // (ratecoeff.cc:840-875: the estimator above or an integral over the radiation field model)
    const int nlevels_total = get_includedlevels();
    std::vector<double> gammacorr;
    gammacorr.reserve(static_cast<size_t>(nc) * static_cast<size_t>(nlevels_total));
    for (int64_t cell = 0; cell < nc; cell++) {
      for (int element = 0; element < get_nelements(); element++) {
        for (int ion = 0; ion < get_nions(element); ion++) {
          for (int level = 0; level < get_nlevels(element, ion); level++) {
            const int ntargets = get_nphixstargets(element, ion, level);
            const bool ionises = (ion < get_nions(element) - 1) && (level < get_nlevels_ionising(element, ion));
            for (int k = 0; k < ntargets; k++) {
              gammacorr.push_back(ionises ? get_corrphotoioncoeff(element, ion, level, k, static_cast<int>(cell), false) : 0.);
            }
          }
        }
      }
    }
    s.arr("cell.corrphotoioncoeff", gammacorr.data(), static_cast<int64_t>(gammacorr.size()));
  }
  if constexpr (NT_ON) {
    // non-thermal routing state (ref_access/ref_nonthermal.cc): rate coefficients and channel probabilities per ion
    std::vector<double> ratecoeff;
    std::vector<double> energyrate;
    std::vector<float> prob;
    std::vector<float> enfrac;
    std::vector<float> fracion;
    nonthermal::b200_nt_cell_state(ratecoeff, energyrate, prob, enfrac, fracion);
    s.arr("cell.nt_ionisation_ratecoeff", ratecoeff.data(), static_cast<int64_t>(ratecoeff.size()));
    s.arr("cell.nt_ion_energyrate", energyrate.data(), static_cast<int64_t>(energyrate.size()));
    s.arr("cell.nt_prob_num_auger", prob.data(), static_cast<int64_t>(prob.size()));
    s.arr("cell.nt_ionenfrac_num_auger", enfrac.data(), static_cast<int64_t>(enfrac.size()));
    s.arr("cell.nt_frac_ionisation", fracion.data(), static_cast<int64_t>(fracion.size()));
    if constexpr (NT_EXCITATION_ON) {
      int stride = 0;
      std::vector<int> count;
      std::vector<int> alltransindex;
      std::vector<double> fracdep;
      std::vector<double> rateperdep;
      std::vector<double> deprate;
      std::vector<float> fracexc;
      nonthermal::b200_nt_excitations(stride, count, alltransindex, fracdep, rateperdep, deprate, fracexc);
      s.i64("scalar.nt_excitations_stored", stride);
      s.arr("cell.nt_exc_count", count.data(), static_cast<int64_t>(count.size()));
      s.arr("cell.nt_exc_alltransindex", alltransindex.data(), static_cast<int64_t>(alltransindex.size()));
      s.arr("cell.nt_exc_frac_deposition", fracdep.data(), static_cast<int64_t>(fracdep.size()));
      s.arr("cell.nt_exc_ratecoeffperdeposition", rateperdep.data(), static_cast<int64_t>(rateperdep.size()));
      s.arr("cell.nt_deposition_rate_density", deprate.data(), static_cast<int64_t>(deprate.size()));
      s.arr("cell.nt_frac_excitation", fracexc.data(), static_cast<int64_t>(fracexc.size()));
    }
  }
  {
    // element number densities of the timestep (grid.cc:1693-1697), read by the XCOM photoelectric opacity (gammapkt.cc:456)
    // and by the LTE ion balance on the device (artisb200_update_grid_lte)
    const int nel = get_nelements();
    std::vector<double> numberdens(static_cast<size_t>(nc) * static_cast<size_t>(nel));
    for (int64_t cell = 0; cell < nc; cell++) {
      for (int element = 0; element < nel; element++) {
        numberdens[(static_cast<size_t>(cell) * static_cast<size_t>(nel)) + static_cast<size_t>(element)] = grid::get_elem_numberdens(cell, element);
      }
    }
    s.arr("cell.elem_numberdens", numberdens.data(), static_cast<int64_t>(numberdens.size()));
  }
  if constexpr (RPKT_USE_EXPANSION_OPACITIES) {
    // binned expansion opacities [cm^2/g] of this timestep (rpkt.h:47, written by calculate_expansion_opacities,
    // rpkt.cc:1071-1123, from update_grid): per-timestep cell state like the temperatures
    s.arr("cell.expansionopacities", expansionopacities.data(), static_cast<int64_t>(expansionopacities.size()));
  }
  if constexpr (RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY.has_value()) {
    const auto cumulative = b200_expansionopacity_planck_cumulative();  // rpkt.cc:48 (ref_access/ref_rpkt.cc)
    s.arr("cell.expopac_planck_cumulative", cumulative.data(), static_cast<int64_t>(cumulative.size()));
  }
  if (globals::total_nlte_levels > 0) {
    // NLTE solver populations over rho, one slot per NLTE level and superlevel (nltepop.h:15, nltepop.cc:1955-1968)
    s.arr("cell.nltepops", nltepops_allcells.data(), static_cast<int64_t>(nltepops_allcells.size()));
  }
  if constexpr (MULTIBIN_RADFIELD_MODEL_ON) {
    // fitted (W, T_R) of every frequency bin: radfield::radfield(nu, cell) reads them (radfield.cc:786-797)
    const auto w = radfield::b200_bin_solutions_W();
    const auto tr = radfield::b200_bin_solutions_T_R();
    s.arr("radfield.bin_W", w.data(), static_cast<int64_t>(w.size()));
    s.arr("radfield.bin_T_R", tr.data(), static_cast<int64_t>(tr.size()));
  }
}

void fill_ts_scalars(const int nts, double* out) {
  const auto& ts = globals::timesteps[nts];
  out[ARTISB200_TS_GAMMA_DEP_DISCRETE] = ts.gamma_dep_discrete;
  out[ARTISB200_TS_POSITRON_DEP_DISCRETE] = ts.positron_dep_discrete;
  out[ARTISB200_TS_POSITRON_EMISSION] = ts.positron_emission;
  out[ARTISB200_TS_ELECTRON_DEP_DISCRETE] = ts.electron_dep_discrete;
  out[ARTISB200_TS_ELECTRON_EMISSION] = ts.electron_emission;
  out[ARTISB200_TS_ALPHA_DEP_DISCRETE] = ts.alpha_dep_discrete;
  out[ARTISB200_TS_ALPHA_EMISSION] = ts.alpha_emission;
  out[ARTISB200_TS_SPFISSION_DEP_DISCRETE] = ts.spfission_dep_discrete;
  out[ARTISB200_TS_GAMMA_EMISSION] = ts.gamma_emission;
  out[ARTISB200_TS_NT_ENERGY_DEPOSITED] = 0.;  // nonthermal.cc:200 is file-static and only logged
}

template <class Sink>
void emit_estimators(Sink& s, const int nts) {
  const auto j = radfield::b200_J();
  const auto nuj = radfield::b200_nuJ();
  s.arr("est.J", j.data(), static_cast<int64_t>(j.size()));
  s.arr("est.nuJ", nuj.data(), static_cast<int64_t>(nuj.size()));
  s.arr("est.ffheating", globals::ffheatingestimator.data(), static_cast<int64_t>(globals::ffheatingestimator.size()));
  s.arr("est.colheating", globals::colheatingestimator.data(), static_cast<int64_t>(globals::colheatingestimator.size()));
  s.arr("est.gamma", globals::gammaestimator.data(), static_cast<int64_t>(globals::gammaestimator.size()));
  s.arr("est.bfheating", globals::bfheatingestimator.data(), static_cast<int64_t>(globals::bfheatingestimator.size()));
  s.arr("est.dep_gamma", globals::dep_estimator_gamma.data(), static_cast<int64_t>(globals::dep_estimator_gamma.size()));
  s.arr("est.dep_positron", globals::dep_estimator_positron.data(), static_cast<int64_t>(globals::dep_estimator_positron.size()));
  s.arr("est.dep_electron", globals::dep_estimator_electron.data(), static_cast<int64_t>(globals::dep_estimator_electron.size()));
  s.arr("est.dep_alpha", globals::dep_estimator_alpha.data(), static_cast<int64_t>(globals::dep_estimator_alpha.size()));
  if constexpr (DETAILED_BF_ESTIMATORS_ON) {
    const auto bfrate = radfield::b200_bfrate_raw();
    s.arr("est.bfrate_raw", bfrate.data(), static_cast<int64_t>(bfrate.size()));
  }
  if constexpr (MULTIBIN_RADFIELD_MODEL_ON) {
    const auto jraw = radfield::b200_bins_J_raw();
    const auto nujraw = radfield::b200_bins_nuJ_raw();
    s.arr("est.bins_J_raw", jraw.data(), static_cast<int64_t>(jraw.size()));
    s.arr("est.bins_nuJ_raw", nujraw.data(), static_cast<int64_t>(nujraw.size()));
  }
  double tss[ARTISB200_NTSSCALARS];
  fill_ts_scalars(nts, tss);
  s.arr("ts.scalars", tss, ARTISB200_NTSSCALARS);
  s.i64("ts.pellet_decays", globals::timesteps[nts].pellet_decays);
  int64_t counters[static_cast<int>(stats::Counter::COUNT)];
  for (int i = 0; i < static_cast<int>(stats::Counter::COUNT); i++) {
    counters[i] = stats::get_counter(static_cast<stats::Counter>(i));
  }
  s.arr("counters", counters, static_cast<int>(stats::Counter::COUNT));
}

template <class Sink>
void emit_packets(Sink& s, const std::span<const Packet> packets) {
  s.i64("packets.count", static_cast<int64_t>(packets.size()));
  s.i64("packets.stride", static_cast<int64_t>(sizeof(Packet)));
  s.arr("packets.aos", reinterpret_cast<const unsigned char*>(packets.data()),
        static_cast<int64_t>(packets.size() * sizeof(Packet)));
}

// ---------------------------------------------------------------------------------------------------
// the library, loaded at run time
// ---------------------------------------------------------------------------------------------------

struct Lib {
  void* handle{nullptr};
  decltype(&artisb200_create) create{};
  decltype(&artisb200_last_error) last_error{};
  decltype(&artisb200_set_array) set_array{};
  decltype(&artisb200_get_array) get_array{};
  decltype(&artisb200_array_count) array_count{};
  decltype(&artisb200_set_option) set_option{};
  decltype(&artisb200_commit_static) commit_static{};
  decltype(&artisb200_begin_timestep) begin_timestep{};
  decltype(&artisb200_update_packets_host) update_packets_host{};
  decltype(&artisb200_last_timing_ms) last_timing_ms{};
  decltype(&artisb200_options_summary) options_summary{};
  decltype(&artisb200_options_hash) options_hash{};
  decltype(&artisb200_register_host_buffer) register_host_buffer{};
  decltype(&artisb200_unregister_host_buffer) unregister_host_buffer{};
  void* registered{nullptr};
  artisb200_ctx* ctx{nullptr};
};

Lib lib;

template <class F>
void load_symbol(F& fn, const char* name) {
  fn = reinterpret_cast<F>(dlsym(lib.handle, name));
  if (fn == nullptr) {
    printlnlog("[fatal] artis_b200: symbol {} missing from library", name);
    std::abort();
  }
}

void check(const int rc, const char* what) {
  if (rc != 0) {
    // same convention as assert_always (mpi_logging.h:123-130): log, then abort
    printlnlog("[fatal] artis_b200: {} failed: {}", what, lib.last_error(lib.ctx));
    std::fprintf(stderr, "[fatal] artis_b200: %s failed: %s\n", what, lib.last_error(lib.ctx));
    std::abort();
  }
}

struct LibSink {
  template <class T>
  void arr(const char* name, const T* data, const int64_t count) {
    check(lib.set_array(lib.ctx, name, b200::dtype_of<T>::code, data, count), name);
  }
  void f64(const char* name, const double v) { arr(name, &v, 1); }
  void i64(const char* name, const int64_t v) { arr(name, &v, 1); }
};

auto env_or(const char* name, const char* fallback) -> std::string {
  const char* v = std::getenv(name);
  return (v != nullptr) ? std::string(v) : std::string(fallback);
}

// the run's random number seed: the first line of input.txt (input.cc:1883-1900: a positive value is used as it is,
// anything else asks for a random seed, chosen by rank 0 and broadcast). The library adds the rank itself.
auto read_pre_zseed() -> long long {
  std::int64_t pre_zseed = -1;
  if (FILE* f = std::fopen("input.txt", "r"); f != nullptr) {
    if (std::fscanf(f, "%ld", &pre_zseed) != 1) {
      pre_zseed = -1;
    }
    std::fclose(f);
  }
  if (pre_zseed <= 0) {
    pre_zseed = get_rng_random_seed();
    MPI_Bcast(&pre_zseed, 1, MPI_INT64_T, 0, MPI_COMM_WORLD);
  }
  return pre_zseed;
}

void lib_init() {
  const auto path = env_or("ARTISB200_LIB", "libartis_b200.so");
  lib.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (lib.handle == nullptr) {
    printlnlog("[fatal] artis_b200: cannot load {}: {}", path, dlerror());
    std::fprintf(stderr, "[fatal] artis_b200: cannot load %s\n", path.c_str());
    std::abort();
  }
  load_symbol(lib.create, "artisb200_create");
  load_symbol(lib.last_error, "artisb200_last_error");
  load_symbol(lib.set_array, "artisb200_set_array");
  load_symbol(lib.get_array, "artisb200_get_array");
  load_symbol(lib.array_count, "artisb200_array_count");
  load_symbol(lib.set_option, "artisb200_set_option");
  load_symbol(lib.commit_static, "artisb200_commit_static");
  load_symbol(lib.begin_timestep, "artisb200_begin_timestep");
  load_symbol(lib.update_packets_host, "artisb200_update_packets_host");
  load_symbol(lib.last_timing_ms, "artisb200_last_timing_ms");
  load_symbol(lib.options_summary, "artisb200_options_summary");
  load_symbol(lib.options_hash, "artisb200_options_hash");
  load_symbol(lib.register_host_buffer, "artisb200_register_host_buffer");
  load_symbol(lib.unregister_host_buffer, "artisb200_unregister_host_buffer");

  const int device = std::atoi(env_or("ARTISB200_DEVICE", env_or("LOCAL_RANK", "0").c_str()).c_str());
  if (lib.create(&lib.ctx, device) != 0) {
    printlnlog("[fatal] artis_b200: create failed: {}", lib.last_error(nullptr));
    std::fprintf(stderr, "[fatal] artis_b200: create failed: %s\n", lib.last_error(nullptr));
    std::abort();
  }
  printlnlog("artis_b200: library {} on device {} options [{}]", path, device, lib.options_summary());
  // the library is compiled for one set of artisoptions.h values: it must be THIS program's set (include/artis_b200_options.h)
  if (lib.options_hash() != artisb200_options_hash_here() && env_or("ARTISB200_IGNORE_OPTIONS_HASH", "0") != "1") {
    printlnlog("[fatal] artis_b200: {} was compiled for other artisoptions.h values than this program (hash {:x} vs {:x})", path,
               lib.options_hash(), artisb200_options_hash_here());
    std::fprintf(stderr, "[fatal] artis_b200: %s was compiled for other artisoptions.h values than this program\n", path.c_str());
    std::abort();
  }
  const bool xoshiro = env_or("ARTISB200_RNG", "philox") == "xoshiro";
#ifndef GPU_ON
  if (xoshiro) {
    printlnlog("[fatal] artis_b200: ARTISB200_RNG=xoshiro needs a -DGPU_ON host build (per-packet rngstate)");
    std::abort();
  }
#endif
  check(lib.set_option(lib.ctx, "rng_mode", xoshiro ? ARTISB200_RNG_XOSHIRO : ARTISB200_RNG_PHILOX), "rng_mode");
  check(lib.set_option(lib.ctx, "seed", read_pre_zseed()), "seed");
  check(lib.set_option(lib.ctx, "rank", globals::my_rank), "rank");
  // finished packets are copied back while the others are still propagated; the span comes back permuted, like the
  // reference's own update_packets leaves it (ARTISB200_STREAM_DOWNLOAD=0: ordered download)
  check(lib.set_option(lib.ctx, "stream_download", std::atoi(env_or("ARTISB200_STREAM_DOWNLOAD", "1").c_str())), "stream_download");
  check(lib.set_option(lib.ctx, "nranks", globals::nprocs), "nranks");
  check(lib.set_option(lib.ctx, "max_steps_per_launch", std::atoll(env_or("ARTISB200_MAXSTEPS", "-1").c_str())),
        "max_steps_per_launch");
  LibSink sink;
  emit_static(sink);
  check(lib.commit_static(lib.ctx), "commit_static");
}

template <class T>
void fetch_add(const char* name, std::span<T> dest) {
  if (dest.empty()) {
    return;
  }
  // the reference allocates some per-cell estimators with one spare element (nonempty_npts_model + 1); the
  // library's arrays have exactly one entry per non-empty cell
  const int64_t have = lib.array_count(lib.ctx, name);
  if (have < 0 || have > static_cast<int64_t>(dest.size())) {
    printlnlog("[fatal] artis_b200: estimator {} has {} entries, host array has {}", name, have, dest.size());
    std::abort();
  }
  static std::vector<T> tmp;
  tmp.resize(static_cast<size_t>(have));
  check(lib.get_array(lib.ctx, name, b200::dtype_of<T>::code, tmp.data(), have), name);
  for (int64_t i = 0; i < have; i++) {
    dest[i] += tmp[i];
  }
}

void update_packets_gpu(const int nts, std::span<Packet> packets) {
  if (lib.ctx == nullptr) {
    lib_init();
  }
  const auto t0 = std::chrono::steady_clock::now();
  LibSink sink;
  emit_timestep_state(sink, nts);
  check(lib.begin_timestep(lib.ctx, nts), "begin_timestep");
  if (lib.registered != packets.data()) {
    // page-lock the packet array once (sn3d.cc keeps one std::vector<Packet> for the whole run): full-speed, asynchronous copies
    if (lib.registered != nullptr) {
      check(lib.unregister_host_buffer(lib.ctx, lib.registered), "unregister_host_buffer");
    }
    check(lib.register_host_buffer(lib.ctx, packets.data(), static_cast<int64_t>(packets.size_bytes())), "register_host_buffer");
    lib.registered = packets.data();
  }
  check(lib.update_packets_host(lib.ctx, nts, packets.data(), static_cast<int64_t>(packets.size()),
                                static_cast<int>(sizeof(Packet))),
        "update_packets_host");

  // fold the device estimators into the (zeroed, sn3d.cc:718) host estimators
  fetch_add<double>("est.J", radfield::b200_J());
  fetch_add<double>("est.nuJ", radfield::b200_nuJ());
  fetch_add<double>("est.ffheating", globals::ffheatingestimator);
  fetch_add<double>("est.colheating", globals::colheatingestimator);
  fetch_add<double>("est.gamma", globals::gammaestimator);
  fetch_add<double>("est.bfheating", globals::bfheatingestimator);
  fetch_add<double>("est.dep_gamma", globals::dep_estimator_gamma);
  fetch_add<double>("est.dep_positron", globals::dep_estimator_positron);
  fetch_add<double>("est.dep_electron", globals::dep_estimator_electron);
  fetch_add<double>("est.dep_alpha", globals::dep_estimator_alpha);
  if constexpr (DETAILED_BF_ESTIMATORS_ON) {
    fetch_add<double>("est.bfrate_raw", radfield::b200_bfrate_raw());
  }
  if constexpr (MULTIBIN_RADFIELD_MODEL_ON) {
    fetch_add<double>("est.bins_J_raw", radfield::b200_bins_J_raw());
    fetch_add<double>("est.bins_nuJ_raw", radfield::b200_bins_nuJ_raw());
  }
  double tss[ARTISB200_NTSSCALARS];
  check(lib.get_array(lib.ctx, "ts.scalars", 'd', tss, ARTISB200_NTSSCALARS), "ts.scalars");
  auto& ts = globals::timesteps[nts];
  ts.gamma_dep_discrete += tss[ARTISB200_TS_GAMMA_DEP_DISCRETE];
  ts.positron_dep_discrete += tss[ARTISB200_TS_POSITRON_DEP_DISCRETE];
  ts.positron_emission += tss[ARTISB200_TS_POSITRON_EMISSION];
  ts.electron_dep_discrete += tss[ARTISB200_TS_ELECTRON_DEP_DISCRETE];
  ts.electron_emission += tss[ARTISB200_TS_ELECTRON_EMISSION];
  ts.alpha_dep_discrete += tss[ARTISB200_TS_ALPHA_DEP_DISCRETE];
  ts.alpha_emission += tss[ARTISB200_TS_ALPHA_EMISSION];
  ts.spfission_dep_discrete += tss[ARTISB200_TS_SPFISSION_DEP_DISCRETE];
  ts.gamma_emission += tss[ARTISB200_TS_GAMMA_EMISSION];
  int64_t pellet_decays = 0;
  check(lib.get_array(lib.ctx, "ts.pellet_decays", 'q', &pellet_decays, 1), "ts.pellet_decays");
  ts.pellet_decays += static_cast<int>(pellet_decays);
  int64_t counters[static_cast<int>(stats::Counter::COUNT)];
  check(lib.get_array(lib.ctx, "counters", 'q', counters, static_cast<int>(stats::Counter::COUNT)), "counters");
  for (int i = 0; i < static_cast<int>(stats::Counter::COUNT); i++) {
    stats::b200_add_counter(i, counters[i]);
  }
  double total_ms = 0.;
  double prop_ms = 0.;
  double sched_ms = 0.;
  lib.last_timing_ms(lib.ctx, &total_ms, &prop_ms, &sched_ms);
  const auto wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  stats::pkt_action_counters_printout(nts);
  printlnlog(
      "timestep {}: finished update_packets for rank {} on B200 (took {:.3f} seconds wall; device {:.3f} ms total, "
      "{:.3f} ms propagate, {:.3f} ms schedule)",
      nts, globals::my_rank, wall, total_ms, prop_ms, sched_ms);
}

#ifdef ARTISB200_WITH_REFERENCE
// Reference physics in an order-independent schedule: every packet is run through the reference's own
// do_packet() (update_packets.cc:257) from the start to the end of the timestep with its own fresh
// continuum-opacity cache. With the GPU_ON build (per-packet RNG state, every cell's cache slot
// precomputed, update_packets.cc:551-563) a packet's history then depends on nothing but the packet
// itself, which is what makes packet-by-packet comparison with the device path possible.
void update_packets_reference_perpacket(const int nts, std::span<Packet> packets) {
  if constexpr (cellcache_singleslot) {
    printlnlog("[fatal] artis_b200: ARTISB200_MODE=ref_perpacket needs the -DGPU_ON (multi-slot cell cache) oracle build");
    std::abort();
  }
  const double ts_end = globals::timesteps[nts].start + globals::timesteps[nts].width;
  const auto nonempty_npts_model = grid::get_nonempty_npts_model();
  for (int nonemptymgi = 0; nonemptymgi < nonempty_npts_model; nonemptymgi++) {
    cellcacheslot_populate(globals::cellcache.at(nonemptymgi), nonemptymgi);
  }
  for (auto& pkt : packets) {
    ContinuumOpacity chi_rpkt_cont{};
    while (packetprop_update_required(pkt, ts_end)) {
      do_packet(pkt, ts_end, nts, chi_rpkt_cont);
    }
  }
  stats::pkt_action_counters_printout(nts);
}
#endif


#ifdef ARTISB200_WITH_REFERENCE
// Deterministic golden vectors straight from the reference's own functions, evaluated on seeded inputs in
// the state of timestep nts: grid::boundary_distance (grid.cc:2480), closest_transition (rpkt.h:144),
// calculate_chi_rpkt_cont<true> (rpkt.cc:1020), and the reference's per-cell cache tables
// (update_packets.cc:397-464; macroatom.cc:64-200; kpkt.cc:57-229). Inputs are stored next to outputs.
struct KatRng {
  std::uint64_t s;
  auto next() -> double {  // 53-bit LCG-derived uniform in [0,1)
    s = (s * 6364136223846793005ULL) + 1442695040888963407ULL;
    return static_cast<double>(s >> 11U) * 0x1.0p-53;
  }
};

template <class Sink>
void emit_reference_kats(Sink& s, const int nts) {
  KatRng rng{0x9E3779B97F4A7C15ULL + static_cast<std::uint64_t>(nts)};
  const double tstart_base = globals::timesteps[nts].start;
  const double twidth = globals::timesteps[nts].width;
  const auto ncoord = grid::b200_ncoordgrid();
  const auto gridtype = grid::b200_propgridtype();
  const int ndim = (gridtype == GridType::SPHERICAL1D) ? 1 : ((gridtype == GridType::CYLINDRICAL2D) ? 2 : 3);
  const auto ngrid = static_cast<int>(grid::ngrid);

  // ---- boundary_distance ----
  // (evaluated without the max_path_step cap of grid.cc:2750 so that real face distances are pinned; the
  //  cap itself is exercised by every packet history)
  const double saved_max_path_step = globals::max_path_step;
  globals::max_path_step = std::numeric_limits<double>::max();
  s.f64("kat.bd.max_path_step", globals::max_path_step);
  constexpr int NBD = 4000;
  std::vector<double> bd_in(static_cast<size_t>(NBD) * 7);
  std::vector<int> bd_cell(NBD);
  std::vector<double> bd_dist(NBD);
  std::vector<int> bd_next(NBD);
  for (int k = 0; k < NBD; k++) {
    const int cellindex = static_cast<int>(rng.next() * ngrid) % ngrid;
    const double t = tstart_base + (rng.next() * twidth);
    const double trat = t / globals::tmin;
    std::array<int, 3> idx{};
    int rem = cellindex;
    std::array<double, 3> cmin{};
    std::array<double, 3> cmax{};
    for (int d = 0; d < ndim; d++) {
      idx[d] = rem % ncoord[d];
      rem /= ncoord[d];
      const auto coords = grid::b200_coord_pos_min_tmin(d);
      cmin[d] = coords[idx[d]];
      cmax[d] = (idx[d] < ncoord[d] - 1) ? coords[idx[d] + 1] : globals::rmax;
    }
    Vec3d pos{};
    // a fraction of the points sit exactly on / a rounding error beyond a face to exercise the tolerance branch
    const auto pick = [&](const double lo, const double hi) {
      const double u = rng.next();
      if (gridtype == GridType::CARTESIAN3D) {  // curved faces have no valid "exactly on the face" state in general
        if (u < 0.03) { return lo; }
        if (u < 0.06) { return hi; }
        if (u < 0.08) { return hi * (1. + 2e-16); }
      }
      return lo + ((hi - lo) * rng.next());
    };
    if (gridtype == GridType::CARTESIAN3D) {
      for (int d = 0; d < 3; d++) {
        pos[d] = pick(cmin[d], cmax[d]) * trat;
      }
    } else if (gridtype == GridType::CYLINDRICAL2D) {
      const double rcyl = pick(cmin[0], cmax[0]) * trat;
      const double phi = 2 * PI * rng.next();
      pos = {rcyl * std::cos(phi), rcyl * std::sin(phi), pick(cmin[1], cmax[1]) * trat};
    } else {
      const double r = pick(cmin[0], cmax[0]) * trat;
      const double mu = (2 * rng.next()) - 1;
      const double phi = 2 * PI * rng.next();
      const double st = std::sqrt(1 - (mu * mu));
      pos = {r * st * std::cos(phi), r * st * std::sin(phi), r * mu};
    }
    const double mu = (2 * rng.next()) - 1;
    const double phi = 2 * PI * rng.next();
    const double st = std::sqrt(1 - (mu * mu));
    Vec3d dir{st * std::cos(phi), st * std::sin(phi), mu};
    if (rng.next() < 0.02) {
      dir = {0., 0., (mu > 0) ? 1. : -1.};  // exactly along z: the dirxylen == 0 branch of the 2D grid
    }
    const auto [dist, next] = grid::boundary_distance(dir, pos, t, cellindex);
    for (int d = 0; d < 3; d++) {
      bd_in[(static_cast<size_t>(k) * 7) + d] = pos[d];
      bd_in[(static_cast<size_t>(k) * 7) + 3 + d] = dir[d];
    }
    bd_in[(static_cast<size_t>(k) * 7) + 6] = t;
    bd_cell[k] = cellindex;
    bd_dist[k] = dist;
    bd_next[k] = next;
  }
  globals::max_path_step = saved_max_path_step;
  s.arr("kat.bd.in", bd_in.data(), static_cast<int64_t>(bd_in.size()));
  s.arr("kat.bd.cell", bd_cell.data(), NBD);
  s.arr("kat.bd.dist", bd_dist.data(), NBD);
  s.arr("kat.bd.next", bd_next.data(), NBD);

  // ---- closest_transition ----
  constexpr int NCT = 4000;
  std::vector<double> ct_nu(NCT);
  std::vector<int> ct_next(NCT);
  std::vector<int> ct_out(NCT);
  const auto linenu = globals::linelist.nu.span();
  const double lognumax = std::log(linenu.front() * 1.05);
  const double lognumin = std::log(linenu.back() * 0.95);
  for (int k = 0; k < NCT; k++) {
    double nu = std::exp(lognumin + ((lognumax - lognumin) * rng.next()));
    const double u = rng.next();
    if (u < 0.1) {
      nu = linenu[static_cast<size_t>(rng.next() * globals::nlines) % globals::nlines];  // exactly on a line
    }
    int next_trans = -1;
    const double v = rng.next();
    if (v < 0.2) {
      next_trans = static_cast<int>(rng.next() * globals::nlines);
    } else if (v < 0.25) {
      next_trans = globals::nlines + 1;
    } else if (v < 0.3) {
      next_trans = 0;
    }
    ct_nu[k] = nu;
    ct_next[k] = next_trans;
    ct_out[k] = closest_transition(nu, next_trans, linenu);
  }
  s.arr("kat.ct.nu", ct_nu.data(), NCT);
  s.arr("kat.ct.next_trans", ct_next.data(), NCT);
  s.arr("kat.ct.out", ct_out.data(), NCT);

  // ---- continuum opacity ----
  constexpr int NCHI = 2000;
  std::vector<double> chi_nu;
  std::vector<int> chi_cell;
  std::vector<double> chi_out;
  const int nc = grid::get_nonempty_npts_model();
  ContinuumOpacity chi{};
  for (int k = 0; k < NCHI; k++) {
    const int cell = static_cast<int>(rng.next() * nc) % nc;
    if (grid::thick_allcells[cell] == grid::CellThickness::THICK) {
      continue;
    }
    const double nu = std::exp(std::log(NU_MIN_R) + ((std::log(NU_MAX_R) - std::log(NU_MIN_R)) * rng.next()));
    chi.nonemptymgi = -1;  // force an evaluation
    calculate_chi_rpkt_cont<true>(nu, chi, cell);
    chi_nu.push_back(nu);
    chi_cell.push_back(cell);
    chi_out.push_back(chi.chi_escatter);
    chi_out.push_back(chi.chi_freefree_heat);
    chi_out.push_back(chi.chi_boundfree);
  }
  s.arr("kat.chi.nu", chi_nu.data(), static_cast<int64_t>(chi_nu.size()));
  s.arr("kat.chi.cell", chi_cell.data(), static_cast<int64_t>(chi_cell.size()));
  s.arr("kat.chi.out", chi_out.data(), static_cast<int64_t>(chi_out.size()));

  // ---- select_continuum_nu (ratecoeff.cc:563-638): free-bound emission frequencies ----
  // random continuum of the allcont list, T_e log-uniform over the LUT's temperature range, the draw taken from a
  // seeded generator exactly as a packet would take it (zrand = 1 - rng_uniform)
  constexpr int NSC = 600;
  std::vector<double> sc_in(static_cast<size_t>(NSC) * 2);
  std::vector<int> sc_cont(NSC);
  std::vector<double> sc_out(NSC);
  for (int k = 0; k < NSC; k++) {
    const int ci = static_cast<int>(rng.next() * globals::nbfcontinua) % globals::nbfcontinua;
    const auto T_e = static_cast<float>(std::exp(std::log(MINTEMP) + ((std::log(MAXTEMP) - std::log(MINTEMP)) * rng.next())));
    rngstate_type state(static_cast<std::uint32_t>(rng.next() * 4294967295.));
    if (k % 50 == 49) {
      // the extremes of the draw: uniform = 0 -> zrand = 1 (first piece) and uniform = 1 - 2^-24 -> zrand = 2^-24
      // (topmost piece, ratecoeff.cc:624-631); xoshiro128++ returns rotl(s0 + s3, 7) + s0
      const std::uint32_t extreme[2][4] = {{0U, 1U, 1U, 0U}, {0U, 1U, 1U, std::rotr(0xFFFFFF00U, 7)}};
      std::memcpy(&state, extreme[(k / 50) % 2], sizeof(state));
    }
    static_assert(sizeof(rngstate_type) == 16);
    rngstate_type peek = state;
    const double zrand = 1. - rng_uniform(peek);
    sc_in[(static_cast<size_t>(k) * 2) + 0] = T_e;
    sc_in[(static_cast<size_t>(k) * 2) + 1] = zrand;
    sc_cont[k] = ci;
    sc_out[k] = select_continuum_nu(globals::allcont.element[ci], globals::allcont.ion[ci], globals::allcont.level[ci],
                                    globals::allcont.phixstargetindex[ci], T_e, state);
  }
  // ... and the sampled DISTRIBUTION for two fixed (continuum, T_e) pairs: NKS reference draws each, for a two-sample
  // Kolmogorov-Smirnov test against draws made by the device function with independent random numbers
  constexpr int NKS = 5000;
  std::vector<double> ks_setup(4);
  std::vector<double> ks_out(static_cast<size_t>(2) * NKS);
  for (int c = 0; c < 2; c++) {
    const int ci = (c == 0) ? 0 : globals::nbfcontinua / 2;
    const auto T_e = static_cast<float>((c == 0) ? 8000. : 25000.);
    ks_setup[(c * 2) + 0] = ci;
    ks_setup[(c * 2) + 1] = T_e;
    rngstate_type state(static_cast<std::uint32_t>(1234567U + c));
    for (int k = 0; k < NKS; k++) {
      ks_out[(static_cast<size_t>(c) * NKS) + k] =
          select_continuum_nu(globals::allcont.element[ci], globals::allcont.ion[ci], globals::allcont.level[ci],
                              globals::allcont.phixstargetindex[ci], T_e, state);
    }
  }
  s.arr("kat.scks.setup", ks_setup.data(), 4);
  s.arr("kat.scks.out", ks_out.data(), static_cast<int64_t>(ks_out.size()));
  s.arr("kat.sc.in", sc_in.data(), static_cast<int64_t>(sc_in.size()));
  s.arr("kat.sc.cont", sc_cont.data(), NSC);
  s.arr("kat.sc.out", sc_out.data(), NSC);
}

// the reference's own per-cell cache tables (GPU_ON: one slot per cell, all filled up front)
template <class Sink>
void emit_reference_cellcache(Sink& s) {
  if constexpr (!cellcache_singleslot) {
    const int nc = grid::get_nonempty_npts_model();
    std::vector<double> pops;
    std::vector<double> marates;
    std::vector<double> matrans;
    std::vector<double> cooling;
    std::vector<double> nnlevel;
    std::vector<std::uint64_t> keepbits;
    std::vector<double> chiff;
    // ARTISB200_DUMP_CELLS=<n>: only n evenly spaced cells (bench-scale fixtures: one cell's tables are megabytes)
    const int ncells_wanted = std::atoi(env_or("ARTISB200_DUMP_CELLS", "0").c_str());
    std::vector<int> cells;
    if (ncells_wanted > 0 && ncells_wanted < nc) {
      for (int k = 0; k < ncells_wanted; k++) {
        cells.push_back(static_cast<int>((static_cast<long long>(k) * nc) / ncells_wanted) + (nc / (2 * ncells_wanted)));
      }
      s.arr("ref.cells", cells.data(), static_cast<int64_t>(cells.size()));
    } else {
      for (int cell = 0; cell < nc; cell++) {
        cells.push_back(cell);
      }
    }
    for (const int cell : cells) {
      const auto& slot = globals::cellcache.at(cell);
      pops.insert(pops.end(), slot.alllevels_pops.begin(), slot.alllevels_pops.end());
      marates.insert(marates.end(), slot.alllevels_maprocessrates.begin(), slot.alllevels_maprocessrates.end());
      matrans.insert(matrans.end(), slot.allmacroatomictransitions.begin(), slot.allmacroatomictransitions.end());
      cooling.insert(cooling.end(), slot.cooling_contrib.begin(), slot.cooling_contrib.end());
      nnlevel.insert(nnlevel.end(), slot.allcont_nnlevel.begin(), slot.allcont_nnlevel.end());
      keepbits.insert(keepbits.end(), slot.allcont_keepbits.begin(), slot.allcont_keepbits.end());
      chiff.push_back(slot.chi_ff_nnionpart[0]);
    }
    s.arr("ref.levelpops", pops.data(), static_cast<int64_t>(pops.size()));
    s.arr("ref.maprocessrates", marates.data(), static_cast<int64_t>(marates.size()));
    s.arr("ref.matrans", matrans.data(), static_cast<int64_t>(matrans.size()));
    s.arr("ref.cooling_contrib", cooling.data(), static_cast<int64_t>(cooling.size()));
    s.arr("ref.cont_nnlevel", nnlevel.data(), static_cast<int64_t>(nnlevel.size()));
    s.arr("ref.cont_keepbits", keepbits.data(), static_cast<int64_t>(keepbits.size()));
    s.arr("ref.chi_ff_nnionpart", chiff.data(), static_cast<int64_t>(chiff.size()));
  }
}
#endif

#ifdef ARTISB200_WITH_REFERENCE
// Known answers for the spectra / light-curve binning of the packets as update_packets leaves them:
// the reference's own add_to_spec_res / add_to_lc_res (spectrum_lightcurve.cc:544-713) for the angle-averaged bin with the
// emission / absorption decomposition, and flux + light curves for each of the MABINS direction bins.
template <class Sink>
void emit_reference_spectra(Sink& s, const std::span<const Packet> packets) {
  B200BinnedPackets b;
  b200_bin_escaped_packets(packets, -1, true, b);
  s.i64("ref.spec.nnubins", static_cast<int64_t>(MNUBINS));
  s.i64("ref.spec.mabins", static_cast<int64_t>(MABINS));
  s.i64("ref.spec.nprocs_exspec", static_cast<int64_t>(globals::nprocs_exspec));
  s.f64("ref.spec.nu_min", NU_MIN_R);
  s.f64("ref.spec.nu_max", NU_MAX_R);
  s.arr("ref.spec.lower_freq", b.lower_freq.data(), static_cast<int64_t>(b.lower_freq.size()));
  s.arr("ref.spec.delta_freq", b.delta_freq.data(), static_cast<int64_t>(b.delta_freq.size()));
  s.arr("ref.spec.flux", b.flux.data(), static_cast<int64_t>(b.flux.size()));
  s.arr("ref.spec.emission", b.emission.data(), static_cast<int64_t>(b.emission.size()));
  s.arr("ref.spec.trueemission", b.trueemission.data(), static_cast<int64_t>(b.trueemission.size()));
  s.arr("ref.spec.absorption", b.absorption.data(), static_cast<int64_t>(b.absorption.size()));
  s.arr("ref.lc.lum", b.lc_lum.data(), static_cast<int64_t>(b.lc_lum.size()));
  s.arr("ref.lc.lumcmf", b.lc_lumcmf.data(), static_cast<int64_t>(b.lc_lumcmf.size()));
  s.arr("ref.lc.gamma_lum", b.gamma_lc_lum.data(), static_cast<int64_t>(b.gamma_lc_lum.size()));
  s.arr("ref.lc.gamma_lumcmf", b.gamma_lc_lumcmf.data(), static_cast<int64_t>(b.gamma_lc_lumcmf.size()));
  std::vector<double> flux_res;
  std::vector<double> lum_res;
  std::vector<double> lumcmf_res;
  for (int dirbin = 0; dirbin < MABINS; dirbin++) {
    b200_bin_escaped_packets(packets, dirbin, false, b);
    flux_res.insert(flux_res.end(), b.flux.begin(), b.flux.end());
    lum_res.insert(lum_res.end(), b.lc_lum.begin(), b.lc_lum.end());
    lumcmf_res.insert(lumcmf_res.end(), b.lc_lumcmf.begin(), b.lc_lumcmf.end());
  }
  s.arr("ref.spec.flux_res", flux_res.data(), static_cast<int64_t>(flux_res.size()));
  s.arr("ref.lc.lum_res", lum_res.data(), static_cast<int64_t>(lum_res.size()));
  s.arr("ref.lc.lumcmf_res", lumcmf_res.data(), static_cast<int64_t>(lumcmf_res.size()));
  // the direction bin of every packet by the reference's get_escapedirectionbin (vectors.h:147-175): index parity
  std::vector<int> dirbins(packets.size());
  for (size_t i = 0; i < packets.size(); i++) {
    dirbins[i] = (packets[i].type == TYPE_ESCAPE) ? get_escapedirectionbin(packets[i].dir) : -1;
  }
  s.arr("ref.spec.dirbin", dirbins.data(), static_cast<int64_t>(dirbins.size()));

  // timing of the reference's binning on a bench-sized packet array (tools/bench_spectra.py): the packets replicated
  // ARTISB200_TIME_SPECTRA times, binned like write_partial_lightcurve_spectra does at the end of a multi-dimensional run
  // (spectrum_lightcurve.cc:316-337: dirbin -1 and then each of the MABINS direction bins, all packets every pass)
  if (const char* rep = std::getenv("ARTISB200_TIME_SPECTRA"); rep != nullptr) {
    const auto replicas = static_cast<size_t>(std::atol(rep));
    std::vector<Packet> many;
    many.reserve(replicas * packets.size());
    for (size_t r = 0; r < replicas; r++) {
      many.insert(many.end(), packets.begin(), packets.end());
    }
    const auto t0 = std::chrono::steady_clock::now();
    double checksum = 0.;
    for (int dirbin = -1; dirbin < MABINS; dirbin++) {
      b200_bin_escaped_packets(many, dirbin, false, b);
      checksum += b.lc_lum[0] + b.flux[0];
    }
    const auto wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printlnlog("ARTISB200_SPECTRA_TIMING npackets {} passes {} wall_s {:.6f} checksum {:g}", many.size(), 1 + MABINS, wall, checksum);
  }
}
#endif

#ifdef ARTISB200_WITH_REFERENCE
// Known answers for the LTE part of update_grid_cell (update_grid.cc:520-545): for every cell the
// reference's own radfield::get_T_J_from_J on a ladder of J values, and calculate_cellpartfuncts + calculate_ion_balance_nne
// (ltepop.cc:426-532) in Saha mode on the cell state as it is (temperatures, density, composition). The cell state is put back.
template <class Sink>
void emit_reference_lte_gridupdate(Sink& s) {
  const auto nc = static_cast<int64_t>(grid::get_nonempty_npts_model());
  const int nelements = get_nelements();
  const auto nions = static_cast<int64_t>(get_includedions());
  // (cell.elem_numberdens, the one input beyond the temperatures, density and composition, is part of emit_timestep_state)
  s.f64("ref.grid.mintemp", MINTEMP);
  s.f64("ref.grid.maxtemp", MAXTEMP);

  auto J = radfield::b200_J();
  std::vector<double> J_test(nc);
  std::vector<float> T_J(nc);
  for (int64_t cell = 0; cell < nc; cell++) {
    const double save = J[cell];
    const double temperature = 800. * std::pow(400., static_cast<double>(cell) / static_cast<double>(nc > 1 ? nc - 1 : 1));
    J_test[cell] = (cell % 17 == 5) ? std::numeric_limits<double>::infinity() : STEBO / PI * std::pow(temperature, 4.);
    J[cell] = J_test[cell];
    T_J[cell] = radfield::get_T_J_from_J(static_cast<int>(cell));
    J[cell] = save;
  }
  s.arr("ref.grid.J_test", J_test.data(), nc);
  s.arr("ref.grid.T_J_from_J", T_J.data(), nc);

  std::vector<float> save_nne(nc);
  std::vector<float> save_ground(grid::ion_groundlevelpops_allcells.data(), grid::ion_groundlevelpops_allcells.data() + (nc * nions));
  std::vector<float> save_partf(grid::ion_partfuncts_allcells.data(), grid::ion_partfuncts_allcells.data() + (nc * nions));
  std::vector<int> save_upper(nc * nelements);
  for (int64_t cell = 0; cell < nc; cell++) {
    save_nne[cell] = grid::get_nne(static_cast<int>(cell));
    for (int e = 0; e < nelements; e++) {
      save_upper[(cell * nelements) + e] = grid::get_elements_uppermost_ion(static_cast<int>(cell), e);
    }
  }
  const bool save_lte = globals::lte_iteration;
  globals::lte_iteration = true;
  std::vector<float> ref_nne(nc);
  std::vector<int> ref_upper(nc * nelements);
  const auto t0 = std::chrono::steady_clock::now();
  for (int64_t cell = 0; cell < nc; cell++) {
    for (int e = 0; e < nelements; e++) {
      calculate_cellpartfuncts(static_cast<int>(cell), e);
    }
    calculate_ion_balance_nne(static_cast<int>(cell));
    ref_nne[cell] = grid::get_nne(static_cast<int>(cell));
    for (int e = 0; e < nelements; e++) {
      ref_upper[(cell * nelements) + e] = grid::get_elements_uppermost_ion(static_cast<int>(cell), e);
    }
  }
  // machine-readable: the reference's own LTE update of all cells on one core (profiles/)
  printlnlog("ARTISB200_GRID_TIMING cells {} ions {} wall_s {:.6f}", nc, nions,
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  s.arr("ref.grid.nne", ref_nne.data(), nc);
  s.arr("ref.grid.uppermost_ion", ref_upper.data(), nc * nelements);
  s.arr("ref.grid.ion_partfuncts", grid::ion_partfuncts_allcells.data(), nc * nions);
  s.arr("ref.grid.ion_groundlevelpops", grid::ion_groundlevelpops_allcells.data(), nc * nions);

  // the same on a temperature ladder from 60 K to 150 000 K over the cells (densities and compositions as they are): the cold
  // end overflows the Saha factors, which truncates the ion lists (ltepop.cc:343-352) and, at the very bottom, leaves only
  // the lowest ion stages (set_groundlevelpops_neutral, ltepop.cc:254-278)
  std::vector<float> save_Te(grid::Te_allcells.data(), grid::Te_allcells.data() + nc);
  std::vector<float> save_TJ(grid::TJ_allcells.data(), grid::TJ_allcells.data() + nc);
  std::vector<float> ladder(nc);
  for (int64_t cell = 0; cell < nc; cell++) {
    ladder[cell] = static_cast<float>(60. * std::pow(2500., static_cast<double>(cell) / static_cast<double>(nc > 1 ? nc - 1 : 1)));
    grid::Te_allcells[cell] = ladder[cell];
    grid::TJ_allcells[cell] = ladder[cell];
  }
  for (int64_t cell = 0; cell < nc; cell++) {
    for (int e = 0; e < nelements; e++) {
      calculate_cellpartfuncts(static_cast<int>(cell), e);
    }
    calculate_ion_balance_nne(static_cast<int>(cell));
    ref_nne[cell] = grid::get_nne(static_cast<int>(cell));
    for (int e = 0; e < nelements; e++) {
      ref_upper[(cell * nelements) + e] = grid::get_elements_uppermost_ion(static_cast<int>(cell), e);
    }
  }
  s.arr("ref.grid.ladder_T", ladder.data(), nc);
  s.arr("ref.grid.ladder_nne", ref_nne.data(), nc);
  s.arr("ref.grid.ladder_uppermost_ion", ref_upper.data(), nc * nelements);
  s.arr("ref.grid.ladder_ion_partfuncts", grid::ion_partfuncts_allcells.data(), nc * nions);
  s.arr("ref.grid.ladder_ion_groundlevelpops", grid::ion_groundlevelpops_allcells.data(), nc * nions);
  std::copy(save_Te.begin(), save_Te.end(), grid::Te_allcells.data());
  std::copy(save_TJ.begin(), save_TJ.end(), grid::TJ_allcells.data());
  globals::lte_iteration = save_lte;
  std::copy(save_ground.begin(), save_ground.end(), grid::ion_groundlevelpops_allcells.data());
  std::copy(save_partf.begin(), save_partf.end(), grid::ion_partfuncts_allcells.data());
  for (int64_t cell = 0; cell < nc; cell++) {
    grid::set_nne(static_cast<int>(cell), save_nne[cell]);
    for (int e = 0; e < nelements; e++) {
      grid::set_elements_uppermost_ion(static_cast<int>(cell), e, save_upper[(cell * nelements) + e]);
    }
  }
}
#endif

auto dump_requested(const int nts) -> bool {
  const char* dir = std::getenv("ARTISB200_DUMP_DIR");
  if (dir == nullptr) {
    return false;
  }
  const auto list = env_or("ARTISB200_DUMP_TS", "all");
  if (list == "all") {
    return true;
  }
  const std::string needle = "," + std::to_string(nts) + ",";
  return ("," + list + ",").find(needle) != std::string::npos;
}

}  // anonymous namespace

void update_packets(const int nts, std::span<Packet> packets) {
#ifdef ARTISB200_WITH_REFERENCE
  const auto mode = env_or("ARTISB200_MODE", "ref");
#else
  const auto mode = env_or("ARTISB200_MODE", "gpu");
#endif
  const bool dump = dump_requested(nts);
  const auto dumpdir = env_or("ARTISB200_DUMP_DIR", ".");
  if (dump) {
    static bool static_written = false;
    if (!static_written) {
      b200::SnapshotWriter w(dumpdir + "/static.abt");
      emit_static(w);
      static_written = true;
    }
    b200::SnapshotWriter w(dumpdir + "/ts" + std::to_string(nts) + "_before.abt");
    emit_timestep_state(w, nts);
    emit_packets(w, packets);
    emit_estimators(w, nts);  // the pre-existing (normally zero) values, so that "after - before" is this call's work
#ifdef ARTISB200_WITH_REFERENCE
    if (std::getenv("ARTISB200_DUMP_GRID") != nullptr) {
      emit_reference_lte_gridupdate(w);
    }
#endif
  }

  const auto t0 = std::chrono::steady_clock::now();
  if (mode == "gpu") {
    update_packets_gpu(nts, packets);
#ifdef ARTISB200_WITH_REFERENCE
  } else if (mode == "ref") {
    update_packets_reference_impl(nts, packets);
  } else if (mode == "ref_perpacket") {
    update_packets_reference_perpacket(nts, packets);
#endif
  } else {
    printlnlog("[fatal] artis_b200: unknown ARTISB200_MODE '{}' for this build", mode);
    std::abort();
  }
  const auto wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // machine-readable line used by bench.py's cpu_baseline / reference arm
  printlnlog("ARTISB200_TIMING nts {} mode {} npackets {} wall_s {:.6f} interactions {}", nts, mode, packets.size(), wall,
             stats::get_counter(stats::Counter::INTERACTIONS));

  if (dump) {
    b200::SnapshotWriter w(dumpdir + "/ts" + std::to_string(nts) + "_after.abt");
    emit_packets(w, packets);
    emit_estimators(w, nts);
#ifdef ARTISB200_WITH_REFERENCE
    if (mode == "ref_perpacket") {
      emit_reference_cellcache(w);
      emit_reference_kats(w, nts);
    }
    if (std::getenv("ARTISB200_DUMP_SPECTRA") != nullptr) {
      emit_reference_spectra(w, packets);
    }
#endif
  }
  MPI_Barrier_allranks();  // update_packets.cc:631
}
