// Integration shim: reference nonthermal.cc + an accessor that evaluates, per cell and ion, what the non-thermal
// routing of the packet path reads during update_packets (nonthermal.cc:1172-1183, 1509-1580, 2398-2492): the
// non-thermal ionisation rate coefficient, the energy rate going to ionising each ion, the Auger channel
// probabilities and the fraction of the deposition going to ionisation. They are per-timestep cell state, fixed
// while packets propagate.
#include "nonthermal.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace nonthermal {
void b200_nt_cell_state(std::vector<double>& ion_ratecoeff, std::vector<double>& ion_energyrate, std::vector<float>& prob_num_auger,
                        std::vector<float>& ionenfrac_num_auger, std::vector<float>& frac_ionisation) {
  const ptrdiff_t nc = grid::get_nonempty_npts_model();
  const ptrdiff_t nions = get_includedions();
  constexpr int NA = NT_MAX_AUGER_ELECTRONS + 1;
  ion_ratecoeff.assign(nc * nions, 0.);
  ion_energyrate.assign(nc * nions, 0.);
  prob_num_auger.assign(nc * nions * NA, 0.F);
  ionenfrac_num_auger.assign(nc * nions * NA, 0.F);
  frac_ionisation.assign(nc, 0.F);
  if constexpr (!NT_ON) {
    return;
  }
  for (ptrdiff_t nonemptymgi = 0; nonemptymgi < nc; nonemptymgi++) {
    frac_ionisation[nonemptymgi] = get_nt_frac_ionisation(static_cast<int>(nonemptymgi));
    for (int element = 0; element < get_nelements(); element++) {
      for (int ion = 0; ion < get_nions(element); ion++) {
        const ptrdiff_t u = get_uniqueionindex(element, ion);
        if (ion < get_nions(element) - 1) {
          ion_ratecoeff[(nonemptymgi * nions) + u] = nt_ionisation_ratecoeff(static_cast<int>(nonemptymgi), element, ion);
          ion_energyrate[(nonemptymgi * nions) + u] = ion_ntion_energyrate(static_cast<int>(nonemptymgi), element, ion);
        }
        if constexpr (NT_SOLVE_SPENCERFANO) {
          const auto& d = get_cell_allions_data(nonemptymgi)[u];
          for (int a = 0; a < NA; a++) {
            prob_num_auger[(((nonemptymgi * nions) + u) * NA) + a] = d.prob_num_auger[a];
            ionenfrac_num_auger[(((nonemptymgi * nions) + u) * NA) + a] = d.ionenfrac_num_auger[a];
          }
        }
      }
    }
  }
}

// the per-cell lists of non-thermal excitation transitions (nonthermal.cc:202-212, 364-367), flattened with the
// reference's own stride, and what nt_excitation_ratecoeff() multiplies them with (nonthermal.cc:2494-2518)
void b200_nt_excitations(int& stride, std::vector<int>& count, std::vector<int>& alltransindex, std::vector<double>& frac_deposition,
                         std::vector<double>& ratecoeffperdeposition, std::vector<double>& deposition_rate_density,
                         std::vector<float>& frac_excitation) {
  const ptrdiff_t nc = grid::get_nonempty_npts_model();
  stride = (NT_ON && NT_SOLVE_SPENCERFANO && NT_EXCITATION_ON) ? nt_excitations_stored : 0;
  count.assign(nc, 0);
  deposition_rate_density.assign(nc, 0.);
  frac_excitation.assign(nc, 0.F);
  alltransindex.assign(nc * stride, -1);
  frac_deposition.assign(nc * stride, 0.);
  ratecoeffperdeposition.assign(nc * stride, 0.);
  if constexpr (!NT_ON) {
    return;
  }
  for (ptrdiff_t nonemptymgi = 0; nonemptymgi < nc; nonemptymgi++) {
    deposition_rate_density[nonemptymgi] = ntlepton_deposition_rate_density_all_cells[nonemptymgi];
    frac_excitation[nonemptymgi] = get_nt_frac_excitation(static_cast<int>(nonemptymgi));
    if (stride > 0) {
      const auto list = get_cell_ntexcitations(nonemptymgi);
      count[nonemptymgi] = static_cast<int>(list.size());
      for (size_t k = 0; k < list.size(); k++) {
        alltransindex[(nonemptymgi * stride) + k] = list[k].alltransindex;
        frac_deposition[(nonemptymgi * stride) + k] = list[k].frac_deposition;
        ratecoeffperdeposition[(nonemptymgi * stride) + k] = list[k].ratecoeffperdeposition;
      }
    }
  }
}
}  // namespace nonthermal
