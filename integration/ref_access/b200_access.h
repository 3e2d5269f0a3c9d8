// Accessors added around the reference translation units by integration/ref_access/ref_*.cc.
#pragma once
#include <array>
#include <cstddef>
#include <span>
#include <vector>

#include "constants.h"

namespace grid {
auto b200_ncoordgrid() -> std::array<int, 3>;
auto b200_coord_pos_min_tmin(int axis) -> std::span<const double>;
auto b200_propcell_nonemptymgi() -> std::span<const int>;
auto b200_propgridtype() -> GridType;
}  // namespace grid

auto b200_lut_spontrecombcoeffs() -> std::span<const double>;
auto b200_lut_corrphotoioncoeffs() -> std::span<const double>;
auto b200_lut_bfcooling_coeffs() -> std::span<const double>;
auto b200_lut_temperature_grid() -> std::span<const double>;

auto b200_expansionopacity_planck_cumulative() -> std::span<const double>;

namespace gammapkt {
// XCOM photoionisation tables of Z = 1..100 flattened: rows [zstart[Z-1], zstart[Z]) of energy [MeV] / sigma [cm^2]
void b200_xcom_tables(std::vector<int>& zstart, std::vector<double>& energy, std::vector<double>& sigma);
}  // namespace gammapkt

namespace kpkt {
auto b200_coolinglist_type(int i) -> int;
auto b200_coolinglist_level(int i) -> int;
auto b200_coolinglist_phixstargetindex(int i) -> int;
}  // namespace kpkt

namespace radfield {
auto b200_J() -> std::span<double>;
auto b200_nuJ() -> std::span<double>;
auto b200_bins_J_raw() -> std::span<double>;
auto b200_bins_nuJ_raw() -> std::span<double>;
auto b200_bfrate_raw() -> std::span<double>;
auto b200_prev_bfrate_normed() -> std::span<const float>;
auto b200_bin_solutions_W() -> std::span<const float>;
auto b200_bin_solutions_T_R() -> std::span<const float>;
}  // namespace radfield

namespace nonthermal {
void b200_nt_cell_state(std::vector<double>& ion_ratecoeff, std::vector<double>& ion_energyrate, std::vector<float>& prob_num_auger,
                        std::vector<float>& ionenfrac_num_auger, std::vector<float>& frac_ionisation);
void b200_nt_excitations(int& stride, std::vector<int>& count, std::vector<int>& alltransindex, std::vector<double>& frac_deposition,
                         std::vector<double>& ratecoeffperdeposition, std::vector<double>& deposition_rate_density,
                         std::vector<float>& frac_excitation);
}  // namespace nonthermal

// spectra and light curves of a span of packets for one direction bin (-1 = angle-averaged), binned by the reference's own
// add_to_lc_res / add_to_spec_res (ref_spectrum_lightcurve.cc); only linked into the oracle build's snapshot hooks
struct Packet;
struct B200BinnedPackets {
  std::vector<float> lower_freq, delta_freq;
  std::vector<double> flux, emission, trueemission, absorption;
  std::vector<double> lc_lum, lc_lumcmf, gamma_lc_lum, gamma_lc_lumcmf;
};
void b200_bin_escaped_packets(std::span<const Packet> pkts, int dirbin, bool do_emission_absorption, B200BinnedPackets& out);

namespace stats {
void b200_add_counter(int i, std::ptrdiff_t n);
}  // namespace stats
