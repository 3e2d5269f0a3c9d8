// Integration shim: reference spectrum_lightcurve.cc + an accessor that bins a span of packets with the reference's own
// add_to_lc_res / add_to_spec_res (spectrum_lightcurve.cc:544-713) exactly as write_partial_lightcurve_spectra_dirbin
// does (277-287), and hands the arrays back instead of formatting them into spec.out / light_curve.out.
#include "spectrum_lightcurve.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

void b200_bin_escaped_packets(std::span<const Packet> pkts, const int dirbin, const bool do_emission_absorption,
                              B200BinnedPackets& out) {
  static Spectra spectra;
  init_spectra(spectra, NU_MIN_R, NU_MAX_R, do_emission_absorption);
  out.lc_lum.assign(globals::ntimesteps, 0.);
  out.lc_lumcmf.assign(globals::ntimesteps, 0.);
  out.gamma_lc_lum.assign(globals::ntimesteps, 0.);
  out.gamma_lc_lumcmf.assign(globals::ntimesteps, 0.);
  for (const auto& pkt : pkts) {
    if (pkt.type != TYPE_ESCAPE) {
      continue;
    }
    if (pkt.escape_type == TYPE_RPKT) {
      add_to_lc_res(pkt, dirbin, out.lc_lum, out.lc_lumcmf);
      add_to_spec_res(pkt, dirbin, spectra, nullptr, nullptr);
    } else if (KEEP_ESCAPED_GAMMAS && dirbin == -1 && pkt.escape_type == TYPE_GAMMA) {
      add_to_lc_res(pkt, dirbin, out.gamma_lc_lum, out.gamma_lc_lumcmf);
    }
  }
  out.lower_freq.assign(spectra.lower_freq.begin(), spectra.lower_freq.end());
  out.delta_freq.assign(spectra.delta_freq.begin(), spectra.delta_freq.end());
  const auto copy = [](const MPI_shared_array<double>& src, std::vector<double>& dst) {
    dst.assign(src.data(), src.data() + src.size());
  };
  copy(spectra.fluxalltimesteps, out.flux);
  out.emission.clear();
  out.trueemission.clear();
  out.absorption.clear();
  if (do_emission_absorption) {
    copy(spectra.emissionalltimesteps, out.emission);
    copy(spectra.trueemissionalltimesteps, out.trueemission);
    copy(spectra.absorptionalltimesteps, out.absorption);
  }
}
