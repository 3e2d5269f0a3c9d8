// Spectra and light curves of the escaped packets, binned where the packets live.
//
// The reference bins on the host: write_partial_lightcurve_spectra (spectrum_lightcurve.cc:316-337) walks ALL packets once
// for the angle-averaged result and once more for each of the MABINS = 100 direction bins, every pass filtering with
// get_escapedirectionbin (vectors.h:147-175), and exspec (exspec.cc:60-200) does the same over the packet files. Here a
// packet is read ONCE: its direction bin is evaluated once and it is added to the angle-averaged set and to the set of its
// own direction bin (the sets lie one behind the other, set 0 = angle-averaged, set 1 + dirbin). Per packet this follows
// add_to_spec_res (spectrum_lightcurve.cc:544-661) and add_to_lc_res (691-718) in the reference's operation order, so that
// every addend is the reference's to the last bit where only IEEE arithmetic is involved (bin indices go through log / acos:
// a packet within an ulp of a bin edge may fall on the other side, like with any other libm).
//
// HBM traffic per escaped r-packet: kinematics (64 B) + energies (64 B) + type record (32-byte sector) + escape_type /
// escape_time (8 B) and, with the emission / absorption decomposition, both emission records' types and the absorption
// fields (~28 B): ~200 B read, against 4 (8 with direction bins) f64 read-modify-writes into tables that stay in L2.
#pragma once
#include "hd.h"
#include "tables.h"
#include "vec.h"

namespace ab {

constexpr int NPHIBINS = 10;       // exspec.h:10
constexpr int NCOSTHETABINS = 10;  // exspec.h:11
constexpr int MABINS = NPHIBINS * NCOSTHETABINS;
constexpr double PARSEC = 3.0857e+18;  // constants.h:39

// One set = one observer direction bin (-1 first). Layouts are the reference's (Spectra, spectrum_lightcurve.h:15-37):
//   flux[set][nnu][nts], emission / trueemission[set][nnu][nts][proccount], absorption[set][nnu][nts][ioncount]
struct SpectraView {
  int nnubins;     // MNUBINS (exspec.h:8)
  int ntimesteps;
  int nsets;       // 1 (angle-averaged only) or 1 + MABINS
  int nsets_emabs; // sets that carry the emission / absorption decomposition: 0, 1 or 1 + MABINS
  int proccount;   // 2 * nelements * max_nions + 1 (spectrum_lightcurve.cc:167)
  int ioncount;    // nelements * max_nions
  int max_nions;
  int nbflist;
  double nu_min;
  double nu_max;
  double dlognu;
  double tmin;
  double tmax;
  double vmax;
  double nprocs_exspec;
  const double* ts_start;   // [ntimesteps]
  const double* ts_width;   // [ntimesteps]
  const float* delta_freq;  // [nnubins]
  const int* line_elementindex;
  const int* line_ionindex;
  const int* bflist_element;  // [nbflist] element / ion of the continuum behind a bound-free emission type
  const int* bflist_ion;      // (globals::bflist, indexed by -1 - emissiontype: atomic.h:513-531)
  double* flux;
  double* emission;
  double* trueemission;
  double* absorption;
  double* lc_lum;        // [set][nts]
  double* lc_lumcmf;     // [set][nts]
  double* gamma_lc_lum;  // [nts]   (escaped gamma packets: angle-averaged only, spectrum_lightcurve.cc:284)
  double* gamma_lc_lumcmf;
  // optional (null = off): Stokes Q and U spectra with the layouts of flux / emission / absorption (exspec.cc:52-59, POL_ON:
  // every addend of the I arrays times the packet's stokes_q / stokes_u, spectrum_lightcurve.cc:567-624), and the spectrum of
  // the escaped gamma packets on its own frequency grid, angle-averaged only (exspec.cc:61-64, 83-86)
  double* flux_q;
  double* flux_u;
  double* emission_q;
  double* emission_u;
  double* absorption_q;
  double* absorption_u;
  double* gamma_flux;             // [nnubins][nts]
  const float* gamma_delta_freq;  // [nnubins]
  double gamma_nu_min;
  double gamma_nu_max;
  double gamma_dlognu;
  int* dirbin;           // [npackets] direction bin of every escaped packet, -1 for the others (or null)
};

// block-local light curves of the angle-averaged set (every escaped packet adds to one of ntimesteps addresses)
struct LcLocal {
  double* lum;
  double* lumcmf;
  double* gamma_lum;
  double* gamma_lumcmf;
};

// vectors.h:147-175
AHD int escape_direction_bin(const double* dir_in) {
  const double syn_dir[3] = {0., 0., 1.};  // constants.h:94
  const double xhat[3] = {1., 0., 0.};
  const double dirmag = vec_len3(dir_in);
  const double dir[3] = {dir_in[0] / dirmag, dir_in[1] / dirmag, dir_in[2] / dirmag};
  const double costheta = dot3(dir, syn_dir);
  int costhetabin = static_cast<int>((costheta + 1.0) * NCOSTHETABINS / 2.0);
  costhetabin = (costhetabin < 0) ? 0 : ((costhetabin > NCOSTHETABINS - 1) ? NCOSTHETABINS - 1 : costhetabin);
  double vec1[3];
  double vec2[3];
  double vec3[3];
  cross_prod(dir, syn_dir, vec1);
  cross_prod(xhat, syn_dir, vec2);
  const double vec1_len = vec_len3(vec1);
  double cosphi = 1.0;
  if (vec1_len > 1e-12) {
    cosphi = dot3(vec1, vec2) / vec1_len;
    cosphi = (cosphi < -1.0) ? -1.0 : ((1.0 < cosphi) ? 1.0 : cosphi);  // std::clamp
  }
  cross_prod(vec2, syn_dir, vec3);
  const double testphi = dot3(vec1, vec3);
  const double phi = (testphi > 0) ? acos(cosphi) : acos(cosphi) + PI;
  int phibin = static_cast<int>(phi / 2. / PI * NPHIBINS);
  phibin = (phibin < 0) ? 0 : ((phibin > NPHIBINS - 1) ? NPHIBINS - 1 : phibin);
  return (costhetabin * NPHIBINS) + phibin;
}

// spectrum_lightcurve.cc:205-217: the timestep with start <= time < next start (tmax after the last one), found by
// bisection over the ascending starts instead of the reference's scan (same index). Caller guarantees tmin < time < tmax.
AHD int spectra_timestep(const SpectraView& S, const double time) {
  int lo = 0;
  int hi = S.ntimesteps - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (time >= S.ts_start[mid]) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  return lo;
}

// sn3d.h:134-137
AHD int log_bin_index(const double value, const double minvalue, const double dlog, const int nbins) {
  const double x = floor((log(value) - log(minvalue)) / dlog);
  const long long i = static_cast<long long>(x);
  return static_cast<int>((i < 0) ? 0 : ((i > nbins - 1) ? nbins - 1 : i));
}

// spectrum_lightcurve.cc:169-203 (its assert_always sites report through `bad`)
AHD int emission_column(const SpectraView& S, const int et, bool& bad) {
  if (et >= 0) {
    return (S.line_elementindex[et] * S.max_nions) + S.line_ionindex[et];
  }
  if (et == EMTYPE_FREEFREE) {
    return 2 * S.ioncount;
  }
  if (et == EMTYPE_NOTSET) {
    return -1;
  }
  const int bfindex = -1 - et;
  if (S.nbflist == 0) {
    return 2 * S.ioncount;
  }
  if (bfindex >= S.nbflist) {
    bad = true;
    return -1;
  }
  return S.ioncount + (S.bflist_element[bfindex] * S.max_nions) + S.bflist_ion[bfindex];
}

// Everything one packet adds. `local` (or null): block-local light curves of set 0.
AHD void bin_escaped_packet(const Tables& T, const SpectraView& S, const long long ip, const LcLocal* local) {
  const HotC& hc = T.pkt.hc[ip];
  if (hc.type != TYPE_ESCAPE) {
    if (S.dirbin != nullptr) {
      S.dirbin[ip] = -1;
    }
    return;
  }
  const int escape_type = T.pkt.escape_type[ip];
  const bool is_rpkt = (escape_type == TYPE_RPKT);
  if (S.dirbin == nullptr && !is_rpkt && escape_type != TYPE_GAMMA) {
    return;
  }
  const HotA ha = T.pkt.ha[ip];
  const int dirbin = (S.nsets > 1 || S.dirbin != nullptr) ? escape_direction_bin(ha.dir) : -1;
  if (S.dirbin != nullptr) {
    S.dirbin[ip] = dirbin;
  }
  if (!is_rpkt && escape_type != TYPE_GAMMA) {
    return;
  }
  const HotB& hb = T.pkt.hb[ip];
  const double e_rf = hb.e_rf;
  const double e_cmf = hb.e_cmf;
  const double nu_rf = hb.nu_rf;
  const double escape_time = static_cast<double>(T.pkt.escape_time[ip]);
  const double mabins = static_cast<double>(MABINS);

  // add_to_lc_res (spectrum_lightcurve.cc:691-718)
  const double t_arrive = escape_time - (dot3(ha.pos, ha.dir) / CLIGHT_PROP);
  const bool arrives = (t_arrive > S.tmin && t_arrive < S.tmax);
  int nts = -1;
  if (arrives) {
    nts = spectra_timestep(S, t_arrive);
    const double base = e_rf / S.ts_width[nts];
    if (is_rpkt) {
      atomic_add((local != nullptr) ? &local->lum[nts] : &S.lc_lum[nts], base * 1. / S.nprocs_exspec);
      if (S.nsets > 1) {
        atomic_add(&S.lc_lum[((1 + dirbin) * static_cast<long long>(S.ntimesteps)) + nts], base * mabins / S.nprocs_exspec);
      }
    } else {
      atomic_add((local != nullptr) ? &local->gamma_lum[nts] : &S.gamma_lc_lum[nts], base * 1. / S.nprocs_exspec);
    }
  }
  const double inverse_gamma = sqrt(1. - (S.vmax * S.vmax / CLIGHTSQUARED));
  const double t_escape_cmf = escape_time * inverse_gamma;
  if (t_escape_cmf > S.tmin && t_escape_cmf < S.tmax) {
    const int nts_cmf = spectra_timestep(S, t_escape_cmf);
    const double base = e_cmf / S.ts_width[nts_cmf];
    if (is_rpkt) {
      atomic_add((local != nullptr) ? &local->lumcmf[nts_cmf] : &S.lc_lumcmf[nts_cmf], base * 1. / S.nprocs_exspec / inverse_gamma);
      if (S.nsets > 1) {
        atomic_add(&S.lc_lumcmf[((1 + dirbin) * static_cast<long long>(S.ntimesteps)) + nts_cmf],
                   base * mabins / S.nprocs_exspec / inverse_gamma);
      }
    } else {
      atomic_add((local != nullptr) ? &local->gamma_lumcmf[nts_cmf] : &S.gamma_lc_lumcmf[nts_cmf],
                 base * 1. / S.nprocs_exspec / inverse_gamma);
    }
  }
  if (!is_rpkt) {
    // add_to_spec_res for the gamma-ray spectrum (exspec.cc:83-86): flux only, angle-averaged only
    if (S.gamma_flux != nullptr && arrives && nu_rf > S.gamma_nu_min && nu_rf < S.gamma_nu_max) {
      const int nnu = log_bin_index(nu_rf, S.gamma_nu_min, S.gamma_dlognu, S.nnubins);
      const double deltaE = e_rf / S.ts_width[nts] / static_cast<double>(S.gamma_delta_freq[nnu]) / 4.e12 / PI / PARSEC / PARSEC /
                            S.nprocs_exspec * 1.;
      atomic_add(&S.gamma_flux[(nnu * static_cast<long long>(S.ntimesteps)) + nts], deltaE);
    }
    return;
  }

  // add_to_spec_res (spectrum_lightcurve.cc:544-661)
  if (!(arrives && nu_rf > S.nu_min && nu_rf < S.nu_max)) {
    return;
  }
  const int nnu = log_bin_index(nu_rf, S.nu_min, S.dlognu, S.nnubins);
  const double width = S.ts_width[nts];
  const double deltaE_unit = e_rf / width / static_cast<double>(S.delta_freq[nnu]) / 4.e12 / PI / PARSEC / PARSEC / S.nprocs_exspec;
  const long long nt = S.ntimesteps;
  const long long fluxindex = (nnu * nt) + nts;
  const long long fluxsize = static_cast<long long>(S.nnubins) * nt;
  atomic_add(&S.flux[fluxindex], deltaE_unit * 1.);
  if (S.nsets > 1) {
    atomic_add(&S.flux[((1 + dirbin) * fluxsize) + fluxindex], deltaE_unit * mabins);
  }
  const bool stokes = (S.flux_q != nullptr);
  const double stokes_q = stokes ? hb.stokes_q : 0.;
  const double stokes_u = stokes ? hb.stokes_u : 0.;
  if (stokes) {
    atomic_add(&S.flux_q[fluxindex], stokes_q * (deltaE_unit * 1.));
    atomic_add(&S.flux_u[fluxindex], stokes_u * (deltaE_unit * 1.));
    if (S.nsets > 1) {
      atomic_add(&S.flux_q[((1 + dirbin) * fluxsize) + fluxindex], stokes_q * (deltaE_unit * mabins));
      atomic_add(&S.flux_u[((1 + dirbin) * fluxsize) + fluxindex], stokes_u * (deltaE_unit * mabins));
    }
  }
  if (S.nsets_emabs == 0) {
    return;
  }
  const bool emabs_res = (S.nsets_emabs > 1);
  bool bad = false;
  const long long emsize = fluxsize * S.proccount;
  const long long emindex_base = (nnu * nt * S.proccount) + (static_cast<long long>(nts) * S.proccount);
  const int truenproc = emission_column(S, T.pkt.trueem[ip].type, bad);
  if (truenproc >= 0) {
    atomic_add(&S.trueemission[emindex_base + truenproc], deltaE_unit * 1.);
    if (emabs_res) {
      atomic_add(&S.trueemission[((1 + dirbin) * emsize) + emindex_base + truenproc], deltaE_unit * mabins);
    }
  }
  const int nproc = emission_column(S, T.pkt.em[ip].type, bad);
  if (nproc >= 0) {
    atomic_add(&S.emission[emindex_base + nproc], deltaE_unit * 1.);
    if (emabs_res) {
      atomic_add(&S.emission[((1 + dirbin) * emsize) + emindex_base + nproc], deltaE_unit * mabins);
    }
    if (stokes) {
      atomic_add(&S.emission_q[emindex_base + nproc], stokes_q * (deltaE_unit * 1.));
      atomic_add(&S.emission_u[emindex_base + nproc], stokes_u * (deltaE_unit * 1.));
      if (emabs_res) {
        atomic_add(&S.emission_q[((1 + dirbin) * emsize) + emindex_base + nproc], stokes_q * (deltaE_unit * mabins));
        atomic_add(&S.emission_u[((1 + dirbin) * emsize) + emindex_base + nproc], stokes_u * (deltaE_unit * mabins));
      }
    }
  }
  if (bad) {
    // spectrum_lightcurve.cc:197 assert_always(bfindex < globals::nbfcontinua)
#if defined(__CUDA_ARCH__)
    if (atomicCAS(reinterpret_cast<unsigned long long*>(&T.dev_error[0]), 0ULL,
                  static_cast<unsigned long long>(DEVERR_SPECTRA_EMISSIONTYPE)) == 0ULL) {
      T.dev_error[1] = ip;
    }
#else
    if (T.dev_error[0] == 0) {
      T.dev_error[0] = DEVERR_SPECTRA_EMISSIONTYPE;
      T.dev_error[1] = ip;
    }
#endif
  }
  const double absorptionfreq = T.pkt.absorptionfreq[ip];
  if (absorptionfreq > S.nu_min && absorptionfreq < S.nu_max) {
    const int nnu_abs = log_bin_index(absorptionfreq, S.nu_min, S.dlognu, S.nnubins);
    const double deltaE_abs_unit =
        e_rf / width / static_cast<double>(S.delta_freq[nnu_abs]) / 4.e12 / PI / PARSEC / PARSEC / S.nprocs_exspec;
    const int at = T.pkt.absorptiontype[ip];
    if (at >= 0) {
      const long long abssize = fluxsize * S.ioncount;
      const long long absindex = (nnu_abs * nt * S.ioncount) + (static_cast<long long>(nts) * S.ioncount) +
                                 (S.line_elementindex[at] * S.max_nions) + S.line_ionindex[at];
      atomic_add(&S.absorption[absindex], deltaE_abs_unit * 1.);
      if (emabs_res) {
        atomic_add(&S.absorption[((1 + dirbin) * abssize) + absindex], deltaE_abs_unit * mabins);
      }
      if (stokes) {
        atomic_add(&S.absorption_q[absindex], stokes_q * (deltaE_abs_unit * 1.));
        atomic_add(&S.absorption_u[absindex], stokes_u * (deltaE_abs_unit * 1.));
        if (emabs_res) {
          atomic_add(&S.absorption_q[((1 + dirbin) * abssize) + absindex], stokes_q * (deltaE_abs_unit * mabins));
          atomic_add(&S.absorption_u[((1 + dirbin) * abssize) + absindex], stokes_u * (deltaE_abs_unit * mabins));
        }
      }
    }
  }
}

}  // namespace ab
