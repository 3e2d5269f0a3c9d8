// Conversion of one packet between the reference's AoS Packet (240 bytes, or 256 with the GPU_ON rngstate
// prefix; layout in packet.h / SURVEY.md Appendix A) and the device SoA arrays. Every field is carried, so a
// round trip is lossless and the host's packets*.out / checkpoint writers (packet.cc:226-311) see exactly
// what the reference would have produced.
#pragma once
#include <cstring>

#include "hd.h"
#include "packet.h"
#include "tables.h"

namespace ab {

template <class U>
AHD U rd(const unsigned char* base, const int off) {
  U v;
  memcpy(&v, base + off, sizeof(U));
  return v;
}

template <class U>
AHD void wr(unsigned char* base, const int off, const U v) {
  memcpy(base + off, &v, sizeof(U));
}

AHD void aos_to_soa_one(const Tables& T, const unsigned char* aos, const int stride, const long long i) {
  const unsigned char* rec = aos + (i * stride);
  const int b = stride - AosLayout::size;  // 0 or 16
  const unsigned char* q = rec + b;
  const PacketSoA& s = T.pkt;
  using L = AosLayout;
  s.prop_time[i] = rd<double>(q, L::prop_time);
  s.pos_x[i] = rd<double>(q, L::pos);
  s.pos_y[i] = rd<double>(q, L::pos + 8);
  s.pos_z[i] = rd<double>(q, L::pos + 16);
  s.dir_x[i] = rd<double>(q, L::dir);
  s.dir_y[i] = rd<double>(q, L::dir + 8);
  s.dir_z[i] = rd<double>(q, L::dir + 16);
  s.nu_cmf[i] = rd<double>(q, L::nu_cmf);
  s.e_cmf[i] = rd<double>(q, L::e_cmf);
  s.nu_rf[i] = rd<double>(q, L::nu_rf);
  s.e_rf[i] = rd<double>(q, L::e_rf);
  s.next_trans[i] = rd<int>(q, L::next_trans);
  s.nscatterings[i] = rd<int>(q, L::nscatterings);
  s.emissiontype[i] = rd<int>(q, L::emissiontype);
  s.em_pos_x[i] = rd<double>(q, L::em_pos);
  s.em_pos_y[i] = rd<double>(q, L::em_pos + 8);
  s.em_pos_z[i] = rd<double>(q, L::em_pos + 16);
  s.em_time[i] = rd<float>(q, L::em_time);
  s.absorptiontype[i] = rd<int>(q, L::absorptiontype);
  s.absorptionfreq[i] = rd<double>(q, L::absorptionfreq);
  s.stokes_q[i] = rd<double>(q, L::stokes_q);
  s.stokes_u[i] = rd<double>(q, L::stokes_u);
  s.trueemissiontype[i] = rd<int>(q, L::trueemissiontype);
  s.trueem_pos_x[i] = rd<double>(q, L::trueem_pos);
  s.trueem_pos_y[i] = rd<double>(q, L::trueem_pos + 8);
  s.trueem_pos_z[i] = rd<double>(q, L::trueem_pos + 16);
  s.trueem_time[i] = rd<float>(q, L::trueem_time);
  s.type[i] = rd<int>(q, L::type);
  s.cellindex[i] = rd<int>(q, L::cellindex);
  s.escape_type[i] = rd<int>(q, L::escape_type);
  s.escape_time[i] = rd<float>(q, L::escape_time);
  s.tdecay[i] = rd<double>(q, L::tdecay);
  s.number[i] = rd<int>(q, L::number);
  s.originated_from_particlenotgamma[i] = static_cast<int>(rd<unsigned char>(q, L::originated_from_particlenotgamma));
  s.pellet_decaytype[i] = rd<int>(q, L::pellet_decaytype);
  s.pellet_nucindex[i] = rd<int>(q, L::pellet_nucindex);
  if (b == 16) {
    s.rng0[i] = rd<unsigned int>(rec, 0);
    s.rng1[i] = rd<unsigned int>(rec, 4);
    s.rng2[i] = rd<unsigned int>(rec, 8);
    s.rng3[i] = rd<unsigned int>(rec, 12);
  } else {
    s.rng0[i] = 0U;
    s.rng1[i] = 0U;
    s.rng2[i] = 0U;
    s.rng3[i] = 0U;
  }
}

AHD void soa_to_aos_one(const Tables& T, unsigned char* aos, const int stride, const long long i) {
  unsigned char* rec = aos + (i * stride);
  const int b = stride - AosLayout::size;
  unsigned char* q = rec + b;
  const PacketSoA& s = T.pkt;
  using L = AosLayout;
  wr<double>(q, L::prop_time, s.prop_time[i]);
  wr<double>(q, L::pos, s.pos_x[i]);
  wr<double>(q, L::pos + 8, s.pos_y[i]);
  wr<double>(q, L::pos + 16, s.pos_z[i]);
  wr<double>(q, L::dir, s.dir_x[i]);
  wr<double>(q, L::dir + 8, s.dir_y[i]);
  wr<double>(q, L::dir + 16, s.dir_z[i]);
  wr<double>(q, L::nu_cmf, s.nu_cmf[i]);
  wr<double>(q, L::e_cmf, s.e_cmf[i]);
  wr<double>(q, L::nu_rf, s.nu_rf[i]);
  wr<double>(q, L::e_rf, s.e_rf[i]);
  wr<int>(q, L::next_trans, s.next_trans[i]);
  wr<int>(q, L::nscatterings, s.nscatterings[i]);
  wr<int>(q, L::emissiontype, s.emissiontype[i]);
  wr<double>(q, L::em_pos, s.em_pos_x[i]);
  wr<double>(q, L::em_pos + 8, s.em_pos_y[i]);
  wr<double>(q, L::em_pos + 16, s.em_pos_z[i]);
  wr<float>(q, L::em_time, s.em_time[i]);
  wr<int>(q, L::absorptiontype, s.absorptiontype[i]);
  wr<double>(q, L::absorptionfreq, s.absorptionfreq[i]);
  wr<double>(q, L::stokes_q, s.stokes_q[i]);
  wr<double>(q, L::stokes_u, s.stokes_u[i]);
  wr<int>(q, L::trueemissiontype, s.trueemissiontype[i]);
  wr<double>(q, L::trueem_pos, s.trueem_pos_x[i]);
  wr<double>(q, L::trueem_pos + 8, s.trueem_pos_y[i]);
  wr<double>(q, L::trueem_pos + 16, s.trueem_pos_z[i]);
  wr<float>(q, L::trueem_time, s.trueem_time[i]);
  wr<int>(q, L::type, s.type[i]);
  wr<int>(q, L::cellindex, s.cellindex[i]);
  wr<int>(q, L::escape_type, s.escape_type[i]);
  wr<float>(q, L::escape_time, s.escape_time[i]);
  wr<double>(q, L::tdecay, s.tdecay[i]);
  wr<int>(q, L::number, s.number[i]);
  wr<unsigned char>(q, L::originated_from_particlenotgamma,
                    static_cast<unsigned char>(s.originated_from_particlenotgamma[i] != 0 ? 1 : 0));
  wr<int>(q, L::pellet_decaytype, s.pellet_decaytype[i]);
  wr<int>(q, L::pellet_nucindex, s.pellet_nucindex[i]);
  if (b == 16 && T.rng_mode == RNG_XOSHIRO) {
    wr<unsigned int>(rec, 0, s.rng0[i]);
    wr<unsigned int>(rec, 4, s.rng1[i]);
    wr<unsigned int>(rec, 8, s.rng2[i]);
    wr<unsigned int>(rec, 12, s.rng3[i]);
  }
}

// Philox streams restart every timestep: counter word 0 = draw index (reset to 0), key word 1 = packet number
AHD void reset_philox_one(const Tables& T, const long long i) {
  T.pkt.rng0[i] = 0U;
  T.pkt.rng1[i] = static_cast<unsigned int>(T.pkt.number[i]);
  T.pkt.rng2[i] = 0U;
  T.pkt.rng3[i] = 0U;
}

// ---- per-cell table build steps, one work item each (see rates.h) ------------------------------------
}  // namespace ab

#include "rates.h"
#include "rpkt.h"

namespace ab {

enum : int { TESTK_BOUNDARY_DISTANCE = 0, TESTK_CLOSEST_TRANSITION = 1, TESTK_CHI_RPKT_CONT = 2 };

// one element of artisb200_test_kernel()
AHD void test_kernel_item(const Tables& T, const int which, const long long i, const long long tid, const double* in_f64,
                          const int* in_i32, double* out_f64, int* out_i32) {
  int cnt[CNT_COUNT];
  long long diag[NDIAG];
  double tss[NTSSCALARS];
  long long pellet_decays = 0;
  for (int k = 0; k < CNT_COUNT; k++) {
    cnt[k] = 0;
  }
  for (int k = 0; k < NDIAG; k++) {
    diag[k] = 0;
  }
  const Ctx c{T, 0, tid, cnt, diag, tss, &pellet_decays};
  if (which == TESTK_BOUNDARY_DISTANCE) {
    const double* rec = in_f64 + (i * 7);
    const BoundaryHit hit = boundary_distance(T, rec + 3, rec, rec[6], in_i32[i]);
    out_f64[i] = hit.distance;
    out_i32[i] = hit.next_cellindex;
  } else if (which == TESTK_CLOSEST_TRANSITION) {
    out_i32[i] = closest_transition(T, in_f64[i], in_i32[i], c);
  } else {
    ChiCont chi;
    chi.nu = -1.;
    chi.nonemptymgi = -1;
    chi.chi_escatter = 0.;
    chi.chi_freefree_heat = 0.;
    chi.chi_boundfree = 0.;
    calculate_chi_rpkt_cont(c, in_f64[i], chi, in_i32[i]);
    out_f64[(i * 3) + 0] = chi.chi_escatter;
    out_f64[(i * 3) + 1] = chi.chi_freefree_heat;
    out_f64[(i * 3) + 2] = chi.chi_boundfree;
  }
}

AHD void build_levelpop_item(const Tables& T, const int cell, const int ulev) {
  const int uion = T.level_uniqueion[ulev];
  T.cell_levelpops[(static_cast<long long>(cell) * T.nlevels) + ulev] =
      calculate_levelpop(T, cell, T.ion_element[uion], T.ion_index[uion], ulev - T.ion_levelstart[uion]);
}

AHD void build_corrphotoion_item(const Tables& T, const int cell, const int ulev) {
  const int n = T.level_nphixstargets[ulev];
  for (int k = 0; k < n; k++) {
    T.cell_corrphotoioncoeff[(static_cast<long long>(cell) * T.nphixstargets_total) + T.level_phixstargetstart[ulev] + k] =
        calc_corrphotoioncoeff(T, cell, ulev, k);
  }
}

AHD void build_keepword_item(const Tables& T, const int cell, const int word) {
  unsigned long long bits = 0ULL;
  const int begin = word * 64;
  const int end = (begin + 64 < T.nbfcontinua) ? begin + 64 : T.nbfcontinua;
  for (int i = begin; i < end; i++) {
    if (build_cell_continuum(T, cell, i)) {
      bits |= 1ULL << static_cast<unsigned>(i - begin);
    }
  }
  T.cell_cont_keepbits[(static_cast<long long>(cell) * T.keepwords) + word] = bits;
}

}  // namespace ab
