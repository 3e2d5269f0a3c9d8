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

// reference Packet (240 B, or 256 B with the GPU_ON rngstate prefix) -> device records
AHD void aos_to_soa_one(const Tables& T, const unsigned char* aos, const int stride, const long long i) {
  const unsigned char* rec = aos + (i * stride);
  const int b = stride - AosLayout::size;  // 0 or 16
  const unsigned char* q = rec + b;
  const PacketStore& s = T.pkt;
  using L = AosLayout;
  HotA ha;
  ha.prop_time = rd<double>(q, L::prop_time);
  HotB hb;
  for (int d = 0; d < 3; d++) {
    ha.pos[d] = rd<double>(q, L::pos + (8 * d));
    ha.dir[d] = rd<double>(q, L::dir + (8 * d));
  }
  ha.nu_cmf = rd<double>(q, L::nu_cmf);
  hb.e_cmf = rd<double>(q, L::e_cmf);
  hb.nu_rf = rd<double>(q, L::nu_rf);
  hb.e_rf = rd<double>(q, L::e_rf);
  hb.stokes_q = rd<double>(q, L::stokes_q);
  hb.stokes_u = rd<double>(q, L::stokes_u);
  hb.chi_nu = -1.;
  hb.chi_escatter = 0.;
  hb.chi_ff = 0.;
  HotC hc;
  hc.chi_bf = 0.;
  hc.next_trans = rd<int>(q, L::next_trans);
  hc.type = rd<int>(q, L::type);
  hc.cellindex = rd<int>(q, L::cellindex);
  hc.stage = pack_stage(ST_DONE, EV_NONE);
  for (int k = 0; k < 4; k++) {
    hc.rng[k] = (b == 16) ? rd<unsigned int>(rec, 4 * k) : 0U;
    hc.ma[k] = -1;
  }
  hc.chi_mgi = -1;
  hc.nscatterings = rd<int>(q, L::nscatterings);
  EmRec em;
  EmRec trueem;
  for (int d = 0; d < 3; d++) {
    em.pos[d] = rd<double>(q, L::em_pos + (8 * d));
    trueem.pos[d] = rd<double>(q, L::trueem_pos + (8 * d));
  }
  em.time = rd<float>(q, L::em_time);
  em.type = rd<int>(q, L::emissiontype);
  trueem.time = rd<float>(q, L::trueem_time);
  trueem.type = rd<int>(q, L::trueemissiontype);
  s.ha[i] = ha;
  s.hb[i] = hb;
  s.hc[i] = hc;
  s.em[i] = em;
  s.trueem[i] = trueem;
  s.absorptiontype[i] = rd<int>(q, L::absorptiontype);
  s.absorptionfreq[i] = rd<double>(q, L::absorptionfreq);
  s.escape_type[i] = rd<int>(q, L::escape_type);
  s.escape_time[i] = rd<float>(q, L::escape_time);
  s.tdecay[i] = rd<double>(q, L::tdecay);
  s.number[i] = rd<int>(q, L::number);
  s.originated_from_particlenotgamma[i] = static_cast<int>(rd<unsigned char>(q, L::originated_from_particlenotgamma));
  s.pellet_decaytype[i] = rd<int>(q, L::pellet_decaytype);
  s.pellet_nucindex[i] = rd<int>(q, L::pellet_nucindex);
  if (b == 16) {
    // the host's generator state travels with the packet (with Philox the device neither uses nor changes it)
    RngPrefix pre;
    for (int k = 0; k < 4; k++) {
      pre.w[k] = rd<unsigned int>(rec, 4 * k);
    }
    s.rngprefix[i] = pre;
  }
}

// packet i of the device records as one reference Packet at `rec`
AHD void soa_to_aos_rec(const Tables& T, unsigned char* rec, const int stride, const long long i);

AHD void soa_to_aos_one(const Tables& T, unsigned char* aos, const int stride, const long long i) {
  soa_to_aos_rec(T, aos + (i * stride), stride, i);
}

AHD void soa_to_aos_rec(const Tables& T, unsigned char* rec, const int stride, const long long i) {
  const int b = stride - AosLayout::size;
  unsigned char* q = rec + b;
  const PacketStore& s = T.pkt;
  using L = AosLayout;
  const HotA ha = s.ha[i];
  const HotB hb = s.hb[i];
  const HotC hc = s.hc[i];
  const EmRec em = s.em[i];
  const EmRec trueem = s.trueem[i];
  wr<double>(q, L::prop_time, ha.prop_time);
  for (int d = 0; d < 3; d++) {
    wr<double>(q, L::pos + (8 * d), ha.pos[d]);
    wr<double>(q, L::dir + (8 * d), ha.dir[d]);
    wr<double>(q, L::em_pos + (8 * d), em.pos[d]);
    wr<double>(q, L::trueem_pos + (8 * d), trueem.pos[d]);
  }
  wr<double>(q, L::nu_cmf, ha.nu_cmf);
  wr<double>(q, L::e_cmf, hb.e_cmf);
  wr<double>(q, L::nu_rf, hb.nu_rf);
  wr<double>(q, L::e_rf, hb.e_rf);
  wr<int>(q, L::next_trans, hc.next_trans);
  wr<int>(q, L::nscatterings, hc.nscatterings);
  wr<int>(q, L::emissiontype, em.type);
  wr<float>(q, L::em_time, em.time);
  wr<int>(q, L::absorptiontype, s.absorptiontype[i]);
  wr<double>(q, L::absorptionfreq, s.absorptionfreq[i]);
  wr<double>(q, L::stokes_q, hb.stokes_q);
  wr<double>(q, L::stokes_u, hb.stokes_u);
  wr<int>(q, L::trueemissiontype, trueem.type);
  wr<float>(q, L::trueem_time, trueem.time);
  wr<int>(q, L::type, hc.type);
  wr<int>(q, L::cellindex, hc.cellindex);
  wr<int>(q, L::escape_type, s.escape_type[i]);
  wr<float>(q, L::escape_time, s.escape_time[i]);
  wr<double>(q, L::tdecay, s.tdecay[i]);
  wr<int>(q, L::number, s.number[i]);
  wr<unsigned char>(q, L::originated_from_particlenotgamma,
                    static_cast<unsigned char>(s.originated_from_particlenotgamma[i] != 0 ? 1 : 0));
  wr<int>(q, L::pellet_decaytype, s.pellet_decaytype[i]);
  wr<int>(q, L::pellet_nucindex, s.pellet_nucindex[i]);
  if (b == 16) {
    const RngPrefix pre = s.rngprefix[i];
    for (int k = 0; k < 4; k++) {
      wr<unsigned int>(rec, 4 * k, (T.rng_mode == RNG_XOSHIRO) ? hc.rng[k] : pre.w[k]);
    }
  }
}

// ---- per-cell table build steps, one work item each (see rates.h) ------------------------------------
}  // namespace ab

#include "rates.h"
#include "rpkt.h"

namespace ab {

enum : int { TESTK_BOUNDARY_DISTANCE = 0, TESTK_CLOSEST_TRANSITION = 1, TESTK_CHI_RPKT_CONT = 2, TESTK_SELECT_CONTINUUM_NU = 3 };

// one element of artisb200_test_kernel()
AHD void test_kernel_item(const Tables& T, Accum& acc, const int which, const long long i, const long long tid,
                          const double* in_f64, const int* in_i32, double* out_f64, int* out_i32) {
  unsigned int hot[Ctx::NHOT] = {};
  const Ctx c{T, tid, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};  // `tid` selects the scratch column
  if (which == TESTK_BOUNDARY_DISTANCE) {
    const double* rec = in_f64 + (i * 7);
    const BoundaryHit hit = boundary_distance(T, rec + 3, rec, rec[6], in_i32[i]);
    out_f64[i] = hit.distance;
    out_i32[i] = hit.next_cellindex;
  } else if (which == TESTK_CLOSEST_TRANSITION) {
    out_i32[i] = closest_transition(T, in_f64[i], in_i32[i], c);
  } else if (which == TESTK_SELECT_CONTINUUM_NU) {
    // in: (T_e, zrand) per item and the continuum (index into the allcont list); out: the sampled frequency
    const int ci = in_i32[i];
    out_f64[i] = select_continuum_nu_z(T, T.cont_element[ci], T.cont_ion[ci], T.cont_level[ci], T.cont_phixstargetindex[ci],
                                       static_cast<float>(in_f64[(i * 2) + 0]), in_f64[(i * 2) + 1]);
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

// the time-independent factor of the Sobolev optical depth of one line in one cell (rpkt.cc:75-100), for the line walk
AHD void build_linetau_item(const Tables& T, const int cell, const int lineindex) {
  const double* cellpops = T.cell_levelpops + (static_cast<long long>(cell) * T.nlevels);
  const double n_l = cellpops[T.line_lower[lineindex]];
  const double n_u = cellpops[T.line_upper[lineindex]];
  const double B_ul = T.line_B_ul[lineindex];
  const double B_lu = T.line_B_lu[lineindex];
  T.cell_linetau[(static_cast<long long>(cell) * T.nlines) + lineindex] = ((B_lu * n_l) - (B_ul * n_u)) * HCLIGHTOVERFOURPI;
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

// list and per-word ranks of the kept continua of one cell (see Tables::cell_cont_keptlist)
AHD void build_keptlist_cell(const Tables& T, const int cell) {
  const unsigned long long* keepbits = T.cell_cont_keepbits + (static_cast<long long>(cell) * T.keepwords);
  int* list = T.cell_cont_keptlist + (static_cast<long long>(cell) * T.nbfcontinua);
  int* rank = T.cell_cont_keptrank + (static_cast<long long>(cell) * (T.keepwords + 1));
  int n = 0;
  for (int word = 0; word < T.keepwords; word++) {
    rank[word] = n;
    unsigned long long bits = keepbits[word];
    while (bits != 0ULL) {
      list[n++] = (word * 64) + lowest_set_bit(bits);
      bits &= bits - 1;
    }
  }
  rank[T.keepwords] = n;
}

}  // namespace ab
