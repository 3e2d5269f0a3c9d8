// Macro-atom random walk (Lucy 2002/2003 transition-probability scheme) on the precomputed per-cell tables.
// Reference: macroatom.cc:360-596 (do_macroatom), 204-244 (raddeexcitation), 248-294 (radrecomb),
// 298-322 (ionisation), 44-62 (cumulative-array views); sn3d.h:85-92 (index_upperbound).
#pragma once
#include "atomicdata.h"
#include "emit.h"
#include "hd.h"
#include "options.h"
#include "packet.h"
#include "rates.h"

namespace ab {

// First index in [0, n) whose value is > target (std::upper_bound over a non-decreasing array; sn3d.h:85-92).
// The tables searched here are the per-cell cumulative transition rates: gigabytes in HBM, visited at random, so
// every probe of a binary search is a dependent DRAM access. This search issues 7 independent probes per round
// (8-way split) and finishes with one round over the last <= 8 elements: 2 rounds instead of 7 dependent loads
// for a 70-entry row, and the same index as the binary search for any sorted input. The work counter keeps counting
// the probes of the binary search (the compulsory traffic of the roofline model).
// positions of the (up to 7) first-round pivots of an array of length n: pos_k = k * ceil(n / 8) - 1, k = 1..7, while < n
AHD int upperbound_pivot_pos(const int n, const int k) { return (k * ((n + 7) >> 3)) - 1; }

// `pivots` (optional): the values a[pos_k] of the first round, read together with the process rates (tables.h
// cell_marecord) - the first round then needs no access to the array itself
AHD int index_upperbound(const double* a, const int n, const double target, const Ctx& c, const double* pivots = nullptr) {
  int lo = 0;
  int len = n;
  // probes of the binary search over n entries. (This loop, not a count-leading-zeros: with the one-instruction form the
  // macro-atom stage was measured 9 % SLOWER, 106.1 against 97.7 ms, twice - the loop's few cycles sit between the address
  // arithmetic and the first probe loads, and ptxas schedules the kernel differently without it.)
  int probes = 0;
  for (int m = n; m > 0; m >>= 1) {
    probes++;
  }
  c.work<DIAG_BINSEARCH_STEPS>(probes);
  if (pivots != nullptr && len > 8) {
    const int step = (len + 7) >> 3;
    int npassed = 0;
    int nvalid = 0;
#pragma unroll
    for (int k = 1; k <= 7; k++) {
      const int pos = (k * step) - 1;
      if (pos < len) {
        nvalid++;
        npassed += (!(target < pivots[k - 1])) ? 1 : 0;
      }
    }
    lo += npassed * step;
    len = (npassed < nvalid) ? (step - 1) : (len - (npassed * step));
  }
  while (len > 8) {
    const int step = (len + 7) >> 3;
    int npassed = 0;  // pivots that are <= target (the pivots are sorted, so these are the first npassed)
    int nvalid = 0;
#pragma unroll
    for (int k = 1; k <= 7; k++) {
      const int pos = (k * step) - 1;
      if (pos < len) {
        nvalid++;
        npassed += (!(target < a[lo + pos])) ? 1 : 0;
      }
    }
    lo += npassed * step;
    // the answer is after the last passed pivot and not after the first failed one
    len = (npassed < nvalid) ? (step - 1) : (len - (npassed * step));
  }
  int count = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (k < len) {
      count += (!(target < a[lo + k])) ? 1 : 0;
    }
  }
  return lo + count;
}


// Take up to `max_steps` transitions (<= 0: walk to deactivation) of the activation recorded in p.ma.
// The walk ends with p.ma_pending == 0 and either a k-packet, or an r-packet whose re-emission is left pending
// (EV_EMIT_MA, run by finish_ma_emission); otherwise p.ma holds the level reached and the walk continues in the
// packet's next visit to the macro-atom stage.
AHD void do_macroatom(Pkt& p, const Ctx& c, const int max_steps) {
  const Tables& T = c.T;
  const MacroAtomState mastate = p.ma;
  const int cell = T.propcell_nonemptymgi[p.cellindex];
  const auto T_e = T.Te[cell];
  const auto clumpednne_ = T.clumpfactor[cell] * T.nne[cell];

  const int element = mastate.element;
  int ion = mastate.ion;
  int level = mastate.level;
  const int activatingline = mastate.activatingline;

  const double* cellrates = T.cell_maprocessrates + (static_cast<long long>(cell) * T.nlevels * MA_ACTION_COUNT);
  const double* cellmatrans = T.cell_matrans + (static_cast<long long>(cell) * T.matrans_total);
  const double* cellrecords =
      (T.cell_marecord != nullptr) ? T.cell_marecord + (static_cast<long long>(cell) * T.nlevels * MA_RECORD) : nullptr;

  bool end_packet = false;
  int nsteps = 0;
  while (!end_packet) {
    if (max_steps > 0 && nsteps >= max_steps) {
      p.ma = {element, ion, level, activatingline};
      return;
    }
    nsteps++;
    const int ustart = levelstart(T, element, ion);
    const int ulev = ustart + level;
    const double epsilon_current = epsilon(T, ulev);
    // the level's walk record: 9 process rates + the first-round pivots of its three cumulative arrays (tables.h)
    const double* levelrates = (cellrecords != nullptr) ? cellrecords + (static_cast<long long>(ulev) * MA_RECORD)
                                                        : cellrates + (static_cast<long long>(ulev) * MA_ACTION_COUNT);
    const double* pivots = (cellrecords != nullptr) ? levelrates + MA_ACTION_COUNT : nullptr;

    // partial sums of the 9 process rates (macroatom.cc:424-425) and selection by upper_bound (433-437)
    double cumulative[MA_ACTION_COUNT];
    double running = 0.;
#pragma unroll
    for (int a = 0; a < MA_ACTION_COUNT; a++) {
      running = (a == 0) ? levelrates[0] : running + levelrates[a];
      cumulative[a] = running;
    }
    const double total_rate = cumulative[MA_ACTION_COUNT - 1];
    const double randomrate = p.rng.uniform() * total_rate;
    int selected_action = 0;
#pragma unroll
    for (int a = 0; a < MA_ACTION_COUNT; a++) {
      selected_action += (!(randomrate < cumulative[a])) ? 1 : 0;  // == upper_bound on a sorted array
    }
    selected_action = (selected_action < MA_ACTION_COUNT - 1) ? selected_action : MA_ACTION_COUNT - 1;

    c.count<CNT_INTERACTIONS>();
    c.work<DIAG_MA_STEPS>();

    const int ndowntrans = T.level_ndowntrans[ulev];
    const double* transblock = cellmatrans + T.level_matransblock_start[ulev];

    switch (selected_action) {
      case MA_ACTION_RADDEEXC: {
        // macroatom.cc:204-244
        const double targetval = p.rng.uniform() * levelrates[MA_ACTION_RADDEEXC];
        const int downtransindex = index_upperbound(transblock, ndowntrans - 1, targetval, c, pivots);
        const int alltrans_startdown = T.level_alltrans_startdown[ulev];
        const int lineindex = T.trans_lineindex[alltrans_startdown + downtransindex];
        if (lineindex == activatingline) {
          c.count<CNT_RESONANCESCATTERINGS>();
        }
        const int ulevlower = ustart + T.trans_targetlevelindex[alltrans_startdown + downtransindex];
        const double epsilon_trans = epsilon_current - epsilon(T, ulevlower);
        const double oldnucmf = (activatingline >= 0) ? (p.kin_in_memory ? T.pkt.ha[c.ip].nu_cmf : p.nu_cmf) : 0.;
        p.nu_cmf = epsilon_trans / H;
        if (activatingline >= 0) {
          if (oldnucmf < p.nu_cmf) {
            c.count<CNT_UPSCATTER>();
          } else {
            c.count<CNT_DOWNSCATTER>();
          }
        }
        c.count<CNT_MA_STAT_DEACTIVATION_BB>();
        p.type = TYPE_RPKT;
        p.ev_pending = EV_EMIT_MA;
        p.next_trans = lineindex + 1;
        T.pkt.em[c.ip].type = lineindex;
        p.nscatterings = 0;
        end_packet = true;
        break;
      }

      case MA_ACTION_COLDEEXC: {
        c.count<CNT_MA_STAT_DEACTIVATION_COLLDEEXC>();
        p.type = TYPE_KPKT;
        end_packet = true;
        if constexpr (!opt::DIRECT_COL_HEAT) {
          atomic_add(&T.est_colheating[cell], p.kin_in_memory ? T.pkt.hb[c.ip].e_cmf : p.e_cmf);
          c.work<DIAG_ESTIMATOR_ADDS>();
        }
        break;
      }

      case MA_ACTION_INTERNALDOWNSAME: {
        const double targetval = p.rng.uniform() * levelrates[MA_ACTION_INTERNALDOWNSAME];
        const int downtransindex =
            index_upperbound(transblock + ndowntrans, ndowntrans - 1, targetval, c, (pivots != nullptr) ? pivots + 7 : nullptr);
        level = T.trans_targetlevelindex[T.level_alltrans_startdown[ulev] + downtransindex];
        break;
      }

      case MA_ACTION_RADRECOMB: {
        // macroatom.cc:248-294
        const double targetval = p.rng.uniform() * levelrates[MA_ACTION_RADRECOMB];
        double rate = 0;
        const int nlevels = nlevels_ionising(T, element, ion - 1);
        const int lowerionstart = levelstart(T, element, ion - 1);
        int lowerionlevel = -1;
        int selected_phixstargetindex = -1;
        for (int tmp = 0; tmp < nlevels; tmp++) {
          const int phixstargetindex = find_phixstargetindex(T, lowerionstart + tmp, level);
          if (phixstargetindex < 0) {
            continue;
          }
          const double epsilon_trans = epsilon_current - epsilon(T, lowerionstart + tmp);
          const double R = rad_recombination_ratecoeff(T, T_e, clumpednne_, element, ion, tmp, phixstargetindex);
          rate += R * epsilon_trans;
          if (targetval < rate) {
            lowerionlevel = tmp;
            selected_phixstargetindex = phixstargetindex;
            break;
          }
        }
        if (lowerionlevel < 0) {
          // cannot happen for consistent tables (reference: assert_always): reported through the device error record;
          // deactivate to the last possible continuum rather than looping forever
          c.fail(DEVERR_MA_RADRECOMB_NO_LEVEL, ulev);
          for (int tmp = nlevels - 1; tmp >= 0 && lowerionlevel < 0; tmp--) {
            const int phixstargetindex = find_phixstargetindex(T, lowerionstart + tmp, level);
            if (phixstargetindex >= 0) {
              lowerionlevel = tmp;
              selected_phixstargetindex = phixstargetindex;
            }
          }
        }
        const int lowerion = ion - 1;
        p.nu_cmf = select_continuum_nu(T, element, lowerion, lowerionlevel, selected_phixstargetindex, T_e, p.rng);
        c.count<CNT_MA_STAT_DEACTIVATION_FB>();
        p.type = TYPE_RPKT;
        p.ev_pending = EV_EMIT_MA;
        p.next_trans = -1;
        T.pkt.em[c.ip].type = emtype_continuum(T, lowerionstart + lowerionlevel, selected_phixstargetindex);
        p.nscatterings = 0;
        level = lowerionlevel;
        ion -= 1;
        end_packet = true;
        break;
      }

      case MA_ACTION_COLRECOMB: {
        c.count<CNT_MA_STAT_DEACTIVATION_COLLRECOMB>();
        p.type = TYPE_KPKT;
        end_packet = true;
        if constexpr (!opt::DIRECT_COL_HEAT) {
          atomic_add(&T.est_colheating[cell], p.kin_in_memory ? T.pkt.hb[c.ip].e_cmf : p.e_cmf);
          c.work<DIAG_ESTIMATOR_ADDS>();
        }
        break;
      }

      case MA_ACTION_INTERNALDOWNLOWER: {
        c.count<CNT_MA_STAT_INTERNALDOWNLOWER>();
        const double targetrate = p.rng.uniform() * levelrates[MA_ACTION_INTERNALDOWNLOWER];
        double rate = 0.;
        const int nlevels = nlevels_ionising(T, element, ion - 1);
        int lower = -1;
        const int lowerionstart = levelstart(T, element, ion - 1);
        for (int tmp = 0; tmp < nlevels; tmp++) {
          const int phixstargetindex = find_phixstargetindex(T, lowerionstart + tmp, level);
          if (phixstargetindex < 0) {
            continue;
          }
          const double epsilon_target = epsilon(T, lowerionstart + tmp);
          const double epsilon_trans = epsilon_current - epsilon_target;
          const double R = rad_recombination_ratecoeff(T, T_e, clumpednne_, element, ion, tmp, phixstargetindex);
          const double C = col_recombination_ratecoeff(T, T_e, clumpednne_, element, ion, tmp, phixstargetindex, epsilon_trans);
          rate += (R + C) * epsilon_target;
          if (rate > targetrate) {
            lower = tmp;
            break;
          }
        }
        if (lower < 0) {
          c.fail(DEVERR_MA_DOWNLOWER_NO_LEVEL, ulev);  // reference: assert_always(lower >= 0)
          lower = 0;
        }
        ion--;
        level = lower;
        break;
      }

      case MA_ACTION_INTERNALUPSAME: {
        const int nuptrans = T.level_nuptrans[ulev];
        const double targetval = p.rng.uniform() * levelrates[MA_ACTION_INTERNALUPSAME];
        const int uptransindex =
            index_upperbound(transblock + (2 * ndowntrans), nuptrans - 1, targetval, c, (pivots != nullptr) ? pivots + 14 : nullptr);
        level = T.trans_targetlevelindex[alltrans_startup(T, ulev) + uptransindex];
        break;
      }

      case MA_ACTION_INTERNALUPHIGHER: {
        // macroatom.cc:298-322
        c.count<CNT_MA_STAT_INTERNALUPHIGHER>();
        const double targetrate = p.rng.uniform() * levelrates[MA_ACTION_INTERNALUPHIGHER];
        double rate = 0.;
        const int nphixstargets = T.level_nphixstargets[ulev];
        int newlevel = -1;
        for (int phixstargetindex = 0; phixstargetindex < nphixstargets; phixstargetindex++) {
          const double epsilon_trans = phixs_threshold(T, element, ion, level, phixstargetindex);
          const double R = cell_corrphotoioncoeff(T, cell, ulev, phixstargetindex);
          const double C = col_ionisation_ratecoeff(T, T_e, clumpednne_, element, ion, level, phixstargetindex, epsilon_trans);
          rate += (R + C) * epsilon_current;
          if (rate > targetrate) {
            newlevel = phixsupperlevel(T, ulev, phixstargetindex);
            break;
          }
        }
        if (newlevel < 0) {
          c.fail(DEVERR_MA_IONISATION_NO_TARGET, ulev);  // reference: assert_always(false)
          newlevel = phixsupperlevel(T, ulev, nphixstargets - 1);
        }
        level = newlevel;
        ion += 1;
        break;
      }

      default: {  // MA_ACTION_INTERNALUPHIGHERNT (macroatom.cc:562-568)
        ion = nt_random_upperion(T, cell, element, ion, false, p.rng);
        level = 0;
        c.count<CNT_MA_STAT_INTERNALUPHIGHERNT>();
        break;
      }
    }
  }

  p.ma_pending = 0;
  if (p.type != TYPE_RPKT) {
    T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;  // macroatom.cc:588-590
  }
}

// the r-packet emission a radiative deactivation left pending: emit_rpkt (macroatom.cc:237, 282) and the
// true-emission bookkeeping that follows the walk (macroatom.cc:579-587)
AHD void finish_ma_emission(Pkt& p, const Ctx& c) {
  const Tables& T = c.T;
  p.ev_pending = EV_NONE;
  emit_rpkt(p, c);
  if (T.pkt.trueem[c.ip].type == EMTYPE_NOTSET) {
    T.pkt.trueem[c.ip].type = T.pkt.em[c.ip].type;
    set_trueem_here(p, c);
  }
}

}  // namespace ab
