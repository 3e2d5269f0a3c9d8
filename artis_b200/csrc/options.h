// Compile-time physics modes. The reference selects these with artisoptions.h (`constexpr` values pruned by
// `if constexpr`); this library is likewise compiled once per preset:
//   -DARTISB200_PRESET_HEADER='"options/preset_classic.h"'   restated hot-path subset shipped in this repo, or
//   -DARTISB200_REFERENCE_OPTIONS -I<artis source dir>        the user's own artisoptions.h, read directly.
#pragma once

#ifdef ARTISB200_REFERENCE_OPTIONS
#include "artisoptions.h"  // the reference's own header (needs its constants.h for the enums)
#define ARTISB200_PRESET_NAME "reference-artisoptions"
namespace opt {
constexpr bool POL_ON = ::POL_ON;
constexpr bool DIPOLE = ::DIPOLE;
constexpr bool USE_RELATIVISTIC_DOPPLER_SHIFT = ::USE_RELATIVISTIC_DOPPLER_SHIFT;
constexpr bool PHIXS_CLASSIC_NO_INTERPOLATION = ::PHIXS_CLASSIC_NO_INTERPOLATION;
constexpr bool USE_LUT_PHOTOION = ::USE_LUT_PHOTOION;
constexpr bool USE_ION_BFHEATING_ESTIMATORS = ::USE_ION_BFHEATING_ESTIMATORS;
constexpr bool DETAILED_BF_ESTIMATORS_ON = ::DETAILED_BF_ESTIMATORS_ON;
constexpr bool MULTIBIN_RADFIELD_MODEL_ON = ::MULTIBIN_RADFIELD_MODEL_ON;
constexpr int RADFIELDBINCOUNT = ::RADFIELDBINCOUNT;
constexpr int FIRST_NLTE_RADFIELD_TIMESTEP = ::FIRST_NLTE_RADFIELD_TIMESTEP;
constexpr double RADFIELDBINS_NU_MIN = ::RADFIELDBINS_NU_MIN;
constexpr double RADFIELDBINS_NU_MAX = ::RADFIELDBINS_NU_MAX;
constexpr double RADFIELDBINS_T_E_SUPERBIN_NU_MAX = ::RADFIELDBINS_T_E_SUPERBIN_NU_MAX;
constexpr bool DIRECT_COL_HEAT = ::DIRECT_COL_HEAT;
constexpr bool NT_ON = ::NT_ON;
constexpr bool NT_SOLVE_SPENCERFANO = ::NT_SOLVE_SPENCERFANO;
constexpr bool NT_EXCITATION_ON = ::NT_EXCITATION_ON;
constexpr int NT_MAX_AUGER_ELECTRONS = ::NT_MAX_AUGER_ELECTRONS;
constexpr int NTEXCITATION_MAXNLEVELS_LOWER = ::NTEXCITATION_MAXNLEVELS_LOWER;
constexpr int NTEXCITATION_MAXNLEVELS_UPPER = ::NTEXCITATION_MAXNLEVELS_UPPER;
constexpr bool LTEPOP_EXCITATION_USE_TJ = ::LTEPOP_EXCITATION_USE_TJ;
constexpr bool BFCOOLING_USELEVELPOPNOTIONPOP = ::BFCOOLING_USELEVELPOPNOTIONPOP;
constexpr bool RPKT_USE_EXPANSION_OPACITIES = ::RPKT_USE_EXPANSION_OPACITIES;
constexpr bool HAS_BB_THERMALISATION_PROBABILITY = ::RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY.has_value();
constexpr float BB_THERMALISATION_PROBABILITY = ::RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY.value_or(0.F);
constexpr bool USE_XCOM_GAMMAPHOTOION = ::USE_XCOM_GAMMAPHOTOION;
constexpr bool HAS_GAMMA_KAPPA_GREY = ::GAMMA_USE_KAPPA_GREY.has_value();
constexpr double GAMMA_KAPPA_GREY = ::GAMMA_USE_KAPPA_GREY.value_or(0.);
constexpr bool FORCE_SPHERICAL_ESCAPE_SURFACE = ::FORCE_SPHERICAL_ESCAPE_SURFACE;
constexpr int PARTICLE_THERMALISATION_SCHEME = static_cast<int>(::PARTICLE_THERMALISATION_SCHEME);
constexpr int GAMMA_THERMALISATION_SCHEME = static_cast<int>(::GAMMA_THERMALISATION_SCHEME);
constexpr double MINPOP = ::MINPOP;
constexpr double NU_MIN_R = ::NU_MIN_R;
constexpr double NU_MAX_R = ::NU_MAX_R;
constexpr bool HAS_NLTE_LEVELS = (::ION_NLEVELS_EXCITED_NLTE(26, 2) > 0) || (::ION_NLEVELS_EXCITED_NLTE(8, 1) > 0);
}  // namespace opt
#else
#ifndef ARTISB200_PRESET_HEADER
#error "compile with -DARTISB200_PRESET_HEADER='\"options/preset_<name>.h\"' or -DARTISB200_REFERENCE_OPTIONS"
#endif
#include ARTISB200_PRESET_HEADER
#endif

namespace opt {
// ParticleThermalisationScheme / GammaThermalisationScheme enumerator values (reference constants.h:87-96)
constexpr int PTS_INSTANTFULLDEPOSITION = 0;
constexpr int PTS_TIMEDEPENDENT = 1;
constexpr int PTS_TIMEDEPENDENT_WITH_ADIABATIC_LOSS = 2;
constexpr int PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS = 3;
constexpr int PTS_BARNES = 4;
constexpr int PTS_WOLLAEGER = 5;
constexpr int GTS_FREQUENCYDEPENDENT = 0;
constexpr int GTS_BARNES = 1;
constexpr int GTS_WOLLAEGER = 2;
constexpr int GTS_GUTTMAN = 3;

// modes of the reference that this library does not implement yet fail at compile time rather than silently
static_assert(!DETAILED_BF_ESTIMATORS_ON || !USE_LUT_PHOTOION, "DETAILED_BF_ESTIMATORS_ON needs USE_LUT_PHOTOION = false");
static_assert(!NT_EXCITATION_ON || (NT_ON && NT_SOLVE_SPENCERFANO), "NT_EXCITATION_ON needs NT_ON and NT_SOLVE_SPENCERFANO");
static_assert(!NT_SOLVE_SPENCERFANO || NT_ON, "NT_SOLVE_SPENCERFANO needs NT_ON");
static_assert(!RPKT_USE_EXPANSION_OPACITIES && !HAS_BB_THERMALISATION_PROBABILITY,
              "expansion-opacity r-packet modes are not implemented yet");
static_assert(!USE_XCOM_GAMMAPHOTOION, "XCOM gamma photoionisation tables are not implemented yet");
static_assert(GAMMA_THERMALISATION_SCHEME >= GTS_FREQUENCYDEPENDENT && GAMMA_THERMALISATION_SCHEME <= GTS_GUTTMAN,
              "unknown gamma-ray thermalisation scheme");
}  // namespace opt
