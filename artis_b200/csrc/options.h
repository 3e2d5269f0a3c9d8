// Compile-time physics modes. The reference selects these with artisoptions.h (`constexpr` values pruned by
// `if constexpr`); this library is likewise compiled once per preset:
//   -DARTISB200_PRESET_HEADER='"options/preset_classic.h"'   restated hot-path subset shipped in this repo, or
//   -DARTISB200_REFERENCE_OPTIONS -I<artis source dir>        the user's own artisoptions.h, read directly.
#pragma once

#ifdef ARTISB200_REFERENCE_OPTIONS
#include "../../include/artis_b200_options.h"  // namespace opt from the reference's own artisoptions.h
#else
#ifndef ARTISB200_PRESET_HEADER
#error "compile with -DARTISB200_PRESET_HEADER='\"options/preset_<name>.h\"' or -DARTISB200_REFERENCE_OPTIONS"
#endif
#include ARTISB200_PRESET_HEADER
#include "../../include/artis_b200_options.h"  // canonical value list + hash of namespace opt
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
static_assert(GAMMA_THERMALISATION_SCHEME >= GTS_FREQUENCYDEPENDENT && GAMMA_THERMALISATION_SCHEME <= GTS_GUTTMAN,
              "unknown gamma-ray thermalisation scheme");
}  // namespace opt
