/* Compile-time options of an artis_b200 library, and the hash that guards against loading a library built for other
 * options than the host program (the reference selects its physics modes with `constexpr` values in artisoptions.h;
 * this library is compiled once per such set).
 *
 *   - In the library: namespace opt comes from a preset header (artis_b200/csrc/options/preset_*.h) or, with
 *     -DARTISB200_REFERENCE_OPTIONS, from the user's own artisoptions.h through the mapping below.
 *   - In the reference-side binding (integration/update_packets_b200.cc): compiled against the host's artisoptions.h,
 *     it defines ARTISB200_REFERENCE_OPTIONS, includes this header, and compares artisb200_options_hash_here() with the
 *     loaded library's artisb200_options_hash(); a mismatch aborts like assert_always (mpi_logging.h:123-130).
 * C++ header (the reference is C++23; the C ABI itself is include/artis_b200.h). */
#ifndef ARTIS_B200_OPTIONS_H
#define ARTIS_B200_OPTIONS_H
#include <cstdint>
#include <cstring>

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
constexpr int DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP = ::DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP;
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
// any ion with excited NLTE levels (artisoptions: ION_NLEVELS_EXCITED_NLTE(Z, ionstage)), scanned over every ion stage
constexpr bool any_ion_has_excited_nlte_levels() {
  for (int z = 1; z <= 118; z++) {
    for (int ionstage = 1; ionstage <= z + 1; ionstage++) {
      if (::ION_NLEVELS_EXCITED_NLTE(z, ionstage) > 0) {
        return true;
      }
    }
  }
  return false;
}
constexpr bool HAS_NLTE_LEVELS = any_ion_has_excited_nlte_levels();
}  // namespace opt
#endif  /* ARTISB200_REFERENCE_OPTIONS */

/* every compile-time value that changes what the library computes, in a fixed order (new entries go to the END) */
#define ARTISB200_OPTION_VALUE_LIST(X)                                                                                   \
  X(POL_ON) X(DIPOLE) X(USE_RELATIVISTIC_DOPPLER_SHIFT) X(PHIXS_CLASSIC_NO_INTERPOLATION) X(USE_LUT_PHOTOION)            \
  X(USE_ION_BFHEATING_ESTIMATORS) X(DETAILED_BF_ESTIMATORS_ON) X(MULTIBIN_RADFIELD_MODEL_ON) X(RADFIELDBINCOUNT)         \
  X(FIRST_NLTE_RADFIELD_TIMESTEP) X(RADFIELDBINS_NU_MIN) X(RADFIELDBINS_NU_MAX) X(RADFIELDBINS_T_E_SUPERBIN_NU_MAX)      \
  X(DIRECT_COL_HEAT) X(NT_ON) X(NT_SOLVE_SPENCERFANO) X(NT_EXCITATION_ON) X(NT_MAX_AUGER_ELECTRONS)                      \
  X(NTEXCITATION_MAXNLEVELS_LOWER) X(NTEXCITATION_MAXNLEVELS_UPPER) X(LTEPOP_EXCITATION_USE_TJ)                          \
  X(BFCOOLING_USELEVELPOPNOTIONPOP) X(RPKT_USE_EXPANSION_OPACITIES) X(HAS_BB_THERMALISATION_PROBABILITY)                 \
  X(BB_THERMALISATION_PROBABILITY) X(USE_XCOM_GAMMAPHOTOION) X(HAS_GAMMA_KAPPA_GREY) X(GAMMA_KAPPA_GREY)                 \
  X(FORCE_SPHERICAL_ESCAPE_SURFACE) X(PARTICLE_THERMALISATION_SCHEME) X(GAMMA_THERMALISATION_SCHEME) X(MINPOP)           \
  X(NU_MIN_R) X(NU_MAX_R) X(HAS_NLTE_LEVELS) X(DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP)

/* FNV-1a over the option values as IEEE doubles, in list order (the preset NAME is not part of it) */
inline std::uint64_t artisb200_options_hash_here() {
  const double values[] = {
#define X(name) static_cast<double>(opt::name),
      ARTISB200_OPTION_VALUE_LIST(X)
#undef X
  };
  std::uint64_t h = 1469598103934665603ULL;
  for (const double v : values) {
    unsigned char bytes[sizeof(double)];
    std::memcpy(bytes, &v, sizeof(double));
    for (const unsigned char b : bytes) {
      h ^= b;
      h *= 1099511628211ULL;
    }
  }
  return h;
}

#endif /* ARTIS_B200_OPTIONS_H */
