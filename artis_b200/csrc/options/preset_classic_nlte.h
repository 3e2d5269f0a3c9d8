// artisoptions_classic.h with NLTE level populations for some ions (ION_NLEVELS_EXCITED_NLTE > 0): the hot path
// then reads the NLTE solver's populations (ltepop.cc:168-199), used by the classic_nlte_toy parity case.
#pragma once
#define ARTISB200_PRESET_NAME "classic_nlte"
namespace opt {
constexpr bool POL_ON = true;
constexpr bool DIPOLE = true;
constexpr bool USE_RELATIVISTIC_DOPPLER_SHIFT = false;
constexpr bool PHIXS_CLASSIC_NO_INTERPOLATION = true;
constexpr bool USE_LUT_PHOTOION = true;
constexpr bool USE_ION_BFHEATING_ESTIMATORS = true;
constexpr int DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP = 13;
constexpr bool DETAILED_BF_ESTIMATORS_ON = false;
constexpr bool MULTIBIN_RADFIELD_MODEL_ON = false;
constexpr int RADFIELDBINCOUNT = 256;
constexpr int FIRST_NLTE_RADFIELD_TIMESTEP = 12;
constexpr double RADFIELDBINS_NU_MIN = 2.99792458e+10 / 40000e-8;
constexpr double RADFIELDBINS_NU_MAX = 2.99792458e+10 / 1085e-8;
constexpr double RADFIELDBINS_T_E_SUPERBIN_NU_MAX = 2.99792458e+10 / 10e-8;
constexpr bool DIRECT_COL_HEAT = false;
constexpr bool NT_ON = false;
constexpr bool NT_SOLVE_SPENCERFANO = false;
constexpr bool NT_EXCITATION_ON = false;
constexpr int NT_MAX_AUGER_ELECTRONS = 2;
constexpr int NTEXCITATION_MAXNLEVELS_LOWER = 5;
constexpr int NTEXCITATION_MAXNLEVELS_UPPER = 250;
constexpr bool LTEPOP_EXCITATION_USE_TJ = true;
constexpr bool BFCOOLING_USELEVELPOPNOTIONPOP = false;
constexpr bool RPKT_USE_EXPANSION_OPACITIES = false;
constexpr bool HAS_BB_THERMALISATION_PROBABILITY = false;
constexpr float BB_THERMALISATION_PROBABILITY = 0.F;
constexpr bool USE_XCOM_GAMMAPHOTOION = false;
constexpr bool HAS_GAMMA_KAPPA_GREY = false;
constexpr double GAMMA_KAPPA_GREY = 0.;
constexpr bool FORCE_SPHERICAL_ESCAPE_SURFACE = false;
constexpr int PARTICLE_THERMALISATION_SCHEME = 0;  // INSTANTFULLDEPOSITION
constexpr int GAMMA_THERMALISATION_SCHEME = 0;     // FREQUENCYDEPENDENT
constexpr double MINPOP = 1e-30;
constexpr double NU_MIN_R = 1e14;
constexpr double NU_MAX_R = 5e15;
constexpr bool HAS_NLTE_LEVELS = true;
}  // namespace opt
