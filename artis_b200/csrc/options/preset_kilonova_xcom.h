// artisoptions_kilonova_lte.h with USE_XCOM_GAMMAPHOTOION = true (the reference's tests/setup_kilonova_2d_xcomgammaphotoion.sh):
#pragma once
#define ARTISB200_PRESET_NAME "kilonova_xcom"
namespace opt {
constexpr bool POL_ON = false;
constexpr bool DIPOLE = false;
constexpr bool USE_RELATIVISTIC_DOPPLER_SHIFT = true;
constexpr bool PHIXS_CLASSIC_NO_INTERPOLATION = false;
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
constexpr bool DIRECT_COL_HEAT = true;
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
constexpr bool USE_XCOM_GAMMAPHOTOION = true;
constexpr bool HAS_GAMMA_KAPPA_GREY = false;
constexpr double GAMMA_KAPPA_GREY = 0.;
constexpr bool FORCE_SPHERICAL_ESCAPE_SURFACE = false;
constexpr int PARTICLE_THERMALISATION_SCHEME = 1;  // TIMEDEPENDENT
constexpr int GAMMA_THERMALISATION_SCHEME = 0;     // FREQUENCYDEPENDENT
constexpr double MINPOP = 1e-40;
constexpr double NU_MIN_R = 1e13;
constexpr double NU_MAX_R = 5e16;
constexpr bool HAS_NLTE_LEVELS = false;
}  // namespace opt
