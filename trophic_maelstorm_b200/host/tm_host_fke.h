// tm_host_fke.h -- reference-facing host layer of the FKE movement step: the reference's own entry points, same
// names and argument lists, implemented over the C ABI (include/tm_fke.h).  F/ = Rietkerk_FPE/rietkerk_bjork_FKE.cpp.
//
//   set_global_FKE_params                 F/:195   (header rietkerk_bjork_basic.h:303)
//   precompute_gaussian                   F/:1735
//   advdiff_FKE_MultiSp                   F/:3614  (header :448)       host vectors in, host vectors out
//   copyToDeviceConstantMemory_AdvTerms   kernel_advdiff_FKE2D.cu:28   (header :466)   the -DBARRACUDA seam:
//   launchAdvDiffKernel_MultiSp           kernel_advdiff_FKE2D.cu:222  (header :457)   caller-owned DEVICE pointers
//
// The reference fixes the species count at compile time (-DSPB=2|3: Sp = SpB + 2 fields, species 1..SpB-1 move);
// the templates below deduce it from the array arguments, so a reference driver's call compiles unchanged.
// Its window bounds live in process-global inline variables / __constant__ memory; here they live in fke_globals()
// (host) and in one device context per calling thread.
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace tm_b200 {

typedef std::vector<std::vector<double>> D2Vec_Double;
typedef std::vector<std::vector<std::pair<double, double>>> D2Vec_Pair_Double;   // F/ header :188

// what the reference keeps in inline globals (F/ header :117-135), indexed by species s in [0, SpB); s = 0 unused
struct FKEGlobals {
    int SpB = 0, L = 0;
    double Dt_ay = 0, dt2 = 0, dt6 = 0, dx2 = 0, dx1 = 0, dx1_2 = 0, L2 = 0;
    double epsilon = 1.0e-10;                                   // header :119
    std::vector<double> beta_vel, exp_beta_Dt, exp_2beta_Dt, sigma_D, mu_vel_prefactor;
    std::vector<std::pair<double, double>> sigma_v, sigma_vD;
    std::vector<int> sig_D_ScaledBounds;
    std::vector<std::pair<int, int>> sig_vD_ScaledBounds;
};
FKEGlobals &fke_globals();

void set_global_FKE_params_n(const double *D, const double *v, const double *dtAdv, int SpB, double Dt_Analytic,
                             double dx, int L, int CI);
template <int SP, int SPB>
inline void set_global_FKE_params(double (&D)[SP], double (&v)[SPB], double (&dtAdv)[SPB], double Dt_Analytic, double dx,
                                  int L, int CI = 4)
{
    static_assert(SP == SPB + 2, "Sp = SpB + 2 (vegetation, consumers, soil water, surface water)");
    set_global_FKE_params_n(D, v, dtAdv, SPB, Dt_Analytic, dx, L, CI);
}

std::vector<double> precompute_gaussian(double sigma, int Rbound);

// One movement step of species 1..SpB-1 on the GPU; Rho_t[s] is replaced by the moved field (and Rho_tsar[s] holds it
// too), exactly what the reference leaves behind (F/:3690-3691).  Windows: the CPU path's convention (F/:3640-3641).
void advdiff_FKE_MultiSp(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &gamma, D2Vec_Pair_Double &v_eff,
                         D2Vec_Double &gaussian_stencil_D, D2Vec_Double &gaussian_stencil_VXY, int g);

// ---- the -DBARRACUDA seam ------------------------------------------------------------------------------------
struct int2_ { int x, y; };   // layout of CUDA's int2
// Returns 0 (cudaSuccess) or a non-zero code; n_species = the reference's compile-time CuSpNV = SpB - 1.
int copyToDeviceConstantMemory_AdvTerms(const int *sig_D_ScaledBounds, const int2_ *sigma_vD_ScBounds,
                                        const double *mu_vel_prefactor, const int *size_gauss_D,
                                        const int *size_gauss_VXY, int n_species);
// Device pointers: Rho_t, gamma [n_species][L*L]; v_eff [n_species][L*L] pairs; the two stencil arrays are the
// concatenations the reference uploads (offsets size_gauss_D / size_gauss_VXY above); Rho_tsar receives the result
// (it need not be zeroed).  Windows: the CUDA kernel's convention (rows +-.x, columns +-.y, .cu:126-127).
// Returns after the result is complete, like a kernel launch followed by the reference's blocking cudaMemcpy.
void launchAdvDiffKernel_MultiSp(double *Rho_t, double *Rho_tsar, double *gamma, std::pair<double, double> *v_eff,
                                 double *gaussian_stencil_D, double *gaussian_stencil_VXY, int L);

} // namespace tm_b200
