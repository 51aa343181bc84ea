// tm_host_fke.cpp -- see tm_host_fke.h.  Host bookkeeping only; all lattice numerics go through include/tm_fke.h.
#include "tm_host_fke.h"
#include "tm_host.h"
#include "../../include/tm_fke.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

namespace tm_b200 {

FKEGlobals &fke_globals()
{
    static FKEGlobals g;
    return g;
}

// the reference's PI: 15 Bailey-Borwein-Plouffe terms summed upwards (F/ header :241-259, :284)
static double ref_pi()
{
    double pi = 0.0, p16 = 1.0;
    for (int n = 0; n <= 14; n++) {
        pi = (1.0 / p16 * (4.0 / (8 * n + 1.0) - 2.0 / (8 * n + 4.0) - 1.0 / (8 * n + 5.0) - 1.0 / (8 * n + 6.0))) + pi;
        p16 *= 16.0;
    }
    return pi;
}

static double min_positive(const double *a, int n)
{
    double m = -1;
    for (int i = 0; i < n; i++)
        if (a[i] > 0 && (m < 0 || a[i] < m)) m = a[i];
    return m;
}

// F/:195-262
void set_global_FKE_params_n(const double *D, const double *v, const double *dtAdv, int SpB, double Dt_Analytic, double dx,
                             int L, int CI)
{
    FKEGlobals &G = fke_globals();
    if (Dt_Analytic != min_positive(dtAdv, SpB)) {
        if (!options().quiet)
            std::cout << "WARNING: Dt_Analytic is not equal to the smallest non-zero dtAdv. Setting Dt_Analytic to the smallest non-zero dtAdv.\n";
        Dt_Analytic = min_positive(dtAdv, SpB);
    }
    G.SpB = SpB; G.L = L;
    G.Dt_ay = Dt_Analytic;
    G.dt2 = G.Dt_ay / 2.0; G.dt6 = G.Dt_ay / 6.0; G.L2 = (double)L * L;
    G.dx2 = dx * dx; G.dx1 = 1.0 / dx; G.dx1_2 = 1.0 / (dx * dx);
    G.beta_vel.assign(SpB, 0.0); G.exp_beta_Dt.assign(SpB, 0.0); G.exp_2beta_Dt.assign(SpB, 0.0);
    G.sigma_D.assign(SpB, 0.0); G.mu_vel_prefactor.assign(SpB, 0.0);
    G.sigma_v.assign(SpB, {0.0, 0.0}); G.sigma_vD.assign(SpB, {0.0, 0.0});
    G.sig_D_ScaledBounds.assign(SpB, 0); G.sig_vD_ScaledBounds.assign(SpB, {0, 0});
    for (int s = 1; s < SpB; s++) {
        G.beta_vel[s] = 3.0 / Dt_Analytic;                                  // displacement invariance (no MV_INVARIANCE)
        G.exp_beta_Dt[s] = std::exp(-G.beta_vel[s] * Dt_Analytic);
        G.exp_2beta_Dt[s] = std::exp(-2 * G.beta_vel[s] * Dt_Analytic);
        const double sv = std::sqrt(2 * G.beta_vel[s] * 1 * std::abs(v[s]));
        G.sigma_v[s] = {sv, sv};
        G.sigma_D[s] = std::sqrt(2 * D[s] * Dt_Analytic);
        const double sigma_vx = (sv / G.beta_vel[s]) * std::sqrt((G.Dt_ay + (1 / (2 * G.beta_vel[s])) * (1 - G.exp_2beta_Dt[s])
                                                                  - (2 / G.beta_vel[s]) * (1 - G.exp_beta_Dt[s])));
        G.sigma_vD[s] = {sigma_vx, sigma_vx};
        G.sig_D_ScaledBounds[s] = int(CI * G.sigma_D[s] / dx);
        G.sig_vD_ScaledBounds[s] = {int(CI * sigma_vx / dx), int(CI * sigma_vx / dx)};
        G.mu_vel_prefactor[s] = (1 / dx) * (G.Dt_ay - ((1 - G.exp_beta_Dt[s]) / G.beta_vel[s]));
    }
}

// F/:1735-1749
std::vector<double> precompute_gaussian(double sigma, int Rbound)
{
    std::vector<double> stencil(Rbound + 1);
    double sum = 0.0;
    const double factor = 1.0 / (std::sqrt(2 * ref_pi()) * sigma);
    for (int x = 0; x <= Rbound; ++x) {
        stencil[x] = factor * std::exp(-0.5 * (x * x) / (sigma * sigma));
        sum += (x > 0) ? 2 * stencil[x] : stencil[0];
    }
    for (int x = 0; x <= Rbound; ++x) stencil[x] /= sum;
    return stencil;
}

namespace {
// one device context per calling thread (the reference runs one OpenMP thread per a-value, R/:4509)
struct ThreadCtx {
    tm_fke_ctx *ctx = nullptr;
    int L = 0, n = 0;
    ~ThreadCtx() { if (ctx) tm_fke_destroy(ctx); }
    tm_fke_ctx *get(int L_, int n_)
    {
        if (ctx && (L != L_ || n != n_)) { tm_fke_destroy(ctx); ctx = nullptr; }
        if (!ctx) {
            if (tm_fke_create(L_, n_, current_device(), &ctx) != 0)
                throw std::runtime_error(std::string("tm_fke_create: ") + tm_fke_last_error(nullptr));
            L = L_; n = n_;
        }
        return ctx;
    }
};
thread_local ThreadCtx t_ctx;

void ck(tm_fke_ctx *c, int rc, const char *what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tm_fke_last_error(c));
}
}

// F/:3614-3695
void advdiff_FKE_MultiSp(D2Vec_Double &Rho_t, D2Vec_Double &Rho_tsar, D2Vec_Double &gamma, D2Vec_Pair_Double &v_eff,
                         D2Vec_Double &gD, D2Vec_Double &gV, int g)
{
    const FKEGlobals &G = fke_globals();
    const int SpB = G.SpB, nsp = SpB - 1;
    if (nsp < 1 || (int)Rho_t.size() < SpB || (int)gamma.size() < SpB || (int)v_eff.size() < SpB)
        throw std::invalid_argument("advdiff_FKE_MultiSp: call set_global_FKE_params first; arrays are indexed by species 0..SpB-1");
    tm_fke_ctx *c = t_ctx.get(g, nsp);
    const size_t n = (size_t)g * g;
    std::vector<double> rho(nsp * n), gam(nsp * n), vel(2 * nsp * n), out(nsp * n);
    for (int s = 1; s < SpB; s++) {
        tm_fke_terms t{};
        t.bound_D = G.sig_D_ScaledBounds[s];
        t.v_row_lo = t.v_col_lo = G.sig_vD_ScaledBounds[s].first;    // F/:3640-3641: (first, second) = (lower, upper) on both axes
        t.v_row_hi = t.v_col_hi = G.sig_vD_ScaledBounds[s].second;
        t.n_gauss_D = (int)gD[s].size(); t.n_gauss_V = (int)gV[s].size();
        t.mu_vel_prefactor = G.mu_vel_prefactor[s]; t.epsilon = G.epsilon;
        ck(c, tm_fke_set_terms(c, s - 1, &t, gD[s].data(), gV[s].data()), "tm_fke_set_terms");
        for (size_t i = 0; i < n; i++) {
            rho[(s - 1) * n + i] = Rho_t[s][i]; gam[(s - 1) * n + i] = gamma[s][i];
            vel[2 * ((s - 1) * n + i)] = v_eff[s][i].first; vel[2 * ((s - 1) * n + i) + 1] = v_eff[s][i].second;
        }
    }
    ck(c, tm_fke_move_host(c, rho.data(), gam.data(), vel.data(), out.data()), "tm_fke_move_host");
    for (int s = 1; s < SpB; s++)
        for (size_t i = 0; i < n; i++) Rho_t[s][i] = Rho_tsar[s][i] = out[(s - 1) * n + i];
}

// ---- the -DBARRACUDA seam ----------------------------------------------------------------------------------------
namespace {
struct AdvTerms {   // what the reference parks in __constant__ memory (F/ header :96-100)
    int n = 0;
    std::vector<int> bD, offD, offV;
    std::vector<int2_> bV;
    std::vector<double> pref;
};
AdvTerms g_adv;     // process-global like the reference's constant bank; read-only after the copy call
}

int copyToDeviceConstantMemory_AdvTerms(const int *sig_D_ScaledBounds, const int2_ *sigma_vD_ScBounds, const double *mu_vel_prefactor,
                                        const int *size_gauss_D, const int *size_gauss_VXY, int n_species)
{
    if (!sig_D_ScaledBounds || !sigma_vD_ScBounds || !mu_vel_prefactor || !size_gauss_D || !size_gauss_VXY || n_species < 1) return 1;
    g_adv.n = n_species;
    g_adv.bD.assign(sig_D_ScaledBounds, sig_D_ScaledBounds + n_species);
    g_adv.bV.assign(sigma_vD_ScBounds, sigma_vD_ScBounds + n_species);
    g_adv.pref.assign(mu_vel_prefactor, mu_vel_prefactor + n_species);
    g_adv.offD.assign(size_gauss_D, size_gauss_D + n_species);
    g_adv.offV.assign(size_gauss_VXY, size_gauss_VXY + n_species);
    return 0;
}

void launchAdvDiffKernel_MultiSp(double *Rho_t, double *Rho_tsar, double *gamma, std::pair<double, double> *v_eff,
                                 double *gaussian_stencil_D, double *gaussian_stencil_VXY, int L)
{
    static_assert(sizeof(std::pair<double, double>) == 2 * sizeof(double), "pair<double,double> is two packed doubles");
    if (g_adv.n < 1) throw std::logic_error("launchAdvDiffKernel_MultiSp: copyToDeviceConstantMemory_AdvTerms has not been called");
    tm_fke_ctx *c = t_ctx.get(L, g_adv.n);
    for (int s = 0; s < g_adv.n; s++) {
        tm_fke_terms t{};
        t.bound_D = g_adv.bD[s];
        t.v_row_lo = t.v_row_hi = g_adv.bV[s].x;                       // .cu:126: rows +- .x
        t.v_col_lo = t.v_col_hi = g_adv.bV[s].y;                       // .cu:127: columns +- .y
        t.n_gauss_D = t.bound_D + 1;
        t.n_gauss_V = (g_adv.bV[s].x > g_adv.bV[s].y ? g_adv.bV[s].x : g_adv.bV[s].y) + 1;
        t.mu_vel_prefactor = g_adv.pref[s]; t.epsilon = fke_globals().epsilon;
        ck(c, tm_fke_set_terms_device(c, s, &t, gaussian_stencil_D + g_adv.offD[s], gaussian_stencil_VXY + g_adv.offV[s]),
           "tm_fke_set_terms_device");
    }
    ck(c, tm_fke_move_device(c, Rho_t, gamma, reinterpret_cast<const double *>(v_eff), Rho_tsar), "tm_fke_move_device");
    ck(c, tm_fke_sync(c), "tm_fke_sync");
}

} // namespace tm_b200
