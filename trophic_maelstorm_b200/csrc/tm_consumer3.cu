// tm_consumer3.cu -- C ABI (include/tm_consumer3.h) over the SPB = 3 launch chain (tm_consumer3.cuh).  sm_100a only; no CPU path.
#include "../../include/tm_consumer3.h"
#include "../../include/tm_lattice.h" // TM_E_* codes, struct tm_params
#include "../../include/tm_fke.h"     // tm_fke_terms
#include "tm_consumer3.cuh"

#include <cuda_runtime.h>
#include <cmath>
#include <new>
#include <string>

using namespace tml;

static thread_local std::string g_cons3_create_error;

struct tm_cons3_ctx {
    int L = 0, device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_state = nullptr, *d_star = nullptr, *d_ya = nullptr, *d_yb = nullptr, *d_acc = nullptr;   // [5][n], [3][n], [5][n] x 3
    double *d_gamma = nullptr, *d_veff = nullptr, *d_part = nullptr;                                   // [2][n], [2][n][2], partials
    Step3Dyn *d_dyn = nullptr;
    int *d_central = nullptr;
    unsigned long long *d_ovf = nullptr;
    double *d_replay = nullptr; size_t replay_cap = 0;
    double *d_fke_gD[2] = {nullptr, nullptr}, *d_fke_gV[2] = {nullptr, nullptr};   // FPE flavour: the consumers' movement stencils
    bool ready = false;
    int strict_means = 0;          // tm_cons3_set_strict_means
    Step3Consts k{};
    int n_launches = 0;
    std::string err;
};

#define C3_CK(call)                                                                    \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return TM_E_CUDA;                                                          \
        }                                                                              \
    } while (0)

namespace {
struct CudaLaunch3 {
    cudaStream_t st;
    template <class K, class A> void operator()(K kernel, const A &args, unsigned grid, unsigned block, size_t smem) const
    {
        kernel<<<grid, block, smem, st>>>(args);
    }
};

Step3Buffers buffers(tm_cons3_ctx *c)
{
    Step3Buffers b{};
    b.state = c->d_state; b.star = c->d_star; b.ya = c->d_ya; b.yb = c->d_yb; b.acc = c->d_acc; b.gamma = c->d_gamma; b.veff = c->d_veff;
    b.part = c->d_part; b.dyn = c->d_dyn; b.central = c->d_central; b.overflow = c->d_ovf;
    for (int s = 0; s < 2; s++) { b.fke_gD[s] = c->d_fke_gD[s]; b.fke_gV[s] = c->d_fke_gV[s]; }
    b.strict_means = c->strict_means;
    return b;
}
} // namespace

extern "C" {

const char *tm_cons3_last_error(const tm_cons3_ctx *c) { return c ? c->err.c_str() : g_cons3_create_error.c_str(); }

int tm_cons3_destroy(tm_cons3_ctx *c)
{
    if (!c) return TM_E_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_state); cudaFree(c->d_star); cudaFree(c->d_ya); cudaFree(c->d_yb); cudaFree(c->d_acc); cudaFree(c->d_gamma);
    for (int s = 0; s < 2; s++) { cudaFree(c->d_fke_gD[s]); cudaFree(c->d_fke_gV[s]); }
    cudaFree(c->d_veff); cudaFree(c->d_part); cudaFree(c->d_dyn); cudaFree(c->d_central); cudaFree(c->d_ovf); cudaFree(c->d_replay);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return TM_OK;
}

int tm_cons3_create(int L, int device, tm_cons3_ctx **out)
{
    if (!out) return TM_E_ARG;
    *out = nullptr;
    auto fail = [&](tm_cons3_ctx *c, int code, const std::string &msg) { g_cons3_create_error = msg; if (c) tm_cons3_destroy(c); return code; };
    if (L < 2 || L > 32767) return fail(nullptr, TM_E_ARG, "tm_cons3_create: need 2 <= L <= 32767");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(nullptr, TM_E_CUDA, "tm_cons3_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, TM_E_ARG, "tm_cons3_create: bad device index");
    tm_cons3_ctx *c = new (std::nothrow) tm_cons3_ctx;
    if (!c) return fail(nullptr, TM_E_ARG, "tm_cons3_create: out of host memory");
    c->L = L; c->device = device;
    const size_t n = (size_t)L * L * sizeof(double);
#define CR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(c, TM_E_CUDA, std::string("tm_cons3_create: " #call ": ") + cudaGetErrorString(e_)); } while (0)
    CR(cudaSetDevice(device));
    CR(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CR(cudaEventCreate(&c->ev0)); CR(cudaEventCreate(&c->ev1));
    CR(cudaMalloc(&c->d_state, 5 * n)); CR(cudaMalloc(&c->d_star, 3 * n));
    CR(cudaMalloc(&c->d_ya, 5 * n)); CR(cudaMalloc(&c->d_yb, 5 * n)); CR(cudaMalloc(&c->d_acc, 5 * n));
    CR(cudaMalloc(&c->d_gamma, 2 * n)); CR(cudaMalloc(&c->d_veff, 4 * n));
    CR(cudaMalloc(&c->d_part, sizeof(double) * 4 * kConsReduceBlocks));
    CR(cudaMalloc(&c->d_dyn, sizeof(Step3Dyn)));
    CR(cudaMalloc(&c->d_ovf, sizeof(unsigned long long)));
    CR(cudaMemset(c->d_gamma, 0, 2 * n)); CR(cudaMemset(c->d_veff, 0, 4 * n)); CR(cudaMemset(c->d_dyn, 0, sizeof(Step3Dyn)));
    CR(cudaMemset(c->d_ovf, 0, sizeof(unsigned long long)));
#undef CR
    *out = c;
    return TM_OK;
}

int tm_cons3_setup(tm_cons3_ctx *c, const tm_params *veg, const tm_graz_params *graz, const tm_graz_params *pred, const tm_cons3_params *pc,
                   const int32_t *central, int n_central, double fr_graz, double fr_pred)
{
    if (!c) return TM_E_ARG;
    if (!veg || !graz || !pred || !pc || !central || veg->L != c->L || graz->L != c->L || pred->L != c->L || pc->L != c->L || n_central < 1 ||
        n_central > 10000 || !(fr_graz >= 0 && fr_graz <= 1) || !(fr_pred >= 0 && fr_pred <= 1) || !(graz->sigma > 0) || !(pred->sigma > 0) ||
        !(veg->sigma0 > 0) || !(veg->dx > 0)) {
        c->err = "tm_cons3_setup: bad argument (all four L must equal the context's; 1 <= n_central <= 10000; 0 <= fr <= 1; sigma, dx > 0)";
        return TM_E_ARG;
    }
    C3_CK(cudaSetDevice(c->device));
    cudaFree(c->d_central); c->d_central = nullptr;
    C3_CK(cudaMalloc(&c->d_central, sizeof(int) * 2 * n_central));
    C3_CK(cudaMemcpy(c->d_central, central, sizeof(int) * 2 * n_central, cudaMemcpyHostToDevice));
    C3_CK(cudaFuncSetAttribute(cons3_gamma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 10000 * (int)sizeof(int)));
    const Cons3Consts cc = cons3_consts(pc->dt, pc->a, pc->c, pc->gmax, pc->alpha, pc->rW, pc->W0, pc->K1, pc->K2, pc->Dw, pc->Do,
                                        pc->A01, pc->H01, pc->E1, pc->A12, pc->H12, pc->E2);
    const double gD[2] = {graz->D, pred->D}, gv[2] = {graz->v, pred->v}, gM[2] = {graz->M, pred->M}, gs[2] = {graz->sigma, pred->sigma};
    c->k = step3_consts(c->L, n_central, fr_graz, fr_pred, veg->dt, veg->dx, veg->d, veg->D0, veg->sigma0, veg->clamp_eps, veg->clamp_lo,
                        gD, gv, gM, gs, graz->D0_floor, cc);
    c->ready = true;
    return TM_OK;
}

int tm_cons3_set_fpe(tm_cons3_ctx *c, const tm_fke_terms *t, const double *const *gauss_D, const double *const *gauss_V, double M1, double M2)
{
    if (!c) return TM_E_ARG;
    if (!c->ready) { c->err = "tm_cons3_set_fpe: call tm_cons3_setup first"; return TM_E_ARG; }
    if (!t) {                                      // back to the Rietkerk_FD flavour
        c->k.fpe = 0; c->k.c.fpe = 0; c->k.c.M1 = 0.0; c->k.c.M2 = 0.0;
        return TM_OK;
    }
    if (!gauss_D || !gauss_V) { c->err = "tm_cons3_set_fpe: stencils are required"; return TM_E_ARG; }
    for (int s = 0; s < 2; s++) {
        const int vmax = fke_max3(fke_max3(t[s].v_row_lo, t[s].v_row_hi, t[s].v_col_lo), t[s].v_col_hi, 0);
        if (!gauss_D[s] || !gauss_V[s] || t[s].bound_D < 0 || t[s].v_row_lo < 0 || t[s].v_row_hi < 0 || t[s].v_col_lo < 0 || t[s].v_col_hi < 0 ||
            t[s].bound_D > kFkeMaxBound || vmax > kFkeMaxBound || t[s].n_gauss_D < t[s].bound_D + 1 || t[s].n_gauss_V < vmax + 1) {
            c->err = "tm_cons3_set_fpe: window bounds must lie in [0, 512] and the stencils must cover them";
            return TM_E_ARG;
        }
        // every cell of consumer s is displaced by |v_s| mu_vel_prefactor_s cells (F/:3635-3637); the gather-form column pass visits
        // (2 R + 1)^2 candidate sources per target for a largest displacement R (tm_fke.cuh), fine for the reference's fractions of a
        // cell, quadratic beyond: refuse a configuration that would turn the pass into minutes per step instead of running it
        if (!(fabs(c->k.v[s] * t[s].mu_vel_prefactor) <= 64.0)) {
            c->err = "tm_cons3_set_fpe: |v| * mu_vel_prefactor exceeds 64 cells per step (the gather-form movement pass is sized for small displacements)";
            return TM_E_ARG;
        }
    }
    C3_CK(cudaSetDevice(c->device));
    C3_CK(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < 2; s++) {
        cudaFree(c->d_fke_gD[s]); cudaFree(c->d_fke_gV[s]); c->d_fke_gD[s] = c->d_fke_gV[s] = nullptr;
        C3_CK(cudaMalloc(&c->d_fke_gD[s], t[s].n_gauss_D * sizeof(double)));
        C3_CK(cudaMalloc(&c->d_fke_gV[s], t[s].n_gauss_V * sizeof(double)));
        C3_CK(cudaMemcpy(c->d_fke_gD[s], gauss_D[s], t[s].n_gauss_D * sizeof(double), cudaMemcpyHostToDevice));
        C3_CK(cudaMemcpy(c->d_fke_gV[s], gauss_V[s], t[s].n_gauss_V * sizeof(double), cudaMemcpyHostToDevice));
        FkeTerms &f = c->k.fke[s];
        f.L = c->L; f.bD = t[s].bound_D; f.rlo = t[s].v_row_lo; f.rhi = t[s].v_row_hi; f.clo = t[s].v_col_lo; f.chi = t[s].v_col_hi;
        f.pref = t[s].mu_vel_prefactor; f.eps = t[s].epsilon;
    }
    c->k.fpe = 1; c->k.c.fpe = 1; c->k.c.M1 = M1; c->k.c.M2 = M2;
    return TM_OK;
}

int tm_cons3_set_strict_means(tm_cons3_ctx *c, int on)
{
    if (!c) return TM_E_ARG;
    c->strict_means = on ? 1 : 0;
    return TM_OK;
}

int tm_cons3_upload(tm_cons3_ctx *c, const double *fields, const double *gamma)
{
    if (!c) return TM_E_ARG;
    if (!c->ready || !fields) { c->err = "tm_cons3_upload: call tm_cons3_setup first; fields must not be NULL"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * sizeof(double);
    C3_CK(cudaSetDevice(c->device));
    C3_CK(cudaMemcpyAsync(c->d_state, fields, 5 * n, cudaMemcpyHostToDevice, c->stream));
    if (gamma) C3_CK(cudaMemcpyAsync(c->d_gamma, gamma, 2 * n, cudaMemcpyHostToDevice, c->stream));
    else C3_CK(cudaMemsetAsync(c->d_gamma, 0, 2 * n, c->stream));
    C3_CK(cudaMemsetAsync(c->d_ovf, 0, sizeof(unsigned long long), c->stream));
    C3_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_cons3_run(tm_cons3_ctx *c, uint64_t seed, uint32_t trial, uint32_t step0, int nsteps)
{
    if (!c) return TM_E_ARG;
    if (!c->ready || nsteps < 0) { c->err = "tm_cons3_run: call tm_cons3_setup first; nsteps >= 0"; return TM_E_ARG; }
    C3_CK(cudaSetDevice(c->device));
    Step3Buffers b = buffers(c);
    C3_CK(cudaEventRecord(c->ev0, c->stream));
    for (int s = 0; s < nsteps; s++) cons_step3_enqueue(CudaLaunch3{c->stream}, c->k, b, seed, trial, step0 + (uint32_t)s);
    C3_CK(cudaGetLastError());
    C3_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = step3_launches(c->k.fpe) * nsteps;
    return TM_OK;
}

int tm_cons3_replay(tm_cons3_ctx *c, const double *u_veg, const double *u_graz, const double *u_pred, uint32_t stride)
{
    if (!c) return TM_E_ARG;
    if (!c->ready || !u_veg || !u_graz || !u_pred || stride < 1) { c->err = "tm_cons3_replay: call tm_cons3_setup first; three uniform arrays and stride >= 1 are required"; return TM_E_ARG; }
    C3_CK(cudaSetDevice(c->device));
    const size_t nn = (size_t)c->L * c->L * stride, need = nn * sizeof(double);
    if (3 * need > c->replay_cap) {
        cudaFree(c->d_replay); c->d_replay = nullptr; c->replay_cap = 0;
        C3_CK(cudaMalloc(&c->d_replay, 3 * need));
        c->replay_cap = 3 * need;
    }
    const double *src[3] = {u_veg, u_graz, u_pred};
    for (int s = 0; s < 3; s++) C3_CK(cudaMemcpyAsync(c->d_replay + s * nn, src[s], need, cudaMemcpyHostToDevice, c->stream));
    Step3Buffers b = buffers(c);
    for (int s = 0; s < 3; s++) b.replay[s] = c->d_replay + s * nn;
    b.stride = stride;
    C3_CK(cudaEventRecord(c->ev0, c->stream));
    cons_step3_enqueue(CudaLaunch3{c->stream}, c->k, b, 0, 0, 0);
    C3_CK(cudaGetLastError());
    C3_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = step3_launches(c->k.fpe);
    C3_CK(cudaStreamSynchronize(c->stream));   // the uniform arrays are the caller's again
    return TM_OK;
}

int tm_cons3_download(tm_cons3_ctx *c, double *fields, double *gamma, double *v_eff, double *means, uint64_t *overflow)
{
    if (!c) return TM_E_ARG;
    if (!c->ready) { c->err = "tm_cons3_download: call tm_cons3_setup first"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * sizeof(double);
    C3_CK(cudaSetDevice(c->device));
    unsigned long long ovf = 0;
    Step3Dyn d{};
    if (fields) C3_CK(cudaMemcpyAsync(fields, c->d_state, 5 * n, cudaMemcpyDeviceToHost, c->stream));
    if (gamma) C3_CK(cudaMemcpyAsync(gamma, c->d_gamma, 2 * n, cudaMemcpyDeviceToHost, c->stream));
    if (v_eff) C3_CK(cudaMemcpyAsync(v_eff, c->d_veff, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    C3_CK(cudaMemcpyAsync(&d, c->d_dyn, sizeof d, cudaMemcpyDeviceToHost, c->stream));
    C3_CK(cudaMemcpyAsync(&ovf, c->d_ovf, sizeof ovf, cudaMemcpyDeviceToHost, c->stream));
    C3_CK(cudaStreamSynchronize(c->stream));
    if (means) {
        const double cells = (double)c->L * (double)c->L;
        for (int s = 0; s < 3; s++) means[s] = d.sum[s] / cells;
        means[3] = d.occ_veg; means[4] = (double)d.eff[0]; means[5] = (double)d.eff[1];
    }
    if (overflow) *overflow = ovf;
    return TM_OK;
}

int tm_cons3_last_ms(tm_cons3_ctx *c, float *ms, int *n_launches)
{
    if (!c || !ms) return TM_E_ARG;
    C3_CK(cudaSetDevice(c->device));
    C3_CK(cudaEventSynchronize(c->ev1));
    C3_CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (n_launches) *n_launches = c->n_launches;
    return TM_OK;
}

} // extern "C"
