// tm_consumer.cu -- C ABI (include/tm_consumer.h) over the consumer-side kernels (tm_consumer.cuh).  sm_100a only; no CPU path.
#include "../../include/tm_consumer.h"
#include "../../include/tm_lattice.h" // TM_E_* codes
#include "../../include/tm_fke.h"     // tm_fke_terms
#include "tm_consumer.cuh"

#include <cuda_runtime.h>
#include <new>
#include <string>

using namespace tml;

static thread_local std::string g_cons_create_error;

struct tm_cons_ctx {
    int L = 0, device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_state = nullptr, *d_ya = nullptr, *d_yb = nullptr, *d_acc = nullptr;   // [4][L*L] each
    double *d_gamma = nullptr;                                                       // [L*L]
    int *d_off = nullptr; int off_cap = 0;
    double *d_replay = nullptr; size_t replay_cap = 0;
    unsigned long long *d_ovf = nullptr;
    // the resident SPB = 2 stepper (tm_cons_step2_*)
    bool step2_ready = false;
    int rk4_stages = 0;            // tm_cons_step2_set_rk4_stages
    int strict_means = 0;          // tm_cons_step2_set_strict_means
    Step2Consts k2{};
    double *d_star = nullptr, *d_veff = nullptr, *d_part = nullptr;
    Step2Dyn *d_dyn = nullptr;
    int *d_central = nullptr;
    double *d_fke_gD = nullptr, *d_fke_gV = nullptr;   // FPE flavour: the grazer's movement stencils
    int n_launches = 0;
    std::string err;
};

#define CONS_CK(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e_); return TM_E_CUDA; } \
    } while (0)

namespace {
struct CudaLaunch {
    cudaStream_t st;
    template <class K, class A> void operator()(K kernel, const A &args, unsigned grid, unsigned block, size_t smem) const
    {
        kernel<<<grid, block, smem, st>>>(args);
    }
};

static Step2Buffers step2_buffers(tm_cons_ctx *c)
{
    Step2Buffers b{};
    b.state = c->d_state; b.star = c->d_star; b.ya = c->d_ya; b.yb = c->d_yb; b.acc = c->d_acc; b.gamma = c->d_gamma; b.veff = c->d_veff;
    b.part = c->d_part; b.dyn = c->d_dyn; b.central = c->d_central; b.overflow = c->d_ovf;
    b.rk4_stages = c->rk4_stages; b.strict_means = c->strict_means;
    b.fke_gD = c->d_fke_gD; b.fke_gV = c->d_fke_gV;
    return b;
}
} // namespace

extern "C" {

const char *tm_cons_last_error(const tm_cons_ctx *c) { return c ? c->err.c_str() : g_cons_create_error.c_str(); }

int tm_cons_destroy(tm_cons_ctx *c)
{
    if (!c) return TM_E_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_state); cudaFree(c->d_ya); cudaFree(c->d_yb); cudaFree(c->d_acc); cudaFree(c->d_gamma); cudaFree(c->d_off); cudaFree(c->d_replay); cudaFree(c->d_ovf);
    cudaFree(c->d_star); cudaFree(c->d_veff); cudaFree(c->d_part); cudaFree(c->d_dyn); cudaFree(c->d_central);
    cudaFree(c->d_fke_gD); cudaFree(c->d_fke_gV);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return TM_OK;
}

int tm_cons_create(int L, int device, tm_cons_ctx **out)
{
    if (!out) return TM_E_ARG;
    *out = nullptr;
    auto fail = [&](tm_cons_ctx *c, int code, const std::string &msg) { g_cons_create_error = msg; if (c) tm_cons_destroy(c); return code; };
    if (L < 2 || L > 32767) return fail(nullptr, TM_E_ARG, "tm_cons_create: need 2 <= L <= 32767");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(nullptr, TM_E_CUDA, "tm_cons_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, TM_E_ARG, "tm_cons_create: bad device index");
    tm_cons_ctx *c = new (std::nothrow) tm_cons_ctx;
    if (!c) return fail(nullptr, TM_E_ARG, "tm_cons_create: out of host memory");
    c->L = L; c->device = device;
    const size_t n4 = (size_t)4 * L * L * sizeof(double);
#define CR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(c, TM_E_CUDA, std::string("tm_cons_create: " #call ": ") + cudaGetErrorString(e_)); } while (0)
    CR(cudaSetDevice(device));
    CR(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CR(cudaEventCreate(&c->ev0)); CR(cudaEventCreate(&c->ev1));
    CR(cudaMalloc(&c->d_state, n4)); CR(cudaMalloc(&c->d_ya, n4)); CR(cudaMalloc(&c->d_yb, n4)); CR(cudaMalloc(&c->d_acc, n4));
    CR(cudaMalloc(&c->d_gamma, n4 / 4));
#undef CR
    *out = c;
    return TM_OK;
}

int tm_cons_rk4_host(tm_cons_ctx *c, const tm_cons_params *p, double *fields, int nsteps)
{
    if (!c) return TM_E_ARG;
    if (!p || !fields || nsteps < 0 || p->L != c->L) { c->err = "tm_cons_rk4_host: bad argument (p->L must equal the context's L)"; return TM_E_ARG; }
    const size_t n4 = (size_t)4 * c->L * c->L * sizeof(double);
    CONS_CK(cudaSetDevice(c->device));
    CONS_CK(cudaMemcpyAsync(c->d_state, fields, n4, cudaMemcpyHostToDevice, c->stream));
    RkStageArgs A{};
    A.L = c->L;
    A.c.dt = p->dt; A.c.dt2 = p->dt / 2.0; A.c.dt6 = p->dt / 6.0;               // R/:171
    A.c.a = p->a; A.c.cgmax = p->c * p->gmax; A.c.gmax = p->gmax; A.c.alpha = p->alpha; A.c.rW = p->rW; A.c.K1 = p->K1; A.c.K2 = p->K2;
    A.c.K2W0 = p->K2 * p->W0; A.c.Dw = p->Dw; A.c.Do = p->Do;
    A.c.A01 = p->A01; A.c.A01H01 = p->A01 * p->H01; A.c.EA1 = p->E1 * p->A01;   // R/:2452
    const int tiles = (c->L + kConsThreads - 1) / kConsThreads;
    const unsigned grid = (unsigned)((long long)tiles * c->L);
    CONS_CK(cudaEventRecord(c->ev0, c->stream));
    c->n_launches = 0;
    for (int s = 0; s < nsteps; s++) {
        A.base = c->d_state; A.acc = c->d_acc; A.out = c->d_state;
        A.stage = 1; A.Y = c->d_state; A.Ynext = c->d_ya; cons_rk4_stage_kernel<<<grid, kConsThreads, 0, c->stream>>>(A);
        A.stage = 2; A.Y = c->d_ya; A.Ynext = c->d_yb; cons_rk4_stage_kernel<<<grid, kConsThreads, 0, c->stream>>>(A);
        A.stage = 3; A.Y = c->d_yb; A.Ynext = c->d_ya; cons_rk4_stage_kernel<<<grid, kConsThreads, 0, c->stream>>>(A);
        A.stage = 4; A.Y = c->d_ya; A.Ynext = nullptr; cons_rk4_stage_kernel<<<grid, kConsThreads, 0, c->stream>>>(A);
        c->n_launches += 4;
    }
    CONS_CK(cudaGetLastError());
    CONS_CK(cudaEventRecord(c->ev1, c->stream));
    CONS_CK(cudaMemcpyAsync(fields, c->d_state, n4, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_cons_gamma_host(tm_cons_ctx *c, const double *prey, const int32_t *offsets, int n_offsets, double rho_avg_prey,
                       double rho_avg_self, double *gamma)
{
    if (!c) return TM_E_ARG;
    if (!prey || !offsets || !gamma || n_offsets < 1 || n_offsets > 10000) { c->err = "tm_cons_gamma_host: bad argument (1 <= n_offsets <= 10000)"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * sizeof(double);
    CONS_CK(cudaSetDevice(c->device));
    if (n_offsets > c->off_cap) {
        cudaFree(c->d_off); c->d_off = nullptr; c->off_cap = 0;
        CONS_CK(cudaMalloc(&c->d_off, sizeof(int) * 2 * n_offsets));
        c->off_cap = n_offsets;
    }
    CONS_CK(cudaMemcpyAsync(c->d_off, offsets, sizeof(int) * 2 * n_offsets, cudaMemcpyHostToDevice, c->stream));
    CONS_CK(cudaMemcpyAsync(c->d_ya, prey, n, cudaMemcpyHostToDevice, c->stream));     // the stage buffer doubles as the prey field
    CONS_CK(cudaMemcpyAsync(c->d_gamma, gamma, n, cudaMemcpyHostToDevice, c->stream));
    const double eps = 1.0e-12;
    GammaArgs A{};
    A.L = c->L; A.eff = n_offsets; A.prey = c->d_ya; A.offsets = c->d_off; A.gamma = c->d_gamma;
    A.rho_inverse = rho_avg_prey >= eps ? 1 / rho_avg_prey : 1 / eps;                  // R/:2548
    A.zero = (rho_avg_self < eps || rho_avg_prey < eps) ? 1 : 0;                       // R/:2575-2578
    const int tiles = (c->L + kConsThreads - 1) / kConsThreads;
    CONS_CK(cudaFuncSetAttribute(cons_gamma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 10000 * (int)sizeof(int)));
    CONS_CK(cudaEventRecord(c->ev0, c->stream));
    cons_gamma_kernel<<<(unsigned)((long long)tiles * c->L), kConsThreads, sizeof(int) * 2 * n_offsets, c->stream>>>(A);
    CONS_CK(cudaGetLastError());
    CONS_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = 1;
    CONS_CK(cudaMemcpyAsync(gamma, c->d_gamma, n, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_cons_grazer_dornic_host(tm_cons_ctx *c, const tm_graz_params *p, uint64_t seed, uint32_t trial, uint32_t step,
                               const double *replay_u, uint32_t stride, const double *prev, double *gamma, double *rho_star,
                               double *v_eff, uint64_t *overflow)
{
    if (!c) return TM_E_ARG;
    if (!p || !prev || !gamma || !rho_star || !v_eff || p->L != c->L || (replay_u && stride < 1) || !(p->sigma > 0) || !(p->dx > 0)) {
        c->err = "tm_cons_grazer_dornic_host: bad argument (p->L must equal the context's L; sigma, dx > 0)"; return TM_E_ARG;
    }
    const size_t nn = (size_t)c->L * c->L, n = nn * sizeof(double);
    CONS_CK(cudaSetDevice(c->device));
    // staging: d_ya[0] = prev, d_ya[1] = rho*, d_ya[2..3] = v_eff; d_gamma = gamma; d_yb = replay uniforms (grown on demand)
    double *d_prev = c->d_ya, *d_out = c->d_ya + nn, *d_veff = c->d_ya + 2 * nn;
    double *d_replay = nullptr;
    if (replay_u) {
        const size_t need = nn * stride * sizeof(double);
        if (need > c->replay_cap) {
            cudaFree(c->d_replay); c->d_replay = nullptr; c->replay_cap = 0;
            CONS_CK(cudaMalloc(&c->d_replay, need));
            c->replay_cap = need;
        }
        CONS_CK(cudaMemcpyAsync(c->d_replay, replay_u, need, cudaMemcpyHostToDevice, c->stream));
        d_replay = c->d_replay;
    }
    if (!c->d_ovf) CONS_CK(cudaMalloc(&c->d_ovf, sizeof(unsigned long long)));
    CONS_CK(cudaMemsetAsync(c->d_ovf, 0, sizeof(unsigned long long), c->stream));
    CONS_CK(cudaMemcpyAsync(d_prev, prev, n, cudaMemcpyHostToDevice, c->stream));
    CONS_CK(cudaMemcpyAsync(c->d_gamma, gamma, n, cudaMemcpyHostToDevice, c->stream));
    GrazerArgs A{};
    A.L = c->L; A.dt = p->dt; A.dx = p->dx; A.D = p->D; A.v = p->v; A.M = p->M; A.sigma = p->sigma; A.D0floor = p->D0_floor;
    grazer_derive(A);
    A.prev = d_prev; A.gamma = c->d_gamma; A.out = d_out; A.veff = d_veff;
    A.seed = seed; A.trial = trial; A.step = step; A.replay_u = d_replay; A.stride = stride; A.overflow = c->d_ovf;
    const int tiles = (c->L + kConsThreads - 1) / kConsThreads;
    const unsigned grid = (unsigned)((long long)tiles * c->L);
    CONS_CK(cudaEventRecord(c->ev0, c->stream));
    if (replay_u) grazer_dornic_kernel<true><<<grid, kConsThreads, 0, c->stream>>>(A);
    else          grazer_dornic_kernel<false><<<grid, kConsThreads, 0, c->stream>>>(A);
    CONS_CK(cudaGetLastError());
    CONS_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = 1;
    unsigned long long ovf = 0;
    CONS_CK(cudaMemcpyAsync(rho_star, d_out, n, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaMemcpyAsync(v_eff, d_veff, 2 * n, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaMemcpyAsync(gamma, c->d_gamma, n, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaMemcpyAsync(&ovf, c->d_ovf, sizeof ovf, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaStreamSynchronize(c->stream));
    if (overflow) *overflow = ovf;
    return TM_OK;
}

// ---- resident SPB = 2 stepper ---------------------------------------------------------------------------------------

int tm_cons_step2_setup(tm_cons_ctx *c, const tm_params *veg, const tm_graz_params *pg, const tm_cons_params *pc,
                        const int32_t *central, int n_central, double fr_graz)
{
    if (!c) return TM_E_ARG;
    if (!veg || !pg || !pc || !central || veg->L != c->L || pg->L != c->L || pc->L != c->L || n_central < 1 || n_central > 10000 ||
        !(fr_graz >= 0 && fr_graz <= 1) || !(pg->sigma > 0) || !(veg->sigma0 > 0) || !(veg->dx > 0)) {
        c->err = "tm_cons_step2_setup: bad argument (all three L must equal the context's; 1 <= n_central <= 10000; 0 <= fr <= 1; sigma, dx > 0)";
        return TM_E_ARG;
    }
    CONS_CK(cudaSetDevice(c->device));
    const size_t nn = (size_t)c->L * c->L;
    if (!c->d_star) CONS_CK(cudaMalloc(&c->d_star, 2 * nn * sizeof(double)));
    if (!c->d_veff) CONS_CK(cudaMalloc(&c->d_veff, 2 * nn * sizeof(double)));
    if (!c->d_part) CONS_CK(cudaMalloc(&c->d_part, sizeof(double) * 3 * kConsReduceBlocks));
    if (!c->d_dyn) CONS_CK(cudaMalloc(&c->d_dyn, sizeof(Step2Dyn)));
    if (!c->d_ovf) CONS_CK(cudaMalloc(&c->d_ovf, sizeof(unsigned long long)));
    cudaFree(c->d_central); c->d_central = nullptr;
    CONS_CK(cudaMalloc(&c->d_central, sizeof(int) * 2 * n_central));
    CONS_CK(cudaMemcpyAsync(c->d_central, central, sizeof(int) * 2 * n_central, cudaMemcpyHostToDevice, c->stream));
    CONS_CK(cudaMemsetAsync(c->d_ovf, 0, sizeof(unsigned long long), c->stream));
    CONS_CK(cudaMemsetAsync(c->d_veff, 0, 2 * nn * sizeof(double), c->stream));
    CONS_CK(cudaMemsetAsync(c->d_gamma, 0, nn * sizeof(double), c->stream));
    CONS_CK(cudaFuncSetAttribute(cons_gamma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 10000 * (int)sizeof(int)));
    CONS_CK(cudaStreamSynchronize(c->stream));   // `central` is the caller's again
    const ConsConsts cc = cons_consts(pc->dt, pc->a, pc->c, pc->gmax, pc->alpha, pc->rW, pc->W0, pc->K1, pc->K2, pc->Dw, pc->Do,
                                      pc->A01, pc->H01, pc->E1);
    c->k2 = step2_consts(c->L, n_central, fr_graz, veg->dt, veg->dx, veg->d, veg->D0, veg->sigma0, veg->clamp_eps, veg->clamp_lo,
                         pg->D, pg->v, pg->M, pg->sigma, pg->D0_floor, cc);
    c->step2_ready = true;
    return TM_OK;
}

int tm_cons_step2_set_fpe(tm_cons_ctx *c, const tm_fke_terms *t, const double *gauss_D, const double *gauss_V, double M1)
{
    if (!c) return TM_E_ARG;
    if (!c->step2_ready) { c->err = "tm_cons_step2_set_fpe: call tm_cons_step2_setup first"; return TM_E_ARG; }
    if (!t) {                                      // back to the Rietkerk_FD flavour
        c->k2.fpe = 0; c->k2.c.fpe = 0; c->k2.c.M1 = 0.0;
        return TM_OK;
    }
    const int vmax = fke_max3(fke_max3(t->v_row_lo, t->v_row_hi, t->v_col_lo), t->v_col_hi, 0);
    if (!gauss_D || !gauss_V || t->bound_D < 0 || t->v_row_lo < 0 || t->v_row_hi < 0 || t->v_col_lo < 0 || t->v_col_hi < 0 ||
        t->bound_D > kFkeMaxBound || vmax > kFkeMaxBound || t->n_gauss_D < t->bound_D + 1 || t->n_gauss_V < vmax + 1) {
        c->err = "tm_cons_step2_set_fpe: window bounds must lie in [0, 512] and the stencils must cover them";
        return TM_E_ARG;
    }
    CONS_CK(cudaSetDevice(c->device));
    CONS_CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_fke_gD); cudaFree(c->d_fke_gV); c->d_fke_gD = c->d_fke_gV = nullptr;
    CONS_CK(cudaMalloc(&c->d_fke_gD, t->n_gauss_D * sizeof(double)));
    CONS_CK(cudaMalloc(&c->d_fke_gV, t->n_gauss_V * sizeof(double)));
    CONS_CK(cudaMemcpy(c->d_fke_gD, gauss_D, t->n_gauss_D * sizeof(double), cudaMemcpyHostToDevice));
    CONS_CK(cudaMemcpy(c->d_fke_gV, gauss_V, t->n_gauss_V * sizeof(double), cudaMemcpyHostToDevice));
    FkeTerms &f = c->k2.fke;
    f.L = c->L; f.bD = t->bound_D; f.rlo = t->v_row_lo; f.rhi = t->v_row_hi; f.clo = t->v_col_lo; f.chi = t->v_col_hi;
    f.pref = t->mu_vel_prefactor; f.eps = t->epsilon;
    c->k2.fpe = 1; c->k2.c.fpe = 1; c->k2.c.M1 = M1;
    return TM_OK;
}

int tm_cons_step2_set_rk4_stages(tm_cons_ctx *c, int on)
{
    if (!c) return TM_E_ARG;
    c->rk4_stages = on ? 1 : 0;
    return TM_OK;
}

int tm_cons_step2_set_strict_means(tm_cons_ctx *c, int on)
{
    if (!c) return TM_E_ARG;
    c->strict_means = on ? 1 : 0;
    return TM_OK;
}

int tm_cons_step2_upload(tm_cons_ctx *c, const double *fields, const double *gamma)
{
    if (!c) return TM_E_ARG;
    if (!c->step2_ready || !fields) { c->err = "tm_cons_step2_upload: call tm_cons_step2_setup first; fields must not be NULL"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * sizeof(double);
    CONS_CK(cudaSetDevice(c->device));
    CONS_CK(cudaMemcpyAsync(c->d_state, fields, 4 * n, cudaMemcpyHostToDevice, c->stream));
    if (gamma) CONS_CK(cudaMemcpyAsync(c->d_gamma, gamma, n, cudaMemcpyHostToDevice, c->stream));
    else CONS_CK(cudaMemsetAsync(c->d_gamma, 0, n, c->stream));
    CONS_CK(cudaMemsetAsync(c->d_ovf, 0, sizeof(unsigned long long), c->stream));
    CONS_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}


int tm_cons_step2_run(tm_cons_ctx *c, uint64_t seed, uint32_t trial, uint32_t step0, int nsteps)
{
    if (!c) return TM_E_ARG;
    if (!c->step2_ready || nsteps < 0) { c->err = "tm_cons_step2_run: call tm_cons_step2_setup first; nsteps >= 0"; return TM_E_ARG; }
    CONS_CK(cudaSetDevice(c->device));
    Step2Buffers b = step2_buffers(c);
    CONS_CK(cudaEventRecord(c->ev0, c->stream));
    for (int s = 0; s < nsteps; s++) cons_step2_enqueue(CudaLaunch{c->stream}, c->k2, b, seed, trial, step0 + (uint32_t)s);
    c->d_state = b.state; c->d_ya = b.ya;            // the fused RK4 ping-pongs the state between these two
    CONS_CK(cudaGetLastError());
    CONS_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = step2_launches(b, c->k2.fpe) * nsteps;
    return TM_OK;
}

int tm_cons_step2_replay(tm_cons_ctx *c, const double *u_veg, const double *u_graz, uint32_t stride)
{
    if (!c) return TM_E_ARG;
    if (!c->step2_ready || !u_veg || !u_graz || stride < 1) { c->err = "tm_cons_step2_replay: call tm_cons_step2_setup first; both uniform arrays and stride >= 1 are required"; return TM_E_ARG; }
    CONS_CK(cudaSetDevice(c->device));
    const size_t need = (size_t)c->L * c->L * stride * sizeof(double);
    if (2 * need > c->replay_cap) {
        cudaFree(c->d_replay); c->d_replay = nullptr; c->replay_cap = 0;
        CONS_CK(cudaMalloc(&c->d_replay, 2 * need));
        c->replay_cap = 2 * need;
    }
    CONS_CK(cudaMemcpyAsync(c->d_replay, u_veg, need, cudaMemcpyHostToDevice, c->stream));
    CONS_CK(cudaMemcpyAsync((char *)c->d_replay + need, u_graz, need, cudaMemcpyHostToDevice, c->stream));
    Step2Buffers b = step2_buffers(c);
    b.replay_veg = c->d_replay; b.replay_graz = (const double *)((char *)c->d_replay + need); b.stride = stride;
    CONS_CK(cudaEventRecord(c->ev0, c->stream));
    cons_step2_enqueue(CudaLaunch{c->stream}, c->k2, b, 0, 0, 0);
    c->d_state = b.state; c->d_ya = b.ya;
    CONS_CK(cudaGetLastError());
    CONS_CK(cudaEventRecord(c->ev1, c->stream));
    c->n_launches = step2_launches(b, c->k2.fpe);
    CONS_CK(cudaStreamSynchronize(c->stream));   // the uniform arrays are the caller's again
    return TM_OK;
}

int tm_cons_step2_download(tm_cons_ctx *c, double *fields, double *gamma, double *v_eff, double *means, uint64_t *overflow)
{
    if (!c) return TM_E_ARG;
    if (!c->step2_ready) { c->err = "tm_cons_step2_download: call tm_cons_step2_setup first"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * sizeof(double);
    CONS_CK(cudaSetDevice(c->device));
    unsigned long long ovf = 0;
    Step2Dyn d{};
    if (fields) CONS_CK(cudaMemcpyAsync(fields, c->d_state, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    if (gamma) CONS_CK(cudaMemcpyAsync(gamma, c->d_gamma, n, cudaMemcpyDeviceToHost, c->stream));
    if (v_eff) CONS_CK(cudaMemcpyAsync(v_eff, c->d_veff, 2 * n, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaMemcpyAsync(&d, c->d_dyn, sizeof d, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaMemcpyAsync(&ovf, c->d_ovf, sizeof ovf, cudaMemcpyDeviceToHost, c->stream));
    CONS_CK(cudaStreamSynchronize(c->stream));
    if (means) {
        const double cells = (double)c->L * (double)c->L;
        means[0] = d.sum_veg / cells; means[1] = d.sum_graz / cells; means[2] = d.occ_veg; means[3] = (double)d.eff;
    }
    if (overflow) *overflow = ovf;
    return TM_OK;
}

int tm_cons_last_ms(tm_cons_ctx *c, float *ms, int *n_launches)
{
    if (!c || !ms) return TM_E_ARG;
    CONS_CK(cudaSetDevice(c->device));
    CONS_CK(cudaEventSynchronize(c->ev1));
    CONS_CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (n_launches) *n_launches = c->n_launches;
    return TM_OK;
}

} // extern "C"
