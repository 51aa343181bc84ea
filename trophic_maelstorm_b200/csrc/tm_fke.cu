// tm_fke.cu -- C ABI (include/tm_fke.h) over the FKE movement kernels (tm_fke.cuh).  sm_100a only; no CPU path.
#include "../../include/tm_fke.h"
#include "../../include/tm_lattice.h" // TM_E_* codes
#include "tm_fke.cuh"

#include <cuda_runtime.h>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

using namespace tml;

static thread_local std::string g_fke_create_error;

struct tm_fke_ctx {
    int L = 0, n_species = 0, device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<tm_fke_terms> terms;
    std::vector<char> have_terms;
    std::vector<double *> d_gD, d_gV;        // per species
    double *d_s[3] = {nullptr, nullptr, nullptr}, *d_U[2] = {nullptr, nullptr};
    int *d_disp = nullptr, *d_maxdisp = nullptr;   // maxdisp: [n_species][2]
    double *d_rho = nullptr, *d_gamma = nullptr, *d_veff = nullptr, *d_out = nullptr; // staging of the host-buffer form
    int n_launches = 0;
    std::string err;
};

#define FKE_CK(call)                                                                                       \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                   \
            return TM_E_CUDA;                                                                              \
        }                                                                                                  \
    } while (0)

extern "C" {

const char *tm_fke_last_error(const tm_fke_ctx *c) { return c ? c->err.c_str() : g_fke_create_error.c_str(); }

int tm_fke_destroy(tm_fke_ctx *c)
{
    if (!c) return TM_E_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (double *p : c->d_gD) cudaFree(p);
    for (double *p : c->d_gV) cudaFree(p);
    for (double *p : c->d_s) cudaFree(p);
    for (double *p : c->d_U) cudaFree(p);
    cudaFree(c->d_disp); cudaFree(c->d_maxdisp);
    cudaFree(c->d_rho); cudaFree(c->d_gamma); cudaFree(c->d_veff); cudaFree(c->d_out);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return TM_OK;
}

static int fke_create_fail(tm_fke_ctx *c, int code, const std::string &msg)
{
    g_fke_create_error = msg;
    if (c) tm_fke_destroy(c);
    return code;
}

int tm_fke_create(int L, int n_species, int device, tm_fke_ctx **out)
{
    if (!out) return TM_E_ARG;
    *out = nullptr;
    if (L < 2 || L > 32767 || n_species < 1 || n_species > 8)
        return fke_create_fail(nullptr, TM_E_ARG, "tm_fke_create: need 2 <= L <= 32767 and 1 <= n_species <= 8");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fke_create_fail(nullptr, TM_E_CUDA, "tm_fke_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fke_create_fail(nullptr, TM_E_ARG, "tm_fke_create: bad device index");
    tm_fke_ctx *c = new (std::nothrow) tm_fke_ctx;
    if (!c) return fke_create_fail(nullptr, TM_E_ARG, "tm_fke_create: out of host memory");
    c->L = L; c->n_species = n_species; c->device = device;
    c->terms.resize(n_species); c->have_terms.assign(n_species, 0);
    c->d_gD.assign(n_species, nullptr); c->d_gV.assign(n_species, nullptr);
    const size_t n = (size_t)L * L;
#define CR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fke_create_fail(c, TM_E_CUDA, std::string("tm_fke_create: " #call ": ") + cudaGetErrorString(e_)); } while (0)
    CR(cudaSetDevice(device));
    cudaDeviceProp prop;
    CR(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    CR(cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault));
    CR(cudaEventCreate(&c->ev0)); CR(cudaEventCreate(&c->ev1));
    for (int k = 0; k < 3; k++) CR(cudaMalloc(&c->d_s[k], n * sizeof(double)));
    for (int k = 0; k < 2; k++) CR(cudaMalloc(&c->d_U[k], n * sizeof(double)));
    CR(cudaMalloc(&c->d_disp, n * sizeof(int)));
    CR(cudaMalloc(&c->d_maxdisp, 2 * n_species * sizeof(int)));
    CR(cudaMemset(c->d_maxdisp, 0, 2 * n_species * sizeof(int)));
#undef CR
    *out = c;
    return TM_OK;
}

static int fke_set_terms(tm_fke_ctx *c, int s, const tm_fke_terms *t, const double *gauss_D, const double *gauss_V, cudaMemcpyKind kind);

int tm_fke_set_terms(tm_fke_ctx *c, int s, const tm_fke_terms *t, const double *gauss_D, const double *gauss_V)
{
    return fke_set_terms(c, s, t, gauss_D, gauss_V, cudaMemcpyHostToDevice);
}

int tm_fke_set_terms_device(tm_fke_ctx *c, int s, const tm_fke_terms *t, const double *d_gauss_D, const double *d_gauss_V)
{
    return fke_set_terms(c, s, t, d_gauss_D, d_gauss_V, cudaMemcpyDeviceToDevice);
}

static int fke_set_terms(tm_fke_ctx *c, int s, const tm_fke_terms *t, const double *gauss_D, const double *gauss_V, cudaMemcpyKind kind)
{
    if (!c) return TM_E_ARG;
    if (!t || !gauss_D || !gauss_V || s < 0 || s >= c->n_species) { c->err = "tm_fke_set_terms: bad argument"; return TM_E_ARG; }
    const int vmax = fke_max3(fke_max3(t->v_row_lo, t->v_row_hi, t->v_col_lo), t->v_col_hi, 0);
    if (t->bound_D < 0 || t->v_row_lo < 0 || t->v_row_hi < 0 || t->v_col_lo < 0 || t->v_col_hi < 0 ||
        t->bound_D > kFkeMaxBound || vmax > kFkeMaxBound || t->n_gauss_D < t->bound_D + 1 || t->n_gauss_V < vmax + 1) {
        c->err = "tm_fke_set_terms: window bounds must lie in [0, 512] and the stencils must cover them";
        return TM_E_ARG;
    }
    FKE_CK(cudaSetDevice(c->device));
    FKE_CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_gD[s]); cudaFree(c->d_gV[s]); c->d_gD[s] = c->d_gV[s] = nullptr;
    FKE_CK(cudaMalloc(&c->d_gD[s], t->n_gauss_D * sizeof(double)));
    FKE_CK(cudaMalloc(&c->d_gV[s], t->n_gauss_V * sizeof(double)));
    FKE_CK(cudaMemcpy(c->d_gD[s], gauss_D, t->n_gauss_D * sizeof(double), kind));
    FKE_CK(cudaMemcpy(c->d_gV[s], gauss_V, t->n_gauss_V * sizeof(double), kind));
    c->terms[s] = *t; c->have_terms[s] = 1;
    return TM_OK;
}

static int fke_launch_all(tm_fke_ctx *c, const double *d_rho, const double *d_gamma, const double *d_veff, double *d_out)
{
    const size_t n = (size_t)c->L * c->L;
    for (int s = 0; s < c->n_species; s++)
        if (!c->have_terms[s]) { c->err = "tm_fke_move: tm_fke_set_terms has not been called for every species"; return TM_E_STATE; }
    FKE_CK(cudaSetDevice(c->device));
    FKE_CK(cudaEventRecord(c->ev0, c->stream));
    FKE_CK(cudaMemsetAsync(c->d_maxdisp, 0, 2 * c->n_species * sizeof(int), c->stream));
    c->n_launches = 0;
    const int tiles = (c->L + kFkeThreads - 1) / kFkeThreads;
    for (int s = 0; s < c->n_species; s++) {
        const tm_fke_terms &h = c->terms[s];
        FkeArgs A{};
        A.t.L = c->L; A.t.bD = h.bound_D; A.t.rlo = h.v_row_lo; A.t.rhi = h.v_row_hi; A.t.clo = h.v_col_lo; A.t.chi = h.v_col_hi;
        A.t.pref = h.mu_vel_prefactor; A.t.eps = h.epsilon;
        A.rho = d_rho + s * n; A.gamma = d_gamma + s * n; A.veff = d_veff + 2 * s * n; A.out = d_out + s * n;
        A.sA = c->d_s[0]; A.sB = c->d_s[1]; A.sD = c->d_s[2]; A.disp = c->d_disp; A.maxdisp = c->d_maxdisp + 2 * s;
        A.U1 = c->d_U[0]; A.U2 = c->d_U[1]; A.gD = c->d_gD[s]; A.gV = c->d_gV[s];
        fke_prepare_kernel<<<(unsigned)c->L, 256, 0, c->stream>>>(A);   // one CTA per lattice row
        const size_t smem_c = fke_cols_smem_doubles(A.t) * sizeof(double);
        const size_t smem_r = (size_t)(A.t.bD + 1 + fke_max3(A.t.rlo, A.t.rhi, 0) + 1) * sizeof(double);
        fke_cols_kernel<<<(unsigned)((long long)tiles * c->L), kFkeThreads, smem_c, c->stream>>>(A);
        fke_rows_kernel<<<(unsigned)((long long)tiles * c->L), kFkeThreads, smem_r, c->stream>>>(A);
        c->n_launches += 3;
    }
    FKE_CK(cudaGetLastError());
    FKE_CK(cudaEventRecord(c->ev1, c->stream));
    return TM_OK;
}

int tm_fke_move_device(tm_fke_ctx *c, const double *d_rho, const double *d_gamma, const double *d_v_eff, double *d_rho_out)
{
    if (!c) return TM_E_ARG;
    if (!d_rho || !d_gamma || !d_v_eff || !d_rho_out) { c->err = "tm_fke_move_device: null pointer"; return TM_E_ARG; }
    return fke_launch_all(c, d_rho, d_gamma, d_v_eff, d_rho_out);
}

int tm_fke_move_host(tm_fke_ctx *c, const double *rho, const double *gamma, const double *v_eff, double *rho_out)
{
    if (!c) return TM_E_ARG;
    if (!rho || !gamma || !v_eff || !rho_out) { c->err = "tm_fke_move_host: null pointer"; return TM_E_ARG; }
    const size_t n = (size_t)c->L * c->L * c->n_species;
    FKE_CK(cudaSetDevice(c->device));
    if (!c->d_rho) {
        FKE_CK(cudaMalloc(&c->d_rho, n * sizeof(double)));
        FKE_CK(cudaMalloc(&c->d_gamma, n * sizeof(double)));
        FKE_CK(cudaMalloc(&c->d_veff, 2 * n * sizeof(double)));
        FKE_CK(cudaMalloc(&c->d_out, n * sizeof(double)));
    }
    FKE_CK(cudaMemcpyAsync(c->d_rho, rho, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FKE_CK(cudaMemcpyAsync(c->d_gamma, gamma, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FKE_CK(cudaMemcpyAsync(c->d_veff, v_eff, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int rc = fke_launch_all(c, c->d_rho, c->d_gamma, c->d_veff, c->d_out);
    if (rc != TM_OK) return rc;
    FKE_CK(cudaMemcpyAsync(rho_out, c->d_out, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FKE_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_fke_sync(tm_fke_ctx *c)
{
    if (!c) return TM_E_ARG;
    FKE_CK(cudaSetDevice(c->device));
    FKE_CK(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

void *tm_fke_cuda_stream(tm_fke_ctx *c) { return c ? (void *)c->stream : nullptr; }

int tm_fke_last_ms(tm_fke_ctx *c, float *ms, int *n_launches)
{
    if (!c || !ms) return TM_E_ARG;
    FKE_CK(cudaSetDevice(c->device));
    FKE_CK(cudaEventSynchronize(c->ev1));
    FKE_CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    if (n_launches) *n_launches = c->n_launches;
    return TM_OK;
}

int tm_fke_last_displacement(tm_fke_ctx *c, int s, int32_t out[2])
{
    if (!c || !out || s < 0 || s >= c->n_species) return TM_E_ARG;
    FKE_CK(cudaSetDevice(c->device));
    FKE_CK(cudaStreamSynchronize(c->stream));
    FKE_CK(cudaMemcpy(out, c->d_maxdisp + 2 * s, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    return TM_OK;
}

} // extern "C"
