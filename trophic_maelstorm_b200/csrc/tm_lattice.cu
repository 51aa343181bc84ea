// tm_lattice.cu -- the extern "C" layer of include/tm_lattice.h over the sm_100a kernels.
// Compiled with: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo
// (fused multiply-adds in the kernels are explicit; see tm_math.cuh).
#include "../../include/tm_lattice.h"
#include "tm_kernels.cuh"
#include "tm_host_setup.h"
#include "tm_dpdet.cuh"

#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace tml;

static thread_local std::string g_create_error;

constexpr int kTimedSteps = 64;

struct tm_ctx {
    tm_params p{};
    Geom g{};
    int device = 0, n_sm = 0, riet_occ = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *buf[2] = {nullptr, nullptr};
    int cur = 0;
    bool have_state = false, stats_valid = false, strict = false;
    std::vector<double> a_host;
    double *a_dev = nullptr;
    DpTrial *dp_trials_dev = nullptr;
    bool trial_consts_dirty = true;
    PartialStats *partials = nullptr;
    int n_records = 0, grid = 0, dp_blocks_per_trial = 0;
    TmStatsDev *stats_dev = nullptr;
    unsigned long long *counters = nullptr;
    double *replay_dev = nullptr;
    size_t replay_cap = 0;
    double *dornic_dev = nullptr;
    double *logtab_dev = nullptr;  // FAST-mode log table
    double *pstar_dev = nullptr;   // rho* scratch field of the split Dornic / RK4 kernel pair
    unsigned long long *spread_dev = nullptr; // [n_trials][3] accumulators of tm_measure_spread
    bool fused = false;            // default: Dornic sweep + RK4 kernel pair; true: the single fused kernel
    int rk4_kind = 1;              // RK4 kernel of the pair: 2 = riet_rk4_kernel (two columns per thread, tm_rk4.cuh), 1 = riet_step_kernel<kSrcPre>
    int rk4_occ = 1;               // resident CTAs per SM of riet_rk4_kernel
    int dornic_kind = 1;           // Dornic sweep: 2 = riet_dornic2_kernel (two columns per thread), 1 = riet_dornic_kernel
    double *halo_dev[2] = {nullptr, nullptr};
    // overlapped slab stepping (tm_slab_step_begin / _end): boundary bands and the halo traffic on their own high-priority stream
    cudaStream_t band_stream = nullptr;
    cudaEvent_t ev_band[2] = {nullptr, nullptr}, ev_inner[2] = {nullptr, nullptr}, ev_join = nullptr, ev_main = nullptr;
    double *pstar_band_dev = nullptr;   // rho* scratch of the band launches (the inner launch's rows overlap theirs by 3)
    int slab_parity = 0, slab_reserve_sm = 0;
    bool slab_open = false, slab_fresh = true;   // fresh: the main stream may hold work the band stream has not been ordered after
    uint64_t seed = 0, step = 0;
    uint32_t trial0 = 0;
    float last_ms = 0.f;
    int last_launches = 0, launches_this_call = 0;
    bool kernel_timing = false;    // tm_set_kernel_timing: events around each kernel of the pair
    std::vector<cudaEvent_t> kev;  // [3 * kTimedSteps]: before the Dornic sweep, between, after the RK4 kernel
    int kev_steps = 0;
    std::string err;
};

#define TM_FAIL(ctx, code, ...)                                             \
    do {                                                                    \
        char _b[512];                                                       \
        snprintf(_b, sizeof _b, __VA_ARGS__);                               \
        (ctx)->err = _b;                                                    \
        return (code);                                                      \
    } while (0)

#define TM_CUDA(ctx, call)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) TM_FAIL(ctx, TM_E_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

extern "C" {

int tm_abi_version(void) { return TM_ABI_VERSION; }

int tm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tm_device_mem(int device, uint64_t *free_bytes, uint64_t *total_bytes)
{
    if (!free_bytes || !total_bytes) return TM_E_ARG;
    int cur = 0;
    size_t f = 0, t = 0;
    if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) {
        cudaGetLastError();
        g_create_error = "tm_device_mem: no such CUDA device";
        return TM_E_CUDA;
    }
    cudaSetDevice(cur);
    *free_bytes = f; *total_bytes = t;
    return TM_OK;
}

int tm_ctx_bytes(const tm_params *p, int n_trials, uint64_t *bytes)
{
    if (!p || !bytes || n_trials < 1 || p->L < 8) return TM_E_ARG;
    Geom g{};
    make_geom(*p, n_trials, 0, p->L, &g);
    const uint64_t state = sizeof(double) * (uint64_t)g.n_trials * (uint64_t)g.trial_stride;
    const uint64_t pstar = is_riet(p->model) ? sizeof(double) * (uint64_t)g.n_trials * (uint64_t)g.field_stride : 0;
    *bytes = 2 * state + pstar + (uint64_t)n_trials * 512 + (1u << 20);   // two ping-pong sides, the rho* scratch field, per-trial records
    return TM_OK;
}

const char *tm_last_error(const tm_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int tm_halo_rows(void) { return kHalo; }

static int create_fail(tm_ctx *c, int code, const char *msg)
{
    g_create_error = msg;
    if (c) { g_create_error += c->err.empty() ? "" : (": " + c->err); tm_destroy(c); }
    return code;
}

static int alloc_all(tm_ctx *c)
{
    Geom &g = c->g;
    const size_t nbytes = sizeof(double) * (size_t)g.n_trials * g.trial_stride;
    for (int s = 0; s < 2; s++) {
        TM_CUDA(c, cudaMalloc(&c->buf[s], nbytes));
        TM_CUDA(c, cudaMemsetAsync(c->buf[s], 0, nbytes, c->stream));
    }
    TM_CUDA(c, cudaMalloc(&c->a_dev, sizeof(double) * g.n_trials));
    TM_CUDA(c, cudaMalloc(&c->dp_trials_dev, sizeof(DpTrial) * g.n_trials));
    TM_CUDA(c, cudaMalloc(&c->stats_dev, sizeof(TmStatsDev) * g.n_trials));
    {
        double tab[256];
        make_log_table(tab);
        TM_CUDA(c, cudaMalloc(&c->logtab_dev, sizeof tab));
        TM_CUDA(c, cudaMemcpy(c->logtab_dev, tab, sizeof tab, cudaMemcpyHostToDevice));
    }
    TM_CUDA(c, cudaMalloc(&c->counters, sizeof(unsigned long long) * 4));
    TM_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 4, c->stream));
    // launch geometry
    if (is_riet(c->p.model)) {
        riet_launch_shape(g, c->n_sm * c->riet_occ, &c->grid);
        c->n_records = 0;
    } else {
        // DP kernels take whole rows per block: at most one block per row, ~8 blocks per SM overall
        long long bpt = g.nrows;
        const long long cap = (8LL * c->n_sm + g.n_trials - 1) / g.n_trials;
        if (bpt > cap) bpt = cap;
        if (bpt < 1) bpt = 1;
        c->dp_blocks_per_trial = (int)bpt;
        c->n_records = (int)bpt * g.n_trials;
    }
    // the stand-alone stats kernel uses its own (possibly larger) record set
    {
        long long bpt = g.nrows;
        const long long cap = (4LL * c->n_sm + g.n_trials - 1) / g.n_trials;
        if (bpt > cap) bpt = cap;
        if (bpt < 1) bpt = 1;
        const int need = (int)bpt * g.n_trials;
        if (need > c->n_records) c->n_records = need;
    }
    TM_CUDA(c, cudaMalloc(&c->partials, sizeof(PartialStats) * c->n_records));
    const size_t halo_bytes = sizeof(double) * (size_t)g.n_trials * g.n_fields * kHalo * g.Lp;
    for (int s = 0; s < 2; s++) TM_CUDA(c, cudaMalloc(&c->halo_dev[s], halo_bytes));
    return TM_OK;
}

int tm_create_slab(const tm_params *p, int n_trials, int device, int row0, int nrows, tm_ctx **out)
{
    if (!p || !out) return create_fail(nullptr, TM_E_ARG, "tm_create: null argument");
    *out = nullptr;
    if (p->model < 0 || p->model > 3) return create_fail(nullptr, TM_E_ARG, "tm_create: unknown model");
    if (p->L < 8 || p->L > 65535) return create_fail(nullptr, TM_E_ARG, "tm_create: L must be in [8, 65535]");
    if (n_trials < 1) return create_fail(nullptr, TM_E_ARG, "tm_create: n_trials < 1");
    if (row0 < 0 || nrows < 1 || row0 + nrows > p->L) return create_fail(nullptr, TM_E_ARG, "tm_create: bad row slab");
    if (nrows != p->L && nrows < kHalo) return create_fail(nullptr, TM_E_ARG, "tm_create: a slab needs at least 4 rows");
    if (!(p->dt > 0) || !(p->dx > 0)) return create_fail(nullptr, TM_E_ARG, "tm_create: dt and dx must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return create_fail(nullptr, TM_E_CUDA, "tm_create: no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return create_fail(nullptr, TM_E_ARG, "tm_create: bad device index");
    tm_ctx *c = new (std::nothrow) tm_ctx();
    if (!c) return create_fail(nullptr, TM_E_ARG, "tm_create: out of host memory");
    c->p = *p; c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) return create_fail(c, TM_E_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return create_fail(c, TM_E_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 9) return create_fail(c, TM_E_CUDA, "tm_create: device lacks TMA bulk copies (needs sm_90+, built for sm_100a)");
    c->n_sm = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return create_fail(c, TM_E_CUDA, "cudaStreamCreate failed");
    if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) return create_fail(c, TM_E_CUDA, "cudaEventCreate failed");

    Geom &g = c->g;
    make_geom(*p, n_trials, row0, nrows, &g);
    c->a_host.assign(n_trials, p->a);

    // opt in to the large dynamic shared memory the Rietkerk kernel needs
    if (is_riet(p->model)) {
        const int smem_s = (int)riet_smem_bytes(g.nthreads, false), smem_d = (int)riet_smem_bytes(g.nthreads, true);
        bool ok = true;
#define TM_OPTIN(K, B) ok = ok && (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, B) == cudaSuccess)
        TM_OPTIN((riet_step_kernel<false, false, kSrcPhilox, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, true, kSrcPhilox, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, false, kSrcReplay, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, true, kSrcReplay, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, false, kSrcPre, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, true, kSrcPre, false>), smem_s);
        TM_OPTIN((riet_step_kernel<false, true, kSrcPhilox, true>), smem_s);
        TM_OPTIN((riet_step_kernel<false, true, kSrcReplay, true>), smem_s);
        TM_OPTIN((riet_step_kernel<true, false, kSrcPhilox, false>), smem_d);
        TM_OPTIN((riet_step_kernel<true, true, kSrcPhilox, false>), smem_d);
#undef TM_OPTIN
        if (!ok) { cudaGetLastError(); return create_fail(c, TM_E_CUDA, "tm_create: cudaFuncSetAttribute(shared memory) failed"); }
        int occ = 1; // resident CTAs per SM of the production variant: the persistent grid fills exactly that
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, riet_step_kernel<false, false, kSrcPhilox, false>, g.nthreads,
                                                          p->model == TM_MODEL_RIETKERK_DET ? smem_d : smem_s) != cudaSuccess || occ < 1) {
            cudaGetLastError(); occ = 1;
        }
        c->riet_occ = occ;
        const int smem_r = (int)riet_rk4_smem_bytes();
        if (cudaFuncSetAttribute(riet_rk4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r) != cudaSuccess ||
            cudaFuncSetAttribute(riet_rk4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r) != cudaSuccess) {
            cudaGetLastError(); return create_fail(c, TM_E_CUDA, "tm_create: cudaFuncSetAttribute(shared memory, riet_rk4_kernel) failed");
        }
        int occ2 = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, riet_rk4_kernel<false>, kRkThreads, smem_r) != cudaSuccess || occ2 < 1) {
            cudaGetLastError(); occ2 = 1;
        }
        c->rk4_occ = occ2;
        if (const char *e = getenv("TM_RK4_KIND")) c->rk4_kind = atoi(e) == 1 ? 1 : 2;   // kernel-variant experiments only
        if (const char *e = getenv("TM_DORNIC_KIND")) c->dornic_kind = atoi(e) == 1 ? 1 : 2;
    }
    if (alloc_all(c) != TM_OK) return create_fail(c, TM_E_CUDA, "tm_create: device allocation failed");
    *out = c;
    return TM_OK;
}

int tm_create(const tm_params *p, int n_trials, int device, tm_ctx **out)
{
    return tm_create_slab(p, n_trials, device, 0, p ? p->L : 0, out);
}

int tm_destroy(tm_ctx *c)
{
    if (!c) return TM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int s = 0; s < 2; s++) { cudaFree(c->buf[s]); cudaFree(c->halo_dev[s]); }
    cudaFree(c->a_dev); cudaFree(c->dp_trials_dev); cudaFree(c->partials); cudaFree(c->stats_dev);
    cudaFree(c->counters); cudaFree(c->replay_dev); cudaFree(c->dornic_dev); cudaFree(c->pstar_dev); cudaFree(c->logtab_dev); cudaFree(c->spread_dev);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto &e : c->kev) cudaEventDestroy(e);
    if (c->band_stream) { cudaStreamSynchronize(c->band_stream); cudaStreamDestroy(c->band_stream); }
    for (int s = 0; s < 2; s++) { if (c->ev_band[s]) cudaEventDestroy(c->ev_band[s]); if (c->ev_inner[s]) cudaEventDestroy(c->ev_inner[s]); }
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    cudaFree(c->pstar_band_dev);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return TM_OK;
}

int tm_n_fields(const tm_ctx *c) { return c ? c->g.n_fields : TM_E_ARG; }

int tm_set_trial_a(tm_ctx *c, int trial, double a)
{
    if (!c) return TM_E_ARG;
    if (trial < 0 || trial >= c->g.n_trials) TM_FAIL(c, TM_E_ARG, "tm_set_trial_a: trial %d out of range", trial);
    c->a_host[trial] = a;
    c->trial_consts_dirty = true;
    return TM_OK;
}

int tm_set_strict(tm_ctx *c, int strict)
{
    if (!c) return TM_E_ARG;
    c->strict = strict != 0;
    return TM_OK;
}

int tm_set_fused(tm_ctx *c, int fused)
{
    if (!c) return TM_E_ARG;
    c->fused = fused != 0;
    return TM_OK;
}

static int push_trial_consts(tm_ctx *c)
{
    if (!c->trial_consts_dirty) return TM_OK;
    TM_CUDA(c, cudaMemcpyAsync(c->a_dev, c->a_host.data(), sizeof(double) * c->g.n_trials, cudaMemcpyHostToDevice, c->stream));
    if (!is_riet(c->p.model)) {
        std::vector<DpTrial> t(c->g.n_trials);
        for (int i = 0; i < c->g.n_trials; i++) t[i].dk = dornic_consts(c->p, c->a_host[i]);
        TM_CUDA(c, cudaMemcpyAsync(c->dp_trials_dev, t.data(), sizeof(DpTrial) * t.size(), cudaMemcpyHostToDevice, c->stream));
        TM_CUDA(c, cudaStreamSynchronize(c->stream)); // t goes out of scope
    }
    c->trial_consts_dirty = false;
    return TM_OK;
}

static int fill_ghost_cols(tm_ctx *c, double *buf)
{
    GhostArgs A{c->g, buf};
    const long long n = (long long)c->g.n_trials * c->g.n_fields * c->g.rows_alloc * 8;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8 * c->n_sm) blocks = 8 * c->n_sm;
    ghost_cols_kernel<<<blocks, 256, 0, c->stream>>>(A);
    TM_CUDA(c, cudaGetLastError());
    return TM_OK;
}

int tm_upload(tm_ctx *c, const double *host)
{
    if (!c || !host) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    const int roff = g.wrap ? 0 : kHalo;
    // all (trial, field) planes are equally spaced on both sides: one 2-D copy per plane
    for (int t = 0; t < g.n_trials; t++)
        for (int f = 0; f < g.n_fields; f++) {
            double *dst = c->buf[c->cur] + (long long)t * g.trial_stride + (long long)f * g.field_stride + (long long)roff * g.Lp + kHalo;
            const double *src = host + ((size_t)t * g.n_fields + f) * (size_t)g.nrows * g.L;
            TM_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(double) * g.Lp, src, sizeof(double) * g.L, sizeof(double) * g.L, g.nrows,
                                         cudaMemcpyHostToDevice, c->stream));
        }
    if (fill_ghost_cols(c, c->buf[c->cur]) != TM_OK) return TM_E_CUDA;
    c->have_state = true; c->stats_valid = false;
    return TM_OK;
}

int tm_download(tm_ctx *c, double *host)
{
    if (!c || !host) return TM_E_ARG;
    if (!c->have_state) TM_FAIL(c, TM_E_STATE, "tm_download: no state uploaded");
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    const int roff = g.wrap ? 0 : kHalo;
    for (int t = 0; t < g.n_trials; t++)
        for (int f = 0; f < g.n_fields; f++) {
            const double *src = c->buf[c->cur] + (long long)t * g.trial_stride + (long long)f * g.field_stride + (long long)roff * g.Lp + kHalo;
            double *dst = host + ((size_t)t * g.n_fields + f) * (size_t)g.nrows * g.L;
            TM_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(double) * g.L, src, sizeof(double) * g.Lp, sizeof(double) * g.L, g.nrows,
                                         cudaMemcpyDeviceToHost, c->stream));
        }
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_seed(tm_ctx *c, uint64_t seed, uint32_t trial0)
{
    if (!c) return TM_E_ARG;
    c->seed = seed; c->trial0 = trial0;
    return TM_OK;
}
int tm_set_step_counter(tm_ctx *c, uint64_t step) { if (!c) return TM_E_ARG; c->step = step; return TM_OK; }
uint64_t tm_step_counter(const tm_ctx *c) { return c ? c->step : 0; }

// The production pair (Dornic sweep into rho*, then the RK4 kernel) on lattice-local rows [ra, rb) of the context's slab, on
// `stream`, with at most n_slots resident CTAs for the RK4 kernel.  A row view is the same kernels on shifted base pointers
// with nrows = rb - ra and row0 advanced by ra (the Philox cell index is global, so the values do not depend on the cut); the
// whole slab is the view [0, nrows).  ev (3 events) brackets the two launches when kernel timing is on.
static int launch_pair_view(tm_ctx *c, const StepArgs &A0, int ra, int rb, double *pstar, cudaStream_t stream, int n_slots, cudaEvent_t *ev)
{
    StepArgs A = A0;
    Geom &g = A.g;
    const long long shift = (long long)ra * g.Lp;
    if (ra != 0 || rb != g.nrows) {
        if (g.wrap) TM_FAIL(c, TM_E_STATE, "row views need a slab context");
        g.row0 += ra; g.nrows = rb - ra;
    }
    A.in = A0.in + shift; A.out = A0.out + shift; A.pstar = pstar + shift;
    DornicArgs D{};
    D.g = g; D.dk = A.c.dk; D.inv_lambda = A.c.inv_lambda; D.mu_scale = A.c.mu_scale; D.in = A.in; D.pstar = pstar + shift;
    D.seed = c->seed; D.trial0 = c->trial0; D.step = (unsigned int)c->step; D.rk = A.rk; D.logtab = c->logtab_dev;
    D.row_lo = g.wrap ? 0 : -(kHalo - 1); D.row_hi = g.wrap ? g.nrows : g.nrows + (kHalo - 1);
    // rows per CTA: enough (trial, column block, run) items for ~2 rounds over 4 resident CTAs per SM
    const int dcols = c->dornic_kind == 2 ? kDornic2Cols : kDornicThreads;   // columns per CTA
    const int ncb = (g.L + dcols - 1) / dcols, nrow_d = D.row_hi - D.row_lo;
    long long want_runs = ((long long)c->n_sm * 8 + (long long)g.n_trials * ncb - 1) / ((long long)g.n_trials * ncb);
    if (want_runs < 1) want_runs = 1;
    int rpc = (int)((nrow_d + want_runs - 1) / want_runs);
    if (rpc < 16) rpc = nrow_d < 16 ? nrow_d : 16;
    D.rows_per_cta = rpc;
    const long long ditems = (long long)g.n_trials * ncb * ((nrow_d + rpc - 1) / rpc);
    long long dgrid = ditems;
    if (dgrid > (long long)c->n_sm * 64) dgrid = (long long)c->n_sm * 64;
    const size_t dsmem = riet_dornic_smem_bytes();
    if (ev) TM_CUDA(c, cudaEventRecord(ev[0], stream));
    if (c->dornic_kind == 2) {
        if (c->strict) riet_dornic2_kernel<true><<<(unsigned)dgrid, kDornic2Threads, dsmem, stream>>>(D);
        else           riet_dornic2_kernel<false><<<(unsigned)dgrid, kDornic2Threads, dsmem, stream>>>(D);
    } else {
        if (c->strict) riet_dornic_kernel<true><<<(unsigned)dgrid, kDornicThreads, dsmem, stream>>>(D);
        else           riet_dornic_kernel<false><<<(unsigned)dgrid, kDornicThreads, dsmem, stream>>>(D);
    }
    TM_CUDA(c, cudaGetLastError());
    if (ev) TM_CUDA(c, cudaEventRecord(ev[1], stream));
    if (c->rk4_kind == 2) {
        Rk4Args R{};
        R.g = g; R.c = A.c; R.in = A.in; R.pstar = A.pstar; R.out = A.out; R.a_trial = c->a_dev; R.counters = c->counters;
        int grid2 = 1;
        rk4_launch_shape(g, c->n_sm * c->rk4_occ, &R, &grid2);
        if (c->strict) riet_rk4_kernel<true><<<grid2, kRkThreads, riet_rk4_smem_bytes(), stream>>>(R);
        else           riet_rk4_kernel<false><<<grid2, kRkThreads, riet_rk4_smem_bytes(), stream>>>(R);
    } else {
        int grid = 1;
        riet_launch_shape(g, n_slots, &grid);
        const size_t smem = riet_smem_bytes(g.nthreads, false);
        if (c->strict) riet_step_kernel<false, true, kSrcPre, false><<<grid, g.nthreads, smem, stream>>>(A);
        else           riet_step_kernel<false, false, kSrcPre, false><<<grid, g.nthreads, smem, stream>>>(A);
    }
    TM_CUDA(c, cudaGetLastError());
    if (ev) TM_CUDA(c, cudaEventRecord(ev[2], stream));
    c->launches_this_call += 1;
    return TM_OK;
}

// one launch: in = buf[cur] -> out = buf[cur^1]
static int launch_step(tm_ctx *c, bool stats, bool replay, uint32_t stride, bool dornic_only)
{
    const Geom &g = c->g;
    const tm_params &p = c->p;
    if (stats && !is_riet(p.model)) TM_CUDA(c, cudaMemsetAsync(c->partials, 0xFF, sizeof(PartialStats) * c->n_records, c->stream));
    if (is_riet(p.model)) {
        StepArgs A{};
        A.g = g; A.c = riet_consts(p);
        A.in = c->buf[c->cur]; A.out = c->buf[c->cur ^ 1];
        A.rk = philox_keys(c->seed); A.logtab = c->logtab_dev;
        A.a_trial = c->a_dev; A.seed = c->seed; A.trial0 = c->trial0; A.step = (unsigned int)c->step;
        A.replay_u = c->replay_dev; A.stride = stride;
        A.counters = c->counters; A.dornic_out = c->dornic_dev;
        const bool det = p.model == TM_MODEL_RIETKERK_DET;
        const size_t smem = riet_smem_bytes(g.nthreads, det);
        const dim3 grid(c->grid), block(g.nthreads);
#define TM_LAUNCH(...) riet_step_kernel<__VA_ARGS__><<<grid, block, smem, c->stream>>>(A)
        (void)stats; // Rietkerk reductions run as stats_kernel on demand (tm_measure)
        if (det) {
            if (c->strict) TM_LAUNCH(true, true, kSrcPhilox, false); else TM_LAUNCH(true, false, kSrcPhilox, false);
        } else if (dornic_only) {
            // the Dornic-only entry point is a checking aid: always the STRICT (IEEE) arithmetic
            if (replay) TM_LAUNCH(false, true, kSrcReplay, true); else TM_LAUNCH(false, true, kSrcPhilox, true);
        } else if (replay) {
            if (c->strict) TM_LAUNCH(false, true, kSrcReplay, false); else TM_LAUNCH(false, false, kSrcReplay, false);
        } else if (c->fused) {
            if (c->strict) TM_LAUNCH(false, true, kSrcPhilox, false); else TM_LAUNCH(false, false, kSrcPhilox, false);
        } else {
            // Dornic sweep at full occupancy into a rho* scratch field, then the RK4 fronts only
            if (!c->pstar_dev) {
                const size_t nb = sizeof(double) * (size_t)g.n_trials * g.field_stride;
                TM_CUDA(c, cudaMalloc(&c->pstar_dev, nb));
                TM_CUDA(c, cudaMemsetAsync(c->pstar_dev, 0, nb, c->stream));
            }
            const bool timed = c->kernel_timing && c->kev_steps < kTimedSteps;
            const int rc = launch_pair_view(c, A, 0, g.nrows, c->pstar_dev, c->stream, c->n_sm * c->riet_occ, timed ? &c->kev[3 * c->kev_steps] : nullptr);
            if (rc != TM_OK) return rc;
            if (timed) c->kev_steps++;
        }
#undef TM_LAUNCH
    } else {
        DpArgs A{};
        A.g = g;
        A.c.dt = p.dt; A.c.dt2 = p.dt / 2.0; A.c.dt6 = p.dt / 6.0; A.c.b = p.b; A.c.c = p.c; // P/:1780
        A.in = c->buf[c->cur]; A.out = c->buf[c->cur ^ 1];
        A.trials = c->dp_trials_dev; A.seed = c->seed; A.trial0 = c->trial0; A.step = (unsigned int)c->step;
        A.replay_u = c->replay_dev; A.stride = stride;
        A.partials = c->partials; A.counters = c->counters; A.dornic_out = c->dornic_dev;
        A.blocks_per_trial = c->dp_blocks_per_trial;
        const dim3 grid(c->dp_blocks_per_trial * g.n_trials), block(256);
        const size_t smem = sizeof(double) * 8 * 256;
#define TM_LAUNCH(M, ...) dp_step_kernel<M, __VA_ARGS__><<<grid, block, smem, c->stream>>>(A)
        if (p.model == TM_MODEL_DP) {
            if (dornic_only) { if (replay) TM_LAUNCH(kDP, true, false, true); else TM_LAUNCH(kDP, false, false, true); }
            else if (replay) TM_LAUNCH(kDP, true, true, false);
            else if (stats) TM_LAUNCH(kDP, false, true, false);
            else TM_LAUNCH(kDP, false, false, false);
        } else {
            if (dornic_only) { if (replay) TM_LAUNCH(kDPCubic, true, false, true); else TM_LAUNCH(kDPCubic, false, false, true); }
            else if (replay) TM_LAUNCH(kDPCubic, true, true, false);
            else if (stats) TM_LAUNCH(kDPCubic, false, true, false);
            else TM_LAUNCH(kDPCubic, false, false, false);
        }
#undef TM_LAUNCH
    }
    TM_CUDA(c, cudaGetLastError());
    return TM_OK;
}

static int check_slab_step(tm_ctx *c, const char *who)
{
    if (!c->have_state) TM_FAIL(c, TM_E_STATE, "%s: no state uploaded", who);
    return TM_OK;
}

int tm_step(tm_ctx *c, int nsteps)
{
    if (!c) return TM_E_ARG;
    if (nsteps < 0) TM_FAIL(c, TM_E_ARG, "tm_step: nsteps < 0");
    if (check_slab_step(c, "tm_step") != TM_OK) return TM_E_STATE;
    TM_CUDA(c, cudaSetDevice(c->device));
    if (push_trial_consts(c) != TM_OK) return TM_E_CUDA;
    if (!c->g.wrap && nsteps > 1) TM_FAIL(c, TM_E_ARG, "tm_step: a row slab advances one step per halo exchange");
    c->launches_this_call = 0;
    c->kev_steps = 0;
    TM_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int s = 0; s < nsteps; s++) {
        const int rc = launch_step(c, s == nsteps - 1, false, 0, false);
        if (rc != TM_OK) return rc;
        c->cur ^= 1; c->step++;
    }
    TM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    c->last_launches = nsteps + c->launches_this_call;
    if (nsteps > 0) c->stats_valid = !is_riet(c->p.model); // DP kernels fuse the reductions; Rietkerk runs stats_kernel on demand
    return TM_OK;
}

static int ensure_replay(tm_ctx *c, size_t n)
{
    if (n <= c->replay_cap) return TM_OK;
    cudaFree(c->replay_dev); c->replay_dev = nullptr; c->replay_cap = 0;
    TM_CUDA(c, cudaMalloc(&c->replay_dev, sizeof(double) * n));
    c->replay_cap = n;
    return TM_OK;
}

int tm_step_replay(tm_ctx *c, const double *host_uniforms, uint64_t stride)
{
    if (!c || !host_uniforms) return TM_E_ARG;
    if (c->p.model == TM_MODEL_RIETKERK_DET) TM_FAIL(c, TM_E_ARG, "tm_step_replay: the noise-off model draws no uniforms");
    if (stride < 1 || stride > (1u << 20)) TM_FAIL(c, TM_E_ARG, "tm_step_replay: bad stride");
    if (check_slab_step(c, "tm_step_replay") != TM_OK) return TM_E_STATE;
    TM_CUDA(c, cudaSetDevice(c->device));
    if (push_trial_consts(c) != TM_OK) return TM_E_CUDA;
    const size_t n = (size_t)c->g.n_trials * c->g.L * c->g.L * stride;
    if (ensure_replay(c, n) != TM_OK) return TM_E_CUDA;
    TM_CUDA(c, cudaMemcpyAsync(c->replay_dev, host_uniforms, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    TM_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    const int rc = launch_step(c, true, true, (uint32_t)stride, false);
    if (rc != TM_OK) return rc;
    TM_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    c->cur ^= 1; c->step++;
    c->last_launches = 1; c->stats_valid = !is_riet(c->p.model);
    return TM_OK;
}

int tm_dornic_only(tm_ctx *c, double *host_out)
{
    if (!c || !host_out) return TM_E_ARG;
    if (c->p.model == TM_MODEL_RIETKERK_DET) TM_FAIL(c, TM_E_ARG, "tm_dornic_only: the noise-off model has no Dornic sub-step");
    if (check_slab_step(c, "tm_dornic_only") != TM_OK) return TM_E_STATE;
    TM_CUDA(c, cudaSetDevice(c->device));
    if (push_trial_consts(c) != TM_OK) return TM_E_CUDA;
    const size_t n = (size_t)c->g.n_trials * c->g.nrows * c->g.L;
    if (!c->dornic_dev) TM_CUDA(c, cudaMalloc(&c->dornic_dev, sizeof(double) * n));
    const int rc = launch_step(c, false, false, 0, true);
    if (rc != TM_OK) return rc;
    TM_CUDA(c, cudaMemcpyAsync(host_out, c->dornic_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_halo_bytes(const tm_ctx *c, uint64_t *bytes)
{
    if (!c || !bytes) return TM_E_ARG;
    *bytes = sizeof(double) * (uint64_t)c->g.n_trials * c->g.n_fields * kHalo * c->g.Lp;
    return TM_OK;
}

// which = 0: this slab's first 4 rows (the upper neighbour's bottom ghosts); 1: its last 4 rows
int tm_halo_pack(tm_ctx *c, int which, void **dev_buf)
{
    if (!c || !dev_buf || which < 0 || which > 1) return TM_E_ARG;
    if (c->g.wrap) TM_FAIL(c, TM_E_STATE, "tm_halo_pack: not a slab context");
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    const long long r = which == 0 ? kHalo : g.nrows; // physical row of the first packed row
    TM_CUDA(c, cudaMemcpy2DAsync(c->halo_dev[which], sizeof(double) * kHalo * g.Lp,
                                 c->buf[c->cur] + r * g.Lp, sizeof(double) * g.field_stride,
                                 sizeof(double) * kHalo * g.Lp, (size_t)g.n_trials * g.n_fields,
                                 cudaMemcpyDeviceToDevice, c->stream));
    *dev_buf = c->halo_dev[which];
    return TM_OK;
}

// which = 0: fill the ghost rows ABOVE row 0; 1: the ghost rows BELOW the last row
int tm_halo_unpack(tm_ctx *c, int which, const void *dev_buf)
{
    if (!c || !dev_buf || which < 0 || which > 1) return TM_E_ARG;
    if (c->g.wrap) TM_FAIL(c, TM_E_STATE, "tm_halo_unpack: not a slab context");
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    const long long r = which == 0 ? 0 : g.nrows + kHalo;
    TM_CUDA(c, cudaMemcpy2DAsync(c->buf[c->cur] + r * g.Lp, sizeof(double) * g.field_stride,
                                 dev_buf, sizeof(double) * kHalo * g.Lp,
                                 sizeof(double) * kHalo * g.Lp, (size_t)g.n_trials * g.n_fields,
                                 cudaMemcpyDeviceToDevice, c->stream));
    return TM_OK;
}

// ---- overlapped slab stepping (BASELINE config 5) ------------------------------------------------------------
// One step of a row slab as three pieces: the two boundary bands (first and last kBand rows) on a high-priority stream, the halo
// traffic of the NEW state behind them on the same stream (the caller's NCCL send/recv, tm_slab.cpp), and the inner rows on the
// context's main stream at the same time.  Step n, state n -> n + 1 (buffers alternate):
//     band stream:  wait inner(n-1) | band(n) | record ev_band[n&1] | <caller: send new edge rows, receive new ghost rows>
//     main stream:  wait band(n-1)  | inner(n) | record ev_inner[n&1]
// band(n) reads rows -4 .. kBand+3 of state n: its ghost rows arrived in order on the band stream, rows kBand.. come from
// inner(n-1).  inner(n) reads rows kBand-4 ..: band rows of state n come from band(n-1).  Writes into the buffer that held
// state n-1 are behind the same two waits.
constexpr int kBand = 8;

int tm_slab_band_rows(void) { return kBand; }

static int slab_prepare(tm_ctx *c)
{
    if (c->band_stream) return TM_OK;
    int lo = 0, hi = 0;
    TM_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TM_CUDA(c, cudaStreamCreateWithPriority(&c->band_stream, cudaStreamNonBlocking, hi));
    for (int s = 0; s < 2; s++) {
        TM_CUDA(c, cudaEventCreateWithFlags(&c->ev_band[s], cudaEventDisableTiming));
        TM_CUDA(c, cudaEventCreateWithFlags(&c->ev_inner[s], cudaEventDisableTiming));
    }
    TM_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    TM_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    const size_t nb = sizeof(double) * (size_t)c->g.n_trials * c->g.field_stride;
    if (!c->pstar_dev) { TM_CUDA(c, cudaMalloc(&c->pstar_dev, nb)); TM_CUDA(c, cudaMemsetAsync(c->pstar_dev, 0, nb, c->stream)); }
    TM_CUDA(c, cudaMalloc(&c->pstar_band_dev, nb));
    TM_CUDA(c, cudaMemsetAsync(c->pstar_band_dev, 0, nb, c->stream));
    return TM_OK;
}

void *tm_slab_comm_stream(tm_ctx *c)
{
    if (!c || c->g.wrap) return nullptr;
    cudaSetDevice(c->device);
    if (slab_prepare(c) != TM_OK) return nullptr;
    return (void *)c->band_stream;
}

// the two state buffers (whole cudaMalloc allocations: what cudaIpcGetMemHandle wants), the index of the current one, and the
// layout a peer needs to address this slab's rows: padded row length and rows per field plane
int tm_state_buffers(tm_ctx *c, void **buf0, void **buf1, uint64_t *bytes, int *cur, int *row_doubles, int *rows_alloc)
{
    if (!c) return TM_E_ARG;
    if (buf0) *buf0 = c->buf[0];
    if (buf1) *buf1 = c->buf[1];
    if (bytes) *bytes = sizeof(double) * (uint64_t)c->g.n_trials * (uint64_t)c->g.trial_stride;
    if (cur) *cur = c->cur;
    if (row_doubles) *row_doubles = c->g.Lp;
    if (rows_alloc) *rows_alloc = c->g.rows_alloc;
    return TM_OK;
}

// SMs' worth of resident-CTA slots the inner RK4 launch leaves free (room for the band kernels and the transport's own kernels)
int tm_slab_reserve_sms(tm_ctx *c, int n_sm)
{
    if (!c || n_sm < 0 || n_sm >= c->n_sm) return TM_E_ARG;
    c->slab_reserve_sm = n_sm;
    return TM_OK;
}

// Edge rows to send and ghost rows to fill, of the state being made by the open step (next != 0; between tm_slab_step_begin
// and tm_slab_step_end) or of the current state (next == 0).  side 0 = top: send rows 0..3, ghosts above; side 1 = bottom:
// send the last four rows, ghosts below.  A (trial, field) plane's four rows are contiguous: *bytes each.
int tm_slab_halo_ptrs(tm_ctx *c, int next, int side, int trial, int field, void **send_ptr, void **recv_ptr, uint64_t *bytes)
{
    if (!c || !send_ptr || !recv_ptr || !bytes || side < 0 || side > 1) return TM_E_ARG;
    const Geom &g = c->g;
    if (g.wrap) TM_FAIL(c, TM_E_STATE, "tm_slab_halo_ptrs: not a slab context");
    if (trial < 0 || trial >= g.n_trials || field < 0 || field >= g.n_fields) TM_FAIL(c, TM_E_ARG, "tm_slab_halo_ptrs: no such plane");
    double *base = c->buf[next ? c->cur ^ 1 : c->cur] + (long long)trial * g.trial_stride + (long long)field * g.field_stride;
    const long long rs = side == 0 ? kHalo : g.nrows, rr = side == 0 ? 0 : g.nrows + kHalo;   // physical rows
    *send_ptr = base + rs * g.Lp; *recv_ptr = base + rr * g.Lp;
    *bytes = sizeof(double) * (uint64_t)kHalo * g.Lp;
    return TM_OK;
}

static int slab_args(tm_ctx *c, StepArgs *A)
{
    const tm_params &p = c->p;
    A->g = c->g; A->c = riet_consts(p);
    A->in = c->buf[c->cur]; A->out = c->buf[c->cur ^ 1];
    A->rk = philox_keys(c->seed); A->logtab = c->logtab_dev;
    A->a_trial = c->a_dev; A->seed = c->seed; A->trial0 = c->trial0; A->step = (unsigned int)c->step;
    A->counters = c->counters;
    return TM_OK;
}

int tm_slab_step_begin(tm_ctx *c)
{
    if (!c) return TM_E_ARG;
    const Geom &g = c->g;
    if (g.wrap || c->p.model != TM_MODEL_RIETKERK) TM_FAIL(c, TM_E_STATE, "tm_slab_step_begin: needs a Rietkerk row-slab context");
    if (g.nrows < 2 * kBand + 2 * kHalo) TM_FAIL(c, TM_E_STATE, "tm_slab_step_begin: a slab needs at least %d rows", 2 * kBand + 2 * kHalo);
    if (c->slab_open) TM_FAIL(c, TM_E_STATE, "tm_slab_step_begin: the previous step is still open");
    if (check_slab_step(c, "tm_slab_step_begin") != TM_OK) return TM_E_STATE;
    TM_CUDA(c, cudaSetDevice(c->device));
    if (push_trial_consts(c) != TM_OK) return TM_E_CUDA;
    if (slab_prepare(c) != TM_OK) return TM_E_CUDA;
    const int par = c->slab_parity;
    if (c->slab_fresh) {   // everything queued so far on the main stream (uploads, plain steps, a join) precedes this band launch
        TM_CUDA(c, cudaEventRecord(c->ev_inner[par ^ 1], c->stream));
        TM_CUDA(c, cudaEventRecord(c->ev_band[par ^ 1], c->band_stream));
        c->slab_fresh = false;
    }
    StepArgs A{};
    slab_args(c, &A);
    c->launches_this_call = 0;
    TM_CUDA(c, cudaStreamWaitEvent(c->band_stream, c->ev_inner[par ^ 1], 0));
    int rc = launch_pair_view(c, A, 0, kBand, c->pstar_band_dev, c->band_stream, c->n_sm * c->riet_occ, nullptr);
    if (rc == TM_OK) rc = launch_pair_view(c, A, g.nrows - kBand, g.nrows, c->pstar_band_dev, c->band_stream, c->n_sm * c->riet_occ, nullptr);
    if (rc != TM_OK) return rc;
    TM_CUDA(c, cudaEventRecord(c->ev_band[par], c->band_stream));
    c->slab_open = true;
    return TM_OK;
}

int tm_slab_step_end(tm_ctx *c)
{
    if (!c) return TM_E_ARG;
    if (!c->slab_open) TM_FAIL(c, TM_E_STATE, "tm_slab_step_end: no open step");
    const Geom &g = c->g;
    TM_CUDA(c, cudaSetDevice(c->device));
    const int par = c->slab_parity;
    StepArgs A{};
    slab_args(c, &A);
    TM_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_band[par ^ 1], 0));
    const int rc = launch_pair_view(c, A, kBand, g.nrows - kBand, c->pstar_dev, c->stream, (c->n_sm - c->slab_reserve_sm) * c->riet_occ, nullptr);
    if (rc != TM_OK) return rc;
    TM_CUDA(c, cudaEventRecord(c->ev_inner[par], c->stream));
    c->slab_open = false; c->slab_parity ^= 1;
    c->cur ^= 1; c->step++;
    c->stats_valid = false;
    c->last_launches = 6;
    return TM_OK;
}

// after the last overlapped step: the main stream waits for the band stream (band rows and ghost rows of the final state), so
// that tm_measure / tm_download / tm_step on the main stream see a complete slab
int tm_slab_join(tm_ctx *c)
{
    if (!c) return TM_E_ARG;
    if (!c->band_stream) return TM_OK;
    if (c->slab_open) TM_FAIL(c, TM_E_STATE, "tm_slab_join: a step is open");
    TM_CUDA(c, cudaSetDevice(c->device));
    TM_CUDA(c, cudaEventRecord(c->ev_join, c->band_stream));
    TM_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->slab_fresh = true;
    return TM_OK;
}

// Copy-engine transport of the slab ring for contexts of ONE process (cudaMemcpyPeerAsync over NVLink: no SM is involved, so the
// transfers cannot compete with the inner launch for CTA slots).  Pull model: `dst` fetches its ghost rows above from `up`'s last
// four rows and its ghost rows below from `down`'s first four.  next != 0: inside an open overlapped step of all three, on dst's band
// stream, behind the neighbours' band launches of this step; next == 0: the current state, on dst's main stream, behind whatever
// the neighbours' main streams hold now.  The neighbours' rows are not overwritten before dst has read them: their step after
// next needs ghost rows that dst sends only after its own next band launch, which follows this copy in stream order.
int tm_slab_pull_halos(tm_ctx *dst, tm_ctx *up, tm_ctx *down, int next)
{
    if (!dst || !up || !down) return TM_E_ARG;
    const Geom &g = dst->g;
    if (g.wrap || up->g.wrap || down->g.wrap) TM_FAIL(dst, TM_E_STATE, "tm_slab_pull_halos: not slab contexts");
    if (up->g.Lp != g.Lp || down->g.Lp != g.Lp || up->g.n_trials != g.n_trials || down->g.n_trials != g.n_trials)
        TM_FAIL(dst, TM_E_ARG, "tm_slab_pull_halos: slabs of different lattices");
    TM_CUDA(dst, cudaSetDevice(dst->device));
    if (slab_prepare(dst) != TM_OK) return TM_E_CUDA;
    cudaStream_t st = next ? dst->band_stream : dst->stream;
    tm_ctx *nb[2] = {up, down};
    for (int k = 0; k < 2; k++) {
        tm_ctx *src = nb[k];
        if (src->device != dst->device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess) cudaGetLastError();   // already enabled, or no direct path: the copy still works (staged)
        }
        if (next) {
            if (!src->slab_open) TM_FAIL(dst, TM_E_STATE, "tm_slab_pull_halos: the neighbour has no open step");
            TM_CUDA(dst, cudaStreamWaitEvent(st, src->ev_band[src->slab_parity], 0));
        } else if (src != dst) {
            TM_CUDA(dst, cudaSetDevice(src->device));
            if (slab_prepare(src) != TM_OK) return TM_E_CUDA;
            TM_CUDA(dst, cudaEventRecord(src->ev_main, src->stream));
            TM_CUDA(dst, cudaSetDevice(dst->device));
            TM_CUDA(dst, cudaStreamWaitEvent(st, src->ev_main, 0));
        }
        const Geom &sg = src->g;
        const long long rs = k == 0 ? sg.nrows : kHalo;            // up: its last four rows; down: its first four (physical rows)
        const long long rr = k == 0 ? 0 : g.nrows + kHalo;         // my ghost rows above / below
        const double *sb = src->buf[next ? src->cur ^ 1 : src->cur];
        double *db = dst->buf[next ? dst->cur ^ 1 : dst->cur];
        for (int t = 0; t < g.n_trials; t++)
            for (int f = 0; f < g.n_fields; f++)
                TM_CUDA(dst, cudaMemcpyPeerAsync(db + (long long)t * g.trial_stride + (long long)f * g.field_stride + rr * g.Lp, dst->device,
                                                 sb + (long long)t * sg.trial_stride + (long long)f * sg.field_stride + rs * sg.Lp, src->device,
                                                 sizeof(double) * kHalo * g.Lp, st));
    }
    return TM_OK;
}

void *tm_cuda_stream(tm_ctx *c) { return c ? (void *)c->stream : nullptr; }

int tm_sync(tm_ctx *c)
{
    if (!c) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_measure(tm_ctx *c, tm_stats *out)
{
    if (!c || !out) return TM_E_ARG;
    if (!c->have_state) TM_FAIL(c, TM_E_STATE, "tm_measure: no state uploaded");
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    int n_records = c->n_records;
    if (!c->stats_valid) {
        long long bpt = g.nrows;
        const long long cap = (4LL * c->n_sm + g.n_trials - 1) / g.n_trials;
        if (bpt > cap) bpt = cap;
        if (bpt < 1) bpt = 1;
        TM_CUDA(c, cudaMemsetAsync(c->partials, 0xFF, sizeof(PartialStats) * c->n_records, c->stream));
        StatsArgs A{g, c->buf[c->cur], c->partials, (int)bpt};
        stats_kernel<<<(int)bpt * g.n_trials, 256, sizeof(double) * 8 * 256, c->stream>>>(A);
        TM_CUDA(c, cudaGetLastError());
        c->stats_valid = true;
    }
    stats_finalize_kernel<<<g.n_trials, 128, 128 * 8 * sizeof(double), c->stream>>>(c->partials, n_records, g.n_trials,
                                                                            (double)((long long)g.nrows * g.L), c->stats_dev);
    TM_CUDA(c, cudaGetLastError());
    std::vector<TmStatsDev> h(g.n_trials);
    unsigned long long hc[4] = {0, 0, 0, 0};
    TM_CUDA(c, cudaMemcpyAsync(h.data(), c->stats_dev, sizeof(TmStatsDev) * g.n_trials, cudaMemcpyDeviceToHost, c->stream));
    TM_CUDA(c, cudaMemcpyAsync(hc, c->counters, sizeof hc, cudaMemcpyDeviceToHost, c->stream));
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    // a lost bulk copy (bounded mbarrier wait) leaves garbage rows: every measurement carries the check, so no runner
    // can log statistics of an invalid state
    if (hc[1]) TM_FAIL(c, TM_E_CUDA, "TMA pipeline timed out in %llu CTA(s): the state is invalid", hc[1]);
    for (int t = 0; t < g.n_trials; t++) {
        for (int f = 0; f < 3; f++) { out[t].mean[f] = h[t].mean[f]; out[t].active[f] = h[t].active[f]; }
        out[t].nonfinite = h[t].nonfinite; out[t].negative = h[t].negative;
    }
    return TM_OK;
}

int tm_measure_spread(tm_ctx *c, int c_x, int c_y, tm_spread *out)
{
    if (!c || !out) return TM_E_ARG;
    if (!c->have_state) TM_FAIL(c, TM_E_STATE, "tm_measure_spread: no state uploaded");
    TM_CUDA(c, cudaSetDevice(c->device));
    const Geom &g = c->g;
    if (!c->spread_dev) TM_CUDA(c, cudaMalloc(&c->spread_dev, sizeof(unsigned long long) * 3 * g.n_trials));
    TM_CUDA(c, cudaMemsetAsync(c->spread_dev, 0, sizeof(unsigned long long) * 3 * g.n_trials, c->stream));
    long long bpt = g.nrows;
    const long long cap = (8LL * c->n_sm + g.n_trials - 1) / g.n_trials;
    if (bpt > cap) bpt = cap;
    if (bpt < 1) bpt = 1;
    SpreadArgs A{g, c->buf[c->cur], c->spread_dev, c_x, c_y, (int)bpt};
    spread_kernel<<<(int)bpt * g.n_trials, 256, 0, c->stream>>>(A);
    TM_CUDA(c, cudaGetLastError());
    std::vector<unsigned long long> h(3 * (size_t)g.n_trials);
    TM_CUDA(c, cudaMemcpyAsync(h.data(), c->spread_dev, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int t = 0; t < g.n_trials; t++) {
        out[t].n_active = h[3 * t]; out[t].sq_dist = (double)h[3 * t + 1]; out[t].border = h[3 * t + 2] ? 1 : 0; out[t].pad_ = 0;
    }
    return TM_OK;
}

int tm_sampler_overflow(tm_ctx *c, uint64_t *count)
{
    if (!c || !count) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    unsigned long long h[4];
    TM_CUDA(c, cudaMemcpyAsync(h, c->counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    TM_CUDA(c, cudaStreamSynchronize(c->stream));
    *count = h[0];
    if (h[1]) TM_FAIL(c, TM_E_CUDA, "TMA pipeline timed out in %llu CTA(s): results are invalid", h[1]);
    return TM_OK;
}

int tm_set_kernel_timing(tm_ctx *c, int on)
{
    if (!c) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    if (on && c->kev.empty()) {
        c->kev.resize(3 * kTimedSteps);
        for (auto &e : c->kev) TM_CUDA(c, cudaEventCreate(&e));
    }
    c->kernel_timing = on != 0;
    c->kev_steps = 0;
    return TM_OK;
}

int tm_kernel_ms(tm_ctx *c, double ms[2], int *n_steps)
{
    if (!c || !ms) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    ms[0] = ms[1] = 0.0;
    for (int s = 0; s < c->kev_steps; s++) {
        float a = 0.f, b = 0.f;
        TM_CUDA(c, cudaEventSynchronize(c->kev[3 * s + 2]));
        TM_CUDA(c, cudaEventElapsedTime(&a, c->kev[3 * s], c->kev[3 * s + 1]));
        TM_CUDA(c, cudaEventElapsedTime(&b, c->kev[3 * s + 1], c->kev[3 * s + 2]));
        ms[0] += a; ms[1] += b;
    }
    if (n_steps) *n_steps = c->kev_steps;
    return TM_OK;
}

int tm_last_step_ms(tm_ctx *c, float *ms, int *n_launches)
{
    if (!c || !ms) return TM_E_ARG;
    TM_CUDA(c, cudaSetDevice(c->device));
    TM_CUDA(c, cudaEventSynchronize(c->ev1));
    TM_CUDA(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    *ms = c->last_ms;
    if (n_launches) *n_launches = c->last_launches;
    return TM_OK;
}

int tm_dp_noise_off_host(int device, int L, double a, double b, double D, double dt, double dx, double *rho, int nsteps, float *ms)
{
    auto fail = [&](int code, const std::string &msg) { g_create_error = msg; return code; };
    if (!rho || L < 1 || L > 32767 || nsteps < 0 || !(dx > 0)) return fail(TM_E_ARG, "tm_dp_noise_off_host: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(TM_E_CUDA, "tm_dp_noise_off_host: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(TM_E_ARG, "tm_dp_noise_off_host: bad device index");
    const size_t nb = sizeof(double) * (size_t)L * L;
    double *d[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaSetDevice(device);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    for (int k = 0; k < 4 && err == cudaSuccess; k++) err = cudaMalloc(&d[k], nb);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d[0], rho, nb, cudaMemcpyHostToDevice, st);
    if (err == cudaSuccess) {
        cudaEventRecord(e0, st);
        auto launch = [st](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { kernel<<<grid, block, smem, st>>>(args); };
        for (int s = 0; s < nsteps; s++) dp_det_step_enqueue(launch, L, a, b, D, dt, dx, d[0], d[1], d[2], d[3]);
        cudaEventRecord(e1, st);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(rho, d[0], nb, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err == cudaSuccess && ms) err = cudaEventElapsedTime(ms, e0, e1);
    for (int k = 0; k < 4; k++) cudaFree(d[k]);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (err != cudaSuccess) return fail(TM_E_CUDA, std::string("tm_dp_noise_off_host: ") + cudaGetErrorString(err));
    return TM_OK;
}

int tm_run_host(tm_ctx *c, double *host_soa, int nsteps, int download, tm_stats *out)
{
    int rc = tm_upload(c, host_soa);
    if (rc != TM_OK) return rc;
    rc = tm_step(c, nsteps);
    if (rc != TM_OK) return rc;
    if (out) { rc = tm_measure(c, out); if (rc != TM_OK) return rc; }
    if (download) { rc = tm_download(c, host_soa); if (rc != TM_OK) return rc; }
    return tm_sync(c);
}

} // extern "C"
