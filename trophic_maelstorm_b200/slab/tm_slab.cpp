// tm_slab.cpp -- libtm_slab.so: the row-slab ring of include/tm_slab.h.  Host code only: the kernels and contexts are
// libtm_lattice.so's (tm_slab_step_begin / _end, tm_slab_halo_ptrs), the transport is ncclSend / ncclRecv.
#include "../../include/tm_slab.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

struct tm_slab {
    tm_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int device = 0, rank = 0, n_ranks = 1, row0 = 0, nrows = 0, n_trials = 1, n_fields = 3;
    bool ring = false;              // false: one rank, the ordinary periodic context
    bool local = false;             // made by tm_slab_create_local: every slab of the ring lives in this process
    bool ghosts_valid = false;      // ghost rows of the current state are in place
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // copy-engine transport between processes (TM_SLAB_PEER_COPY on a ring made by tm_slab_create): the neighbours' state
    // buffers and flag words mapped through CUDA IPC; halo rows and the "rows of exchange #k are in place" flags travel as
    // cudaMemcpyAsync on the copy engines, a slab waits for its two flags with a stream memory operation (no SM involved)
    int transport = TM_SLAB_NCCL;
    bool ipc_ready = false;
    uint32_t *flags = nullptr;        // device: [0] set by my upper neighbour, [1] by my lower one, [2] staging word
    void *peer_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [up, down][buffer 0, 1]
    uint32_t *peer_flags[2] = {nullptr, nullptr};
    int peer_rows_alloc[2] = {0, 0};
    std::vector<void *> ipc_opened;
    uint32_t xfer_seq = 0;            // exchanges posted so far (the same on every rank: the ring is driven in lockstep)
    std::string err;
};

typedef int (*cu_memop32_fn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, uint32_t, unsigned int);
static cu_memop32_fn g_wait32 = nullptr, g_write32 = nullptr;

static thread_local std::string g_err;

static int fail(tm_slab *s, int code, const std::string &msg)
{
    if (s) s->err = msg; else g_err = msg;
    return code;
}

#define SL_NCCL(s, call)                                                                                       \
    do { ncclResult_t _r = (call); if (_r != ncclSuccess) return fail(s, TM_E_CUDA, std::string(#call ": ") + ncclGetErrorString(_r)); } while (0)
#define SL_CUDA(s, call)                                                                                       \
    do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(s, TM_E_CUDA, std::string(#call ": ") + cudaGetErrorString(_e)); } while (0)
#define SL_TM(s, call)                                                                                         \
    do { int _rc = (call); if (_rc != TM_OK) return fail(s, _rc, std::string(#call ": ") + tm_last_error((s)->ctx)); } while (0)

extern "C" {

static_assert(sizeof(ncclUniqueId) == TM_SLAB_ID_BYTES, "ncclUniqueId size");

int tm_slab_unique_id(void *id)
{
    if (!id) return TM_E_ARG;
    ncclUniqueId u;
    SL_NCCL(nullptr, ncclGetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return TM_OK;
}

void tm_slab_partition(int L, int n_ranks, int rank, int *row0, int *nrows)
{
    const int base = L / n_ranks, extra = L % n_ranks;
    if (nrows) *nrows = base + (rank < extra ? 1 : 0);
    if (row0) *row0 = rank * base + (rank < extra ? rank : extra);
}

const char *tm_slab_last_error(const tm_slab *s) { return s ? s->err.c_str() : g_err.c_str(); }

static int make_ctx(tm_slab *s, const tm_params *p, int n_trials, int device, int rank, int n_ranks)
{
    s->device = device; s->rank = rank; s->n_ranks = n_ranks; s->n_trials = n_trials; s->ring = n_ranks > 1;
    tm_slab_partition(p->L, n_ranks, rank, &s->row0, &s->nrows);
    const int rc = s->ring ? tm_create_slab(p, n_trials, device, s->row0, s->nrows, &s->ctx) : tm_create(p, n_trials, device, &s->ctx);
    if (rc != TM_OK) return fail(s, rc, std::string("tm_create_slab: ") + tm_last_error(nullptr));
    s->n_fields = tm_n_fields(s->ctx);
    SL_CUDA(s, cudaSetDevice(device));
    SL_CUDA(s, cudaEventCreate(&s->ev0));
    SL_CUDA(s, cudaEventCreate(&s->ev1));
    return TM_OK;
}

int tm_slab_create(const tm_params *p, int n_trials, int device, int rank, int n_ranks, const void *id, tm_slab **out)
{
    if (!p || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id)) return fail(nullptr, TM_E_ARG, "tm_slab_create: bad argument");
    if (p->model != TM_MODEL_RIETKERK) return fail(nullptr, TM_E_ARG, "tm_slab_create: the slab ring is built for the Rietkerk Dornic + RK4 step");
    *out = nullptr;
    tm_slab *s = new (std::nothrow) tm_slab();
    if (!s) return fail(nullptr, TM_E_ARG, "tm_slab_create: out of host memory");
    int rc = make_ctx(s, p, n_trials, device, rank, n_ranks);
    if (rc == TM_OK && s->ring) {
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        ncclResult_t r = ncclCommInitRank(&s->comm, n_ranks, u, rank);
        if (r != ncclSuccess) rc = fail(s, TM_E_CUDA, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
    }
    if (rc != TM_OK) { g_err = s->err; tm_slab_destroy(s); return rc; }
    *out = s;
    return TM_OK;
}

int tm_slab_create_local(const tm_params *p, int n_trials, int n_devices, const int *devices, int transport, tm_slab **out)
{
    if (!p || !out || n_devices < 1 || !devices) return fail(nullptr, TM_E_ARG, "tm_slab_create_local: bad argument");
    if (p->model != TM_MODEL_RIETKERK) return fail(nullptr, TM_E_ARG, "tm_slab_create_local: the slab ring is built for the Rietkerk Dornic + RK4 step");
    std::vector<tm_slab *> v(n_devices, nullptr);
    int rc = TM_OK;
    for (int i = 0; i < n_devices && rc == TM_OK; i++) {
        v[i] = new (std::nothrow) tm_slab();
        rc = v[i] ? make_ctx(v[i], p, n_trials, devices[i], i, n_devices) : TM_E_ARG;
        if (rc != TM_OK && v[i]) g_err = v[i]->err;
    }
    if (rc == TM_OK && n_devices > 1 && transport == TM_SLAB_NCCL) {
        std::vector<ncclComm_t> comms(n_devices);
        ncclResult_t r = ncclCommInitAll(comms.data(), n_devices, devices);
        if (r != ncclSuccess) rc = fail(nullptr, TM_E_CUDA, std::string("ncclCommInitAll: ") + ncclGetErrorString(r));
        else for (int i = 0; i < n_devices; i++) v[i]->comm = comms[i];
    }
    for (int i = 0; i < n_devices; i++) if (v[i]) { v[i]->transport = transport; v[i]->local = true; }
    if (rc != TM_OK) { for (tm_slab *s : v) tm_slab_destroy(s); return rc; }
    for (int i = 0; i < n_devices; i++) out[i] = v[i];
    return TM_OK;
}

int tm_slab_destroy(tm_slab *s)
{
    if (!s) return TM_OK;
    cudaSetDevice(s->device);
    if (s->ctx) tm_sync(s->ctx);
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(s->flags);
    if (s->comm) ncclCommDestroy(s->comm);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    tm_destroy(s->ctx);
    delete s;
    return TM_OK;
}

tm_ctx *tm_slab_ctx(tm_slab *s) { if (s) s->ghosts_valid = false; return s ? s->ctx : nullptr; }   // the caller may upload: ghosts are stale after that

static int ipc_setup(tm_slab *s);

int tm_slab_set_transport(tm_slab *s, int transport)
{
    if (!s || (transport != TM_SLAB_NCCL && transport != TM_SLAB_PEER_COPY)) return TM_E_ARG;
    if (transport == TM_SLAB_NCCL && s->ring && !s->comm) return fail(s, TM_E_STATE, "this ring was made without a communicator");
    s->transport = transport;
    if (transport == TM_SLAB_PEER_COPY && s->ring && !s->local) {   // rings of tm_slab_create: map the neighbours now (collective)
        SL_TM(s, tm_sync(s->ctx));
        return ipc_setup(s);
    }
    return TM_OK;
}

int tm_slab_rows(const tm_slab *s, int *row0, int *nrows)
{
    if (!s) return TM_E_ARG;
    if (row0) *row0 = s->row0;
    if (nrows) *nrows = s->nrows;
    return TM_OK;
}

int tm_slab_comm_bytes_per_step(const tm_slab *s, uint64_t *bytes)
{
    if (!s || !bytes) return TM_E_ARG;
    *bytes = 0;
    if (!s->ring) return TM_OK;
    void *a, *b; uint64_t plane = 0;
    const int rc = tm_slab_halo_ptrs(s->ctx, 0, 0, 0, 0, &a, &b, &plane);
    if (rc != TM_OK) return rc;
    *bytes = 2 * plane * (uint64_t)s->n_trials * s->n_fields;
    return TM_OK;
}

// the sends and receives of one slab for one exchange (inside the caller's ncclGroupStart / End).  Two-rank rings talk to the same
// peer in both directions: the receives are posted in the order the peer's sends are (its top rows first -> my ghosts below).
static int post_exchange(tm_slab *s, int next, cudaStream_t stream)
{
    const int up = (s->rank + s->n_ranks - 1) % s->n_ranks, down = (s->rank + 1) % s->n_ranks;
    for (int t = 0; t < s->n_trials; t++)
        for (int f = 0; f < s->n_fields; f++) {
            void *send_top, *ghost_above, *send_bot, *ghost_below; uint64_t bytes = 0;
            SL_TM(s, tm_slab_halo_ptrs(s->ctx, next, 0, t, f, &send_top, &ghost_above, &bytes));
            SL_TM(s, tm_slab_halo_ptrs(s->ctx, next, 1, t, f, &send_bot, &ghost_below, &bytes));
            const size_t n = bytes / sizeof(double);
            SL_NCCL(s, ncclSend(send_top, n, ncclDouble, up, s->comm, stream));
            SL_NCCL(s, ncclSend(send_bot, n, ncclDouble, down, s->comm, stream));
            SL_NCCL(s, ncclRecv(ghost_below, n, ncclDouble, down, s->comm, stream));
            SL_NCCL(s, ncclRecv(ghost_above, n, ncclDouble, up, s->comm, stream));
        }
    return TM_OK;
}

// ---- CUDA IPC transport -------------------------------------------------------------------------------------------------
struct IpcCard { cudaIpcMemHandle_t buf[2], flags; int rows_alloc, cur; };

static int ipc_setup(tm_slab *s)
{
    if (s->ipc_ready) return TM_OK;
    if (!s->comm) return fail(s, TM_E_STATE, "peer-copy transport between processes needs the ring's communicator (tm_slab_create)");
    SL_CUDA(s, cudaSetDevice(s->device));
    if (!g_wait32) {
        cudaDriverEntryPointQueryResult q;
        void *f = nullptr;
        SL_CUDA(s, cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q));
        g_wait32 = (cu_memop32_fn)f;
        SL_CUDA(s, cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q));
        g_write32 = (cu_memop32_fn)f;
        if (!g_wait32 || !g_write32) return fail(s, TM_E_CUDA, "stream memory operations are not available");
    }
    // a whole 2 MiB block of its own: smaller allocations are carved out of shared blocks, and it is the block that an IPC handle
    // exports (CUDA programming guide, interprocess communication)
    SL_CUDA(s, cudaMalloc(&s->flags, 2u << 20));
    SL_CUDA(s, cudaMemset(s->flags, 0, 2u << 20));
    IpcCard mine;
    memset(&mine, 0, sizeof mine);
    void *b0, *b1; uint64_t bytes; int row_doubles;
    SL_TM(s, tm_state_buffers(s->ctx, &b0, &b1, &bytes, &mine.cur, &row_doubles, &mine.rows_alloc));
    SL_CUDA(s, cudaIpcGetMemHandle(&mine.buf[0], b0));
    SL_CUDA(s, cudaIpcGetMemHandle(&mine.buf[1], b1));
    SL_CUDA(s, cudaIpcGetMemHandle(&mine.flags, s->flags));
    // everybody's card to everybody (a few hundred bytes, once): ncclAllGather
    std::vector<IpcCard> all(s->n_ranks);
    char *dev = nullptr;
    SL_CUDA(s, cudaMalloc(&dev, sizeof(IpcCard) * (s->n_ranks + 1)));
    SL_CUDA(s, cudaMemcpy(dev + sizeof(IpcCard) * s->n_ranks, &mine, sizeof mine, cudaMemcpyHostToDevice));
    cudaStream_t st = (cudaStream_t)tm_cuda_stream(s->ctx);
    SL_NCCL(s, ncclAllGather(dev + sizeof(IpcCard) * s->n_ranks, dev, sizeof(IpcCard), ncclChar, s->comm, st));
    SL_CUDA(s, cudaStreamSynchronize(st));
    SL_CUDA(s, cudaMemcpy(all.data(), dev, sizeof(IpcCard) * s->n_ranks, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    const int nb[2] = {(s->rank + s->n_ranks - 1) % s->n_ranks, (s->rank + 1) % s->n_ranks};
    for (int k = 0; k < 2; k++) {
        if (k == 1 && nb[1] == nb[0]) {                     // two ranks: the same peer above and below
            for (int b = 0; b < 2; b++) s->peer_buf[1][b] = s->peer_buf[0][b];
            s->peer_flags[1] = s->peer_flags[0]; s->peer_rows_alloc[1] = s->peer_rows_alloc[0];
            continue;
        }
        const IpcCard &c = all[nb[k]];
        if (c.cur != mine.cur) return fail(s, TM_E_STATE, "the ranks of the ring are not in lockstep (different current buffer)");
        for (int b = 0; b < 2; b++) {
            SL_CUDA(s, cudaIpcOpenMemHandle(&s->peer_buf[k][b], c.buf[b], cudaIpcMemLazyEnablePeerAccess));
            s->ipc_opened.push_back(s->peer_buf[k][b]);
        }
        void *pf = nullptr;
        SL_CUDA(s, cudaIpcOpenMemHandle(&pf, c.flags, cudaIpcMemLazyEnablePeerAccess));
        s->ipc_opened.push_back(pf);
        s->peer_flags[k] = (uint32_t *)pf;
        s->peer_rows_alloc[k] = c.rows_alloc;
    }
    // Nobody may queue its flag waits before EVERY rank has finished mapping: opening a handle (and the peer access it enables)
    // can wait for the owner's GPU to drain, and a GPU parked in a flag wait drains only when its neighbours push -- which a
    // neighbour still stuck in its mapping calls cannot do.  (An 8-rank run hung before this barrier was here.)
    {
        int *tok = nullptr;
        SL_CUDA(s, cudaMalloc(&tok, sizeof(int)));
        SL_CUDA(s, cudaMemset(tok, 0, sizeof(int)));
        SL_NCCL(s, ncclAllReduce(tok, tok, 1, ncclInt, ncclSum, s->comm, st));
        SL_CUDA(s, cudaStreamSynchronize(st));
        cudaFree(tok);
    }
    s->ipc_ready = true;
    return TM_OK;
}

// one exchange over the copy engines: my edge rows into the neighbours' ghost rows, then the sequence number into their flag words,
// then this stream waits until both of my own flag words carry it.  The neighbours' ghost rows of a buffer are not overwritten
// before they were read: a push follows my band launch, which waited for the flags the neighbours raise only after THEIR band launch.
static int post_exchange_ipc(tm_slab *s, int next, cudaStream_t stream)
{
    int rc = ipc_setup(s);
    if (rc != TM_OK) return rc;
    void *b0, *b1; uint64_t bytes; int cur, Lp, rows_alloc;
    SL_TM(s, tm_state_buffers(s->ctx, &b0, &b1, &bytes, &cur, &Lp, &rows_alloc));
    const int which = next ? cur ^ 1 : cur;
    const int halo = tm_halo_rows();
    const uint32_t seq = ++s->xfer_seq;
    for (int t = 0; t < s->n_trials; t++)
        for (int f = 0; f < s->n_fields; f++) {
            void *send_top, *ghost_above, *send_bot, *ghost_below; uint64_t plane = 0;
            SL_TM(s, tm_slab_halo_ptrs(s->ctx, next, 0, t, f, &send_top, &ghost_above, &plane));
            SL_TM(s, tm_slab_halo_ptrs(s->ctx, next, 1, t, f, &send_bot, &ghost_below, &plane));
            // up neighbour: its ghost rows BELOW its last row = physical rows [rows_alloc - halo, rows_alloc); down: its ghost rows above = [0, halo)
            const long long fs_up = (long long)s->peer_rows_alloc[0] * Lp, fs_dn = (long long)s->peer_rows_alloc[1] * Lp;
            double *up_dst = (double *)s->peer_buf[0][which] + ((long long)t * s->n_fields + f) * fs_up + (long long)(s->peer_rows_alloc[0] - halo) * Lp;
            double *dn_dst = (double *)s->peer_buf[1][which] + ((long long)t * s->n_fields + f) * fs_dn;
            SL_CUDA(s, cudaMemcpyAsync(up_dst, send_top, plane, cudaMemcpyDefault, stream));
            SL_CUDA(s, cudaMemcpyAsync(dn_dst, send_bot, plane, cudaMemcpyDefault, stream));
        }
    if (g_write32(stream, (unsigned long long)(uintptr_t)(s->flags + 2), seq, 0) != 0) return fail(s, TM_E_CUDA, "cuStreamWriteValue32 failed");
    SL_CUDA(s, cudaMemcpyAsync(s->peer_flags[0] + 1, s->flags + 2, 4, cudaMemcpyDefault, stream));   // I am my upper neighbour's lower one
    SL_CUDA(s, cudaMemcpyAsync(s->peer_flags[1] + 0, s->flags + 2, 4, cudaMemcpyDefault, stream));
    for (int k = 0; k < 2; k++)
        if (g_wait32(stream, (unsigned long long)(uintptr_t)(s->flags + k), seq, 0x1 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
            return fail(s, TM_E_CUDA, "cuStreamWaitValue32 failed");
    return TM_OK;
}

static int exchange_all(tm_slab **v, int n, int next, bool band_stream)
{
    if (v[0]->transport == TM_SLAB_PEER_COPY && !v[0]->local && v[0]->n_ranks > 1) {   // one slab per process: CUDA IPC
        cudaSetDevice(v[0]->device);
        cudaStream_t st = (cudaStream_t)(band_stream ? tm_slab_comm_stream(v[0]->ctx) : tm_cuda_stream(v[0]->ctx));
        return post_exchange_ipc(v[0], next, st);
    }
    if (v[0]->transport == TM_SLAB_PEER_COPY) {           // peer copies: all slabs of the ring are local
        if (n != v[0]->n_ranks) return fail(v[0], TM_E_STATE, "peer-copy transport needs the whole ring in this process");
        for (int i = 0; i < n; i++) SL_TM(v[i], tm_slab_pull_halos(v[i]->ctx, v[(i + n - 1) % n]->ctx, v[(i + 1) % n]->ctx, next));
        return TM_OK;
    }
    SL_NCCL(v[0], ncclGroupStart());
    int rc = TM_OK;
    for (int i = 0; i < n && rc == TM_OK; i++) {
        cudaSetDevice(v[i]->device);
        cudaStream_t st = (cudaStream_t)(band_stream ? tm_slab_comm_stream(v[i]->ctx) : tm_cuda_stream(v[i]->ctx));
        rc = post_exchange(v[i], next, st);
    }
    ncclResult_t r = ncclGroupEnd();
    if (rc == TM_OK && r != ncclSuccess) rc = fail(v[0], TM_E_CUDA, std::string("ncclGroupEnd: ") + ncclGetErrorString(r));
    return rc;
}

int tm_slab_run_local(tm_slab **v, int n, int nsteps, int mode)
{
    if (!v || n < 1 || nsteps < 0 || mode < 0 || mode > 2) return TM_E_ARG;
    for (int i = 0; i < n; i++) if (!v[i]) return TM_E_ARG;
    for (int i = 0; i < n; i++) {
        SL_CUDA(v[i], cudaSetDevice(v[i]->device));
        SL_CUDA(v[i], cudaEventRecord(v[i]->ev0, (cudaStream_t)tm_cuda_stream(v[i]->ctx)));
    }
    if (!v[0]->ring) {                                    // one rank: the ordinary periodic lattice
        SL_TM(v[0], tm_step(v[0]->ctx, nsteps));
    } else {
        if (mode != TM_SLAB_NOCOMM && !v[0]->ghosts_valid && nsteps > 0) {   // ghost rows of the starting state, on the main streams
            const int rc = exchange_all(v, n, 0, false);
            if (rc != TM_OK) return rc;
            for (int i = 0; i < n; i++) v[i]->ghosts_valid = true;
        }
        // TM_SLAB_TRACE=1: device timestamps of the last steps of slab 0 (band done, transfers done, inner begin / end), to stderr
        constexpr int kTraceSteps = 6;
        const bool trace = getenv("TM_SLAB_TRACE") && mode != TM_SLAB_SERIAL && nsteps >= kTraceSteps;
        std::vector<cudaEvent_t> tev;
        if (trace) { cudaSetDevice(v[0]->device); tev.resize(5 * kTraceSteps); for (auto &x : tev) cudaEventCreate(&x); }
        for (int k = 0; k < nsteps; k++) {
            if (mode == TM_SLAB_SERIAL) {
                for (int i = 0; i < n; i++) SL_TM(v[i], tm_step(v[i]->ctx, 1));
                const int rc = exchange_all(v, n, 0, false);
                if (rc != TM_OK) return rc;
            } else {
                const bool tr = trace && k >= nsteps - kTraceSteps;
                cudaStream_t sb = (cudaStream_t)tm_slab_comm_stream(v[0]->ctx), sm = (cudaStream_t)tm_cuda_stream(v[0]->ctx);
                cudaEvent_t *e = tr ? &tev[5 * (k - (nsteps - kTraceSteps))] : nullptr;
                for (int i = 0; i < n; i++) SL_TM(v[i], tm_slab_step_begin(v[i]->ctx));
                if (tr) cudaEventRecord(e[0], sb);                       // band launches done
                if (mode == TM_SLAB_OVERLAP) {
                    const int rc = exchange_all(v, n, 1, true);
                    if (rc != TM_OK) return rc;
                }
                if (tr) { cudaEventRecord(e[1], sb); cudaEventRecord(e[2], sm); }   // transfers done; main stream before the inner launches
                for (int i = 0; i < n; i++) SL_TM(v[i], tm_slab_step_end(v[i]->ctx));
                if (tr) cudaEventRecord(e[3], sm);                       // inner launches done
            }
        }
        if (mode != TM_SLAB_SERIAL) for (int i = 0; i < n; i++) SL_TM(v[i], tm_slab_join(v[i]->ctx));
        if (mode == TM_SLAB_NOCOMM) for (int i = 0; i < n; i++) v[i]->ghosts_valid = false;
        if (trace) {
            cudaSetDevice(v[0]->device);
            cudaStreamSynchronize((cudaStream_t)tm_cuda_stream(v[0]->ctx));
            cudaStreamSynchronize((cudaStream_t)tm_slab_comm_stream(v[0]->ctx));
            for (int k = 0; k < kTraceSteps; k++) {
                float t[4];
                for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], tev[2], tev[5 * k + j]);
                fprintf(stderr, "[tm_slab trace rank %d mode %d] step -%d: band done %.3f  transfers done %.3f  inner begin %.3f  inner end %.3f ms\n",
                        v[0]->rank, mode, kTraceSteps - k, t[0], t[1], t[2], t[3]);
            }
            for (auto &x : tev) cudaEventDestroy(x);
        }
    }
    for (int i = 0; i < n; i++) {
        SL_CUDA(v[i], cudaSetDevice(v[i]->device));
        SL_CUDA(v[i], cudaEventRecord(v[i]->ev1, (cudaStream_t)tm_cuda_stream(v[i]->ctx)));
    }
    return TM_OK;
}

int tm_slab_run(tm_slab *s, int nsteps, int mode) { return tm_slab_run_local(&s, 1, nsteps, mode); }

int tm_slab_sync(tm_slab *s)
{
    if (!s) return TM_E_ARG;
    SL_CUDA(s, cudaSetDevice(s->device));
    SL_TM(s, tm_sync(s->ctx));
    if (s->ring) if (void *st = tm_slab_comm_stream(s->ctx)) SL_CUDA(s, cudaStreamSynchronize((cudaStream_t)st));
    return TM_OK;
}

int tm_slab_last_ms(tm_slab *s, float *ms)
{
    if (!s || !ms) return TM_E_ARG;
    SL_CUDA(s, cudaSetDevice(s->device));
    SL_CUDA(s, cudaEventSynchronize(s->ev1));
    SL_CUDA(s, cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return TM_OK;
}

} // extern "C"
