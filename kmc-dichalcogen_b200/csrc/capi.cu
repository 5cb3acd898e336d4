// capi.cu -- the C ABI (include/kmc_b200.h) over the CUDA kernels.  No torch, no C++ types in the
// signatures; errors become status codes + a thread-local message.  There is NO CPU fallback: without
// a CUDA device every entry point that computes fails with KMC_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>
#include <algorithm>
#include <vector>
#include <chrono>

#include "../../include/kmc_b200.h"
#include "kmc_common.cuh"
#include "noinc_kernel.cuh"
#include "noinc_planes_kernel.cuh"
#include "replica_kernel.cuh"
#include "replica_bp_kernel.cuh"
#include "replica_hw_kernel.cuh"
#include "replica_wide_kernel.cuh"
#include "replica_wide2_kernel.cuh"
#include "sweep_kernel.cuh"
#include "stategen_kernel.cuh"
#include "zobrist_kernel.cuh"

using namespace kmc;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                     __FILE__, __LINE__);                                                      \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver           \
                            ? KMC_ERR_NO_DEVICE : KMC_ERR_CUDA, b_);                           \
        }                                                                                      \
    } while (0)

struct kmc_engine {
    int device = 0, d0 = 0, d1 = 0, n = 0, R = 0;
    double rate[NCA];
    int order_pos[NCA];
    bool mm = false;                         // a MoveMonovacancy class (13..16) is enabled: general kernels only
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_status = nullptr;
    uint8_t *d_slots = nullptr;              // [R][17][n], allocated on first sweep
    uint8_t *d_slots_alt = nullptr; bool slots_double = false;   // measurement: alternate output buffers
    uint32_t *d_rowcnt = nullptr;            // [R][d0][16] (general MoveMonovacancy sweep: [R][d0][32])
    uint32_t *d_rowmm = nullptr;             // [R][d0][4] MoveMonovacancy row counts of the aligned two-kernel sweep
    bool sweep_split = false;                // the last sweep left its row counts in d_rowcnt (stride 16) + d_rowmm
    unsigned long long *d_counts = nullptr;  // [R][13]
    unsigned long long *d_acc = nullptr;     // [R][32] running totals of the aligned sweep (zero between launches)
    unsigned int *d_ticket = nullptr;        // [R]
    unsigned long long *d_steps = nullptr, *d_hist = nullptr, *d_ties = nullptr, *d_hops = nullptr;
    uint32_t *d_flags = nullptr;
    uint32_t *d_anyflag = nullptr;           // OR of every replica's flags: the handle's asynchronous status word
    void *d_events = nullptr; size_t events_cap = 0;
    double *d_draws = nullptr; size_t draws_cap = 0;
    int *d_result = nullptr;
    double *d_clock = nullptr; int clock_mode = 0;           // optional KMC clock per replica: 0 off, 1 mean residence, 2 stochastic
    uint32_t *d_tracer = nullptr; bool tracer_on = false;    // optional divacancy tracers [R][n] (int16 da | int16 db << 16)
    uint32_t *d_hier = nullptr; bool hier_valid = false;   // row hierarchy of HBM-resident lattices
    // wide bit-sliced kernel (rows of 64*W sites): planes + row table live in HBM between launches.  While
    // planes_valid they describe the current lattice; bytes_stale means d_status lags behind them.
    unsigned long long *d_planes = nullptr; uint16_t *d_whier = nullptr, *d_whier_chk = nullptr;
    bool planes_valid = false, bytes_stale = false;
    bool whier_valid = false;                // d_whier matches the planes (the no-incremental plane path does not maintain it)
    uint32_t *d_prowc = nullptr, *d_pblk = nullptr;   // no-incremental path on planes: class-major row counts, CTA sums
    unsigned int *d_pepoch = nullptr;                 // persistent form: events completed in this call
    unsigned int *d_pticket = nullptr;                // and the per-replica ticket that elects the selecting CTA
    void *d_flush = nullptr; size_t flush_bytes = 0;   // L2-flush scratch (measurement helper)
    int warps_per_block = 0;
    int variant = 0;                         // 0 auto, 1 byte lattice, 2 bit planes, 3 byte lattice kept in HBM
    int sweep_variant = 0;                   // aligned shapes: 0 auto (rolling window for 4096-site rows, else
                                             // nibble/PRMT), 1 nibble/PRMT, 2 byte-parallel
    long long launches = 0;
    bool swept = false;
    int max_smem_optin = 0;
};

extern "C" const char *kmc_last_error(void) { return g_err.c_str(); }

extern "C" int kmc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int bind(kmc_engine *h) {
    if (!h) return fail(KMC_ERR_ARG, "null engine handle");
    CU(cudaSetDevice(h->device));
    return KMC_OK;
}

static int create_device_side(kmc_engine *h);

extern "C" int kmc_create(kmc_engine **out, int device, int dim0, int dim1, int n_replicas,
                          const double *class_rate, const int *class_order, int n_order)
{
    if (!out) return fail(KMC_ERR_ARG, "kmc_create: out is null");
    *out = nullptr;
    if (dim0 < 5 || dim1 < 5 || dim0 == 6 || dim1 == 6)
        return fail(KMC_ERR_ARG, "kmc_create: lattice dimensions must be >= 5 and != 6 "
                                 "(smaller lattices alias the trefoil geometry under periodic boundaries)");
    if ((long long)dim0 * dim1 > 0x7FFFFFFFLL / 32) return fail(KMC_ERR_ARG, "kmc_create: lattice too large");
    if (n_replicas < 1) return fail(KMC_ERR_ARG, "kmc_create: n_replicas must be >= 1");
    if (!class_rate || (n_order > 0 && !class_order) || n_order < 0 || n_order > NCA)
        return fail(KMC_ERR_ARG, "kmc_create: bad class_rate / class_order");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(KMC_ERR_NO_DEVICE, "kmc_create: no CUDA device available (this engine has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(KMC_ERR_ARG, "kmc_create: bad device index");
    kmc_engine *h = new kmc_engine();
    h->device = device; h->d0 = dim0; h->d1 = dim1; h->n = dim0 * dim1; h->R = n_replicas;
    for (int c = 0; c < NCA; c++) { h->rate[c] = 0.0; h->order_pos[c] = -1; }
    for (int i = 0; i < n_order; i++) {
        int c = class_order[i];
        if (c < 0 || c >= NCA || h->order_pos[c] >= 0) { delete h; return fail(KMC_ERR_ARG, "kmc_create: class_order must list distinct class ids 0..16"); }
        if (!(class_rate[c] >= 0.0)) { delete h; return fail(KMC_ERR_ARG, "kmc_create: negative or NaN rate"); }   // kmc.py:37-39
        h->order_pos[c] = i; h->rate[c] = class_rate[c];
        if (c >= C_MOVE_MONO) h->mm = true;
    }
    const int rc = create_device_side(h);
    if (rc) { const std::string msg = g_err; kmc_destroy(h); return fail(rc, msg); }   // frees whatever was allocated
    *out = h;
    return KMC_OK;
}

// device half of kmc_create; on failure the caller destroys the half-built engine
static int create_device_side(kmc_engine *h)
{
    const int device = h->device, n_replicas = h->R;
    CU(cudaSetDevice(device));
    CU(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    const size_t R = (size_t)n_replicas;
    CU(cudaMalloc(&h->d_status, R * h->n));
    CU(cudaMalloc(&h->d_counts, R * NCA * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_steps, R * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_hist, R * NCA * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_hops, R * 6 * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_ties, R * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->d_flags, R * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_result, sizeof(int)));
    CU(cudaMemsetAsync(h->d_status, 0, R * h->n, h->stream));
    CU(cudaMemsetAsync(h->d_counts, 0, R * NCA * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_steps, 0, R * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_hist, 0, R * NCA * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_hops, 0, R * 6 * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_ties, 0, R * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_flags, 0, R * sizeof(uint32_t), h->stream));
    CU(cudaMalloc(&h->d_anyflag, sizeof(uint32_t)));
    CU(cudaMemsetAsync(h->d_anyflag, 0, sizeof(uint32_t), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_destroy(kmc_engine *h)
{
    if (!h) return KMC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_status); cudaFree(h->d_slots); cudaFree(h->d_slots_alt); cudaFree(h->d_rowcnt); cudaFree(h->d_rowmm); cudaFree(h->d_counts);
    cudaFree(h->d_acc); cudaFree(h->d_ticket); cudaFree(h->d_flush); cudaFree(h->d_hier); cudaFree(h->d_clock); cudaFree(h->d_tracer);
    cudaFree(h->d_planes); cudaFree(h->d_whier); cudaFree(h->d_whier_chk); cudaFree(h->d_prowc); cudaFree(h->d_pblk); cudaFree(h->d_pticket); cudaFree(h->d_pepoch);
    cudaFree(h->d_steps); cudaFree(h->d_hist); cudaFree(h->d_hops); cudaFree(h->d_ties); cudaFree(h->d_flags); cudaFree(h->d_anyflag);
    cudaFree(h->d_events); cudaFree(h->d_draws); cudaFree(h->d_result);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return KMC_OK;
}

// ---- lattice transfer ---------------------------------------------------------------------------
// State invariants (reference demo1/state.py:102-149): known status codes, complete trefoils.
// Checked on the device after the copy so that a 64 MiB batch upload is not CPU-bound.
__global__ void kmc_check_state_kernel(const uint8_t *status, long long total, int d0, int d1, int *result)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int n = d0 * d1;
    const long long rep = g / n;
    const int i = (int)(g - rep * n);
    const uint8_t *st = status + rep * n;
    const int v = st[i];
    if (v > 9) { atomicOr(result, 1); return; }
    if (v >= 4) {
        const int o = (v - 4) / 3, role = (v - 4) % 3;
        // member offsets: Up {0,0},{2,0},{0,2}; Down {0,0},{2,0},{2,-2}
        const int ma[3] = {0, 2, o ? 2 : 0}, mb[3] = {0, 0, o ? -2 : 2};
        const int a = i / d1 - ma[role], b = i % d1 - mb[role];
        for (int q = 0; q < 3; q++) {
            const int aa = ((a + ma[q]) % d0 + d0) % d0, bb = ((b + mb[q]) % d1 + d1) % d1;
            if (st[aa * d1 + bb] != 4 + 3 * o + q) atomicOr(result, 2);
        }
    }
}

// bring the status bytes up to date after launches of the wide bit-sliced kernel
static int ensure_bytes(kmc_engine *h)
{
    if (!h->bytes_stale) return KMC_OK;
    const long long cells = (long long)h->R * h->d0 * wide_anchor_words(h->d1);
    kmc_wide_unpack_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, h->stream>>>(h->d_planes, h->d_status, h->R, h->d0, h->d1);
    CU(cudaGetLastError());
    h->launches += 1;
    h->bytes_stale = false;
    return KMC_OK;
}

// The zero-rate / validation flags describe the lattice they were raised on: whenever the lattice is
// replaced or edited they are cleared (KmcSim.initialize starts afresh, sim.py:46-68).
static int lattice_replaced(kmc_engine *h)
{
    h->swept = false; h->hier_valid = false; h->planes_valid = false; h->bytes_stale = false;
    CU(cudaMemsetAsync(h->d_flags, 0, (size_t)h->R * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->d_anyflag, 0, sizeof(uint32_t), h->stream));
    // every divacancy of a new lattice is a new tracer
    if (h->d_tracer) CU(cudaMemsetAsync(h->d_tracer, 0, (size_t)h->R * h->n * sizeof(uint32_t), h->stream));
    return KMC_OK;
}

static int check_state_device(kmc_engine *h)
{
    CU(cudaMemsetAsync(h->d_result, 0, sizeof(int), h->stream));
    const long long total = (long long)h->R * h->n;
    kmc_check_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(h->d_status, total, h->d0, h->d1, h->d_result);
    CU(cudaGetLastError());
    h->launches += 1;
    int res = 0;
    CU(cudaMemcpyAsync(&res, h->d_result, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (res & 1) return fail(KMC_ERR_STATE, "invalid status byte (> 9) in the lattice");
    if (res & 2) return fail(KMC_ERR_STATE, "incomplete trefoil in the lattice");
    return KMC_OK;
}

extern "C" int kmc_upload_state(kmc_engine *h, const uint8_t *host_status)
{
    int rc = bind(h); if (rc) return rc;
    if (!host_status) return fail(KMC_ERR_ARG, "kmc_upload_state: null buffer");
    CU(cudaMemcpyAsync(h->d_status, host_status, (size_t)h->R * h->n, cudaMemcpyHostToDevice, h->stream));
    rc = lattice_replaced(h); if (rc) return rc;
    return check_state_device(h);
}

extern "C" int kmc_download_state(kmc_engine *h, uint8_t *host_status)
{
    int rc = bind(h); if (rc) return rc;
    if (!host_status) return fail(KMC_ERR_ARG, "kmc_download_state: null buffer");
    rc = ensure_bytes(h); if (rc) return rc;
    CU(cudaMemcpyAsync(host_status, h->d_status, (size_t)h->R * h->n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_set_state_device(kmc_engine *h, const void *dev_status)
{
    int rc = bind(h); if (rc) return rc;
    if (!dev_status) return fail(KMC_ERR_ARG, "kmc_set_state_device: null pointer");
    CU(cudaMemcpyAsync(h->d_status, dev_status, (size_t)h->R * h->n, cudaMemcpyDeviceToDevice, h->stream));
    rc = lattice_replaced(h); if (rc) return rc;
    return check_state_device(h);
}

extern "C" int kmc_generate_state(kmc_engine *h, int mode, int n_keys, const int *codes, const int64_t *bounds,
                                  const double *cumul, uint64_t seed, uint32_t replica0)
{
    int rc = bind(h); if (rc) return rc;
    if (mode < 0 || mode > 1) return fail(KMC_ERR_ARG, "kmc_generate_state: mode must be 0 (exact) or 1 (approx)");
    if (n_keys < 1 || n_keys > 3 || !codes) return fail(KMC_ERR_ARG, "kmc_generate_state: 1..3 keys");
    GenParams p;
    p.mode = mode; p.n = h->n; p.n_replicas = h->R; p.n_keys = n_keys;
    p.seed = seed; p.replica0 = replica0; p.status = h->d_status; p.perm = nullptr;
    for (int k = 0; k < 3; k++) { p.code[k] = 0; p.cumul[k] = 0.0; }
    for (int k = 0; k < 4; k++) p.bound[k] = 0;
    for (int k = 0; k < n_keys; k++) {
        if (codes[k] != 0 && codes[k] != 1 && codes[k] != 3) return fail(KMC_ERR_ARG, "kmc_generate_state: codes are 0, 1 or 3");
        p.code[k] = codes[k];
    }
    if (mode == 0) {
        if (!bounds) return fail(KMC_ERR_ARG, "kmc_generate_state: exact mode needs bounds");
        for (int k = 0; k <= n_keys; k++) {
            p.bound[k] = bounds[k];
            if (bounds[k] < 0 || bounds[k] > h->n || (k && bounds[k] < bounds[k - 1]))
                return fail(KMC_ERR_ARG, "kmc_generate_state: bounds must be ascending within [0, sites]");
        }
    } else {
        if (!cumul) return fail(KMC_ERR_ARG, "kmc_generate_state: approx mode needs cumulative weights");
        for (int k = 0; k < n_keys; k++) {
            p.cumul[k] = cumul[k];
            if (!(cumul[k] > 0.0) || (k && cumul[k] < cumul[k - 1]))
                return fail(KMC_ERR_ARG, "kmc_generate_state: cumulative weights must be positive and ascending");
        }
    }
    CU(cudaMemsetAsync(h->d_status, 0, (size_t)h->R * h->n, h->stream));
    uint32_t *perm = nullptr;
    if (mode == 0) {
        CU(cudaMalloc(&perm, (size_t)h->R * h->n * sizeof(uint32_t)));
        p.perm = perm;
    }
    if (mode == 1) {
        const long long sites = (long long)h->R * h->n;
        kmc_generate_approx_sites_kernel<<<(unsigned)((sites + 255) / 256), 256, 0, h->stream>>>(p);
        h->launches += 1;
    }
    kmc_generate_state_kernel<<<(unsigned)((h->R + 63) / 64), 64, 0, h->stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (perm) cudaFree(perm);
    if (e != cudaSuccess) return fail(KMC_ERR_CUDA, cudaGetErrorString(e));
    h->launches += 1;
    return lattice_replaced(h);
}

extern "C" int kmc_get_state_device(kmc_engine *h, void *dev_status)
{
    int rc = bind(h); if (rc) return rc;
    if (!dev_status) return fail(KMC_ERR_ARG, "kmc_get_state_device: null pointer");
    rc = ensure_bytes(h); if (rc) return rc;
    CU(cudaMemcpyAsync(dev_status, h->d_status, (size_t)h->R * h->n, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

// ---- K1/K2 sweep ---------------------------------------------------------------------------------
static int sweep_launch(kmc_engine *h, bool with_slots)
{
    const size_t R = (size_t)h->R;
    { int rc = ensure_bytes(h); if (rc) return rc; }
    const int rcs = h->mm ? RCG_STRIDE_MM : RCG_STRIDE, nsl = h->mm ? NSA : NS;
    if (!h->d_rowcnt) CU(cudaMalloc(&h->d_rowcnt, R * h->d0 * rcs * sizeof(uint32_t)));
    if (with_slots && !h->d_slots) CU(cudaMalloc(&h->d_slots, R * nsl * h->n));
    // aligned fast path: 16 sites per thread, CTAs aligned to rows and replicas
    const int cpr = h->d1 >> 4;
    const int cpl_all = h->n >> 4;                     // 16-site chunks per lattice
    // (lattices of fewer than 4096 sites: a CTA of the 16-sites-per-thread kernel holds several whole replicas)
    const bool aligned_small = (h->d1 & 15) == 0 && cpr >= 1 && (256 % cpr) == 0 && (h->n & 15) == 0 && cpl_all < 256 &&
                               (256 % cpl_all) == 0 && ((long long)h->R * cpl_all) % 256 == 0;
    const bool aligned = ((h->d1 & 15) == 0 && cpr <= 256 && (256 % cpr) == 0 && (cpl_all % 256) == 0) || aligned_small;
    // MoveMonovacancy live on an aligned shape: the MoveMonovacancy builds of the nibble kernels write all 29 planes;
    // the row counts are split over two tables, d_rowcnt [..][16] (classes 0..12) and d_rowmm [..][4] (classes 13..16)
    const bool mm_split = h->mm && aligned;
    h->sweep_split = mm_split;
    if (h->mm && !mm_split) {
        // MoveMonovacancy live: the general enumeration kernel (all 9 rules, 29 slot planes, one thread per site)
        CU(cudaMemsetAsync(h->d_rowcnt, 0, R * h->d0 * rcs * sizeof(uint32_t), h->stream));
        SweepParams p;
        p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R; p.total_words = 0;
        p.status = h->d_status; p.slots = with_slots ? h->d_slots : nullptr; p.rowcnt = h->d_rowcnt;
        const long long blocks = ((long long)R * h->n + 255) / 256;
        kmc_sweep_kernel_mm<<<(unsigned)blocks, 256, 0, h->stream>>>(p);
        CU(cudaGetLastError());
        kmc_reduce_counts_mm_kernel<<<h->R, 256, 0, h->stream>>>(h->d_rowcnt, h->d0, h->R, h->d_counts);
        CU(cudaGetLastError());
        h->launches += 2;
        return KMC_OK;
    }
    if (with_slots && h->slots_double) {
        // steady-state measurement: successive sweeps write alternating buffers, so that a launch's stores
        // never hit lines the previous launch left dirty in L2 (the result the getters read is the last one)
        if (!h->d_slots_alt) CU(cudaMalloc(&h->d_slots_alt, R * (h->mm ? NSA : NS) * h->n));
        std::swap(h->d_slots, h->d_slots_alt);
    }
    // (aligned shapes: the nibble kernels do all 29 planes in one pass; the four MoveMonovacancy row counts have their
    // own table)
    if (mm_split && !h->d_rowmm) CU(cudaMalloc(&h->d_rowmm, R * h->d0 * 4 * sizeof(uint32_t)));
    if (aligned) {
        if (!h->d_acc) {
            CU(cudaMalloc(&h->d_acc, R * ROLL_ACC_STRIDE * sizeof(unsigned long long)));
            CU(cudaMalloc(&h->d_ticket, R * sizeof(unsigned int)));
            CU(cudaMemsetAsync(h->d_acc, 0, R * ROLL_ACC_STRIDE * sizeof(unsigned long long), h->stream));
            CU(cudaMemsetAsync(h->d_ticket, 0, R * sizeof(unsigned int), h->stream));
        }
        if (cpr == 256 && h->sweep_variant == 0) {
            // rows that fill a CTA: rolling window over rows; pick rows per CTA so that the grid is about one
            // full wave (148 SMs x 3 resident CTAs)
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
            // rows per CTA: 2..8, favouring ~5 (measured best trade between register reuse and parallel
            // slack) and a grid that fills its last wave (3 resident CTAs per SM)
            const double slots = (double)sms * 3.0;
            int rpc = 1;
            double best = -1.0;
            for (int c = 2; c <= 8 && c <= h->d0; c++) {
                const double ctas = (double)R * ((h->d0 + c - 1) / c);
                const double waves = ctas / slots;
                const double eff = waves / (double)(long long)(waves + 0.999999);
                const double score = eff - 0.02 * (c > 5 ? c - 5 : 5 - c);
                if (score > best) { best = score; rpc = c; }
            }
            if (const char *e = getenv("KMC_SWEEP_ROWS_PER_CTA")) rpc = atoi(e);      // tuning knob
            if (rpc < 1) rpc = 1;
            if (rpc > ROLL_MAX_ROWS) rpc = ROLL_MAX_ROWS;
            SweepRollParams q;
            q.d0 = h->d0; q.d1 = h->d1; q.n = h->n; q.n_replicas = h->R;
            q.status = h->d_status; q.slots = with_slots ? h->d_slots : nullptr; q.rowcnt = h->d_rowcnt;
            q.acc = h->d_acc; q.ticket = h->d_ticket; q.counts = h->d_counts;
            q.rows_per_cta = rpc; q.ctas_per_replica = (h->d0 + rpc - 1) / rpc; q.nplanes = nsl;
            q.rowmm = mm_split ? h->d_rowmm : nullptr;
            const unsigned grid = (unsigned)(R * q.ctas_per_replica);
            if (mm_split) {
                if (with_slots) kmc_sweep_roll_kernel<true, true><<<grid, 256, 0, h->stream>>>(q);
                else kmc_sweep_roll_kernel<false, true><<<grid, 256, 0, h->stream>>>(q);
            } else if (with_slots) kmc_sweep_roll_kernel<true><<<grid, 256, 0, h->stream>>>(q);
            else kmc_sweep_roll_kernel<false><<<grid, 256, 0, h->stream>>>(q);
            CU(cudaGetLastError());
            h->launches += 1;
            return KMC_OK;
        }
        Sweep16Params q;
        q.d0 = h->d0; q.d1 = h->d1; q.n = h->n; q.n_replicas = h->R;
        q.status = h->d_status; q.slots = with_slots ? h->d_slots : nullptr; q.rowcnt = h->d_rowcnt;
        q.acc = h->d_acc; q.ticket = h->d_ticket; q.counts = h->d_counts;
        q.ctas_per_replica = (h->n >> 4) / 256; q.nplanes = nsl; q.rowmm = mm_split ? h->d_rowmm : nullptr;      // (0: several replicas per CTA)
        const long long blocks = aligned_small ? ((long long)R * cpl_all) / 256 : (long long)R * q.ctas_per_replica;
        if (mm_split) kmc_sweep16n_kernel<true><<<(unsigned)blocks, 256, 0, h->stream>>>(q);     // (variant 2 has no MoveMonovacancy build)
        else if (h->sweep_variant == 2 && !aligned_small) kmc_sweep16_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(q);
        else kmc_sweep16n_kernel<false><<<(unsigned)blocks, 256, 0, h->stream>>>(q);
        CU(cudaGetLastError());
        h->launches += 1;
        return KMC_OK;
    }
    CU(cudaMemsetAsync(h->d_rowcnt, 0, R * h->d0 * RCG_STRIDE * sizeof(uint32_t), h->stream));
    SweepParams p;
    p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R;
    p.status = h->d_status; p.slots = with_slots ? h->d_slots : nullptr; p.rowcnt = h->d_rowcnt;
    if ((h->d1 & 3) == 0) {
        p.total_words = (long long)R * (h->n >> 2);
        const long long blocks = (p.total_words + 255) / 256;
        kmc_sweep_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    } else {
        p.total_words = 0;
        const long long blocks = ((long long)R * h->n + 255) / 256;
        kmc_sweep_kernel_generic<<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    }
    CU(cudaGetLastError());
    kmc_reduce_counts_kernel<<<h->R, 256, 0, h->stream>>>(h->d_rowcnt, h->d0, h->R, h->d_counts);
    CU(cudaGetLastError());
    h->launches += 2;
    return KMC_OK;
}

extern "C" int kmc_sweep(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    rc = sweep_launch(h, true); if (rc) return rc;
    h->swept = true;
    return KMC_OK;
}

extern "C" int kmc_get_counts(kmc_engine *h, int64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_counts: null buffer");
    if (!h->swept) return fail(KMC_ERR_ARG, "kmc_get_counts: call kmc_sweep first");
    CU(cudaMemcpyAsync(out, h->d_counts, (size_t)h->R * NCA * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_get_slots(kmc_engine *h, uint8_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_slots: null buffer");
    if (!h->swept || !h->d_slots) return fail(KMC_ERR_ARG, "kmc_get_slots: call kmc_sweep first");
    const int nsl = h->mm ? NSA : NS;                 // slot planes the sweep kernels of this engine write
    std::vector<uint8_t> planes((size_t)nsl * h->n);
    for (int r = 0; r < h->R; r++) {
        CU(cudaMemcpyAsync(planes.data(), h->d_slots + (size_t)r * nsl * h->n, planes.size(),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        uint8_t *o = out + (size_t)r * h->n * NSA;
        for (int i = 0; i < h->n; i++)
            for (int j = 0; j < NSA; j++) o[(size_t)i * NSA + j] = j < nsl ? planes[(size_t)j * h->n + i] : (uint8_t)KMC_SLOT_NONE;
    }
    return KMC_OK;
}

extern "C" int kmc_get_row_counts(kmc_engine *h, uint32_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_row_counts: null buffer");
    if (!h->swept) return fail(KMC_ERR_ARG, "kmc_get_row_counts: call kmc_sweep first");
    const bool split = h->sweep_split;
    const int rcs = (h->mm && !split) ? RCG_STRIDE_MM : RCG_STRIDE, ncx = (h->mm && !split) ? NCA : NC;
    std::vector<uint32_t> tmp((size_t)h->R * h->d0 * rcs), mmv(split ? (size_t)h->R * h->d0 * 4 : 0);
    CU(cudaMemcpyAsync(tmp.data(), h->d_rowcnt, tmp.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (split) CU(cudaMemcpyAsync(mmv.data(), h->d_rowmm, mmv.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < (size_t)h->R * h->d0; i++)
        for (int c = 0; c < NCA; c++)
            out[i * NCA + c] = c < ncx ? tmp[i * rcs + c] : (split && c >= C_MOVE_MONO ? mmv[i * 4 + c - C_MOVE_MONO] : 0u);
    return KMC_OK;
}

// ---- K3/K4 replica kernel ------------------------------------------------------------------------
static size_t record_size(int log_mode) {
    return log_mode == KMC_LOG_FULL ? sizeof(kmc_event_full) : log_mode == KMC_LOG_COMPACT ? sizeof(kmc_event_compact) : 0;
}

static int ensure_events(kmc_engine *h, size_t bytes) {
    if (bytes > h->events_cap) {
        if (h->d_events) CU(cudaFree(h->d_events));
        h->d_events = nullptr; h->events_cap = 0;
        CU(cudaMalloc(&h->d_events, bytes));
        h->events_cap = bytes;
    }
    // records a launch does not reach (everything after a zero-rate stop) read back as CLASS_NONE / site
    // 0xFFFFFFFF instead of whatever an earlier launch left there
    if (bytes) CU(cudaMemsetAsync(h->d_events, 0xFF, bytes, h->stream));
    return KMC_OK;
}

static int check_flags(kmc_engine *h, bool check_validate)
{
    std::vector<uint32_t> fl(h->R);
    CU(cudaMemcpyAsync(fl.data(), h->d_flags, fl.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int zero = -1, bad = -1;
    for (int r = 0; r < h->R; r++) {
        if (fl[r] & FLAG_STALL) return fail(KMC_ERR_CUDA, "no-incremental kernel: a CTA waited in vain for its replica's selector");
        if ((fl[r] & FLAG_VALIDATE) && bad < 0) bad = r;
        if ((fl[r] & FLAG_ZERO_RATE) && zero < 0) zero = r;
    }
    char b[256];
    if (check_validate && bad >= 0) {
        snprintf(b, sizeof b, "validation failed: incremental move tables of replica %d differ from a "
                              "from-scratch enumeration", bad);
        return fail(KMC_ERR_VALIDATE, b);
    }
    if (zero >= 0) {
        snprintf(b, sizeof b, "Cannot choose from total weight of zero! (replica %d)", zero);
        return fail(KMC_ERR_ZERO_RATE, b);
    }
    return KMC_OK;
}

static int replica_launch(kmc_engine *h, long long nsteps, uint64_t seed, uint32_t replica0,
                          const double *d_draws, int log_mode, void *host_events, int validate)
{
    if (nsteps < 0) return fail(KMC_ERR_ARG, "nsteps must be >= 0");
    if (log_mode < 0 || log_mode > 2) return fail(KMC_ERR_ARG, "bad log_mode");
    if (log_mode != KMC_LOG_NONE && !host_events) return fail(KMC_ERR_ARG, "log_mode set but events buffer is null");
    // rows of <= 64 sites run on the bit-sliced kernels (two replicas per warp by default), wider rows on
    // the byte-lattice kernel
    // (two replicas per warp pay off once one replica per warp would already fill the GPU; small batches and
    // single trajectories are latency-bound and run faster with a full warp each)
    // MoveMonovacancy live: only the byte-lattice kernel maintains its four classes (variants 1 and 3)
    // MoveMonovacancy live: the byte-lattice kernel (variants 1 and 3) maintains its four classes for any rule set and
    // shape; the half-warp kernel (variant 4) does when rows have <= 64 sites and one of the own-status classes is
    // disabled (17 classes, 16 lanes: class 16 takes that class's lane)
    int lane16 = -1;
    if (h->mm) {
        const int cand[8] = {0, 1, 7, 8, 9, 10, 11, 12};
        for (int i = 0; i < 8 && lane16 < 0; i++) if (h->order_pos[cand[i]] < 0) lane16 = cand[i];
    }
    const bool hw_mm_ok = h->mm && lane16 >= 0 && h->d1 <= 64 && 2 * replica_hw_bytes(h->d0, true) <= (size_t)h->max_smem_optin;
    // ... and the wide bit-plane kernel (variant 5) in its CTA-per-replica build (replica_wide2_kernel.cuh)
    const bool wide_geom = h->d1 > 64 && h->d1 <= 2048;
    const bool wide_mm_ok = h->mm && wide_geom && wide2_cta_bytes(h->d0, true) <= (size_t)h->max_smem_optin - 1024;
    if (h->mm && ((h->variant == 5 && !wide_mm_ok) || (h->variant == 4 && !hw_mm_ok)))
        return fail(KMC_ERR_ARG, "MoveMonovacancy is enabled: of the bit-sliced kernels, 4 (half-warp) needs rows of <= 64 sites "
                                 "and at least one own-status class disabled, 5 (wide bit planes) needs 64 < dim1 <= 2048");
    const bool wide_ok = h->mm ? wide_mm_ok : (wide_geom && wide_warp_bytes(h->d0) <= (size_t)h->max_smem_optin);
    if (h->variant == 5 && !wide_ok)
        return fail(KMC_ERR_ARG, "the wide bit-plane kernel needs 64 < dim1 <= 2048 and dim0 <= ~8700");
    // auto: lattices whose byte form does not fit shared memory, and batches too large for all replicas to be
    // resident with their byte lattices in shared memory (then the byte kernel runs in waves of a few warps per
    // SM: 1184 replicas of 256x256 take 11.9 us per event there, 4.7 us here)
    bool wide_auto = false;
    if (h->variant == 0 && wide_ok) {
        const size_t bbytes = replica_warp_bytes(h->d0, h->d1, h->mm);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const long long per_sm = bbytes <= (size_t)h->max_smem_optin ? (long long)((size_t)h->max_smem_optin / bbytes) : 0;
        // (and from about 215 x 215 sites even a single trajectory is faster here: the byte kernel's searches grow
        // with the row length, 3.1 / 3.5 / 4.0 / 5.1 us per event at 128 / 192 / 256 / 448 squared, this one stays at 3.7)
        // (round 2, CTA-per-replica build, `profiles/probe_200.py`: the byte kernel's in-row search grows with the row
        // length -- 3.3 / 3.5 / 3.65 / 3.9 us per event at 128 / 160 / 192 / 200 columns -- this one stays at 3.5-3.7 us)
        wide_auto = per_sm == 0 || (long long)h->R > per_sm * sms || h->n > 46000 || (h->R <= sms && h->d1 >= 196);
    }
    if (h->variant == 5 || wide_auto) {
        const int Wa = wide_anchor_words(h->d1), We = wide_row_words(h->d1), d0p = (h->d0 + 1) & ~1;
        const size_t plane_words = (size_t)h->R * N_TYPES * h->d0 * We;
        const int nct = wide_tables(h->mm);
        const size_t hier_elems = (size_t)h->R * nct * d0p;
        if (!h->d_planes) CU(cudaMalloc(&h->d_planes, plane_words * sizeof(unsigned long long)));
        if (!h->d_whier) {
            CU(cudaMalloc(&h->d_whier, hier_elems * sizeof(uint16_t)));
            // rows a >= d0 of the padded table are never written by the kernels but ARE compared by validate
            CU(cudaMemsetAsync(h->d_whier, 0, hier_elems * sizeof(uint16_t), h->stream));
        }
        if (!h->planes_valid) {
            int rc = ensure_bytes(h); if (rc) return rc;
            const long long cells = (long long)h->R * h->d0 * We;
            kmc_wide_pack_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, h->stream>>>(h->d_status, h->d_planes, h->R, h->d0, h->d1);
            CU(cudaGetLastError());
            h->launches += 1;
            h->planes_valid = true; h->whier_valid = false;
        }
        if (!h->whier_valid) {
            const long long rows = (long long)h->R * h->d0;
            kmc_wide_hier_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->d_planes, h->d_whier, h->R, h->d0, h->d1, d0p, nct);
            if (h->mm) kmc_wide_hier_mm_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->d_planes, h->d_whier, h->R, h->d0, h->d1, d0p);
            CU(cudaGetLastError());
            h->launches += h->mm ? 2 : 1;
            h->whier_valid = true;
        }
        const size_t wbytes = wide_warp_bytes(h->d0);
        int wpb = h->warps_per_block > 0 ? h->warps_per_block : (h->R + 147) / 148;    // spread small batches over the SMs
        if (wpb > 8) wpb = 8;
        while (wpb > 1 && wpb * wbytes > (size_t)h->max_smem_optin) wpb--;
        const size_t smem = wpb * wbytes;
        const size_t rs = record_size(log_mode);
        if (rs) { int rc = ensure_events(h, rs * (size_t)h->R * (size_t)nsteps); if (rc) return rc; }
        WideParams q;
        ReplicaParams &p = q.base;
        p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R;
        p.status = h->d_status;
        for (int c = 0; c < NCA; c++) { p.rate[c] = h->rate[c]; p.order_pos[c] = h->order_pos[c]; }
        p.nsteps = nsteps; p.seed = seed; p.replica0 = replica0;
        p.draws = d_draws; p.log_mode = log_mode; p.events = rs ? h->d_events : nullptr;
        p.steps_done = h->d_steps; p.hist = h->d_hist; p.hops = h->d_hops; p.ties = h->d_ties; p.flags = h->d_flags; p.any_flag = h->d_anyflag;
        p.validate = validate;
        p.clock = h->clock_mode ? h->d_clock : nullptr; p.clock_mode = h->clock_mode;
        p.tracer = h->tracer_on ? h->d_tracer : nullptr;
        p.hier = nullptr; p.hier_valid = 0; p.lane16 = -1;
        q.planes = h->d_planes; q.hier = h->d_whier; q.Wa = Wa; q.d0p = d0p;
        // small batches (at most one replica per SM): a whole CTA per replica, warps
        // split by role -- the latency build (replica_wide2_kernel.cuh); KMC_WIDE_CTA=0/1 forces the choice
        int sms2 = 148;
        cudaDeviceGetAttribute(&sms2, cudaDevAttrMultiProcessorCount, h->device);
        bool cta_per_replica = h->R <= sms2 && h->warps_per_block == 0;       // (one CTA per SM; beyond that one warp per replica wins)
        if (const char *e = getenv("KMC_WIDE_CTA")) cta_per_replica = atoi(e) != 0;
        if (h->mm) {
            // MoveMonovacancy: only the CTA-per-replica build maintains its classes (any batch size)
            const size_t smem2 = wide2_cta_bytes(h->d0, true);
            CU(cudaFuncSetAttribute(kmc_replica_wide2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            kmc_replica_wide2_kernel<true><<<(unsigned)h->R, 256, smem2, h->stream>>>(q);
        } else if (cta_per_replica && wide2_cta_bytes(h->d0) <= (size_t)h->max_smem_optin - 1024) {
            const size_t smem2 = wide2_cta_bytes(h->d0);
            CU(cudaFuncSetAttribute(kmc_replica_wide2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            kmc_replica_wide2_kernel<false><<<(unsigned)h->R, 256, smem2, h->stream>>>(q);
        } else {
            CU(cudaFuncSetAttribute(kmc_replica_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kmc_replica_wide_kernel<<<(unsigned)((h->R + wpb - 1) / wpb), wpb * 32, smem, h->stream>>>(q);
        }
        CU(cudaGetLastError());
        h->launches += 1;
#ifdef KMC_WIDE_TRACE
        {
            long long tr[16];
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaMemcpyFromSymbol(tr, g_wide_trace, sizeof tr));
            fprintf(stderr, "wide2 trace (cycles since the start of the event, mean of %lld events of replica 0):", tr[15]);
            for (int i = 1; i < 8; i++) fprintf(stderr, " %d:%.0f", i, (double)tr[i] / (double)(tr[15] ? tr[15] : 1));
            fprintf(stderr, "\n");
            memset(tr, 0, sizeof tr);
            CU(cudaMemcpyToSymbol(g_wide_trace, tr, sizeof tr));
        }
#endif
        if (validate) {          // the incrementally maintained row table against a fresh one from the planes
            if (!h->d_whier_chk) CU(cudaMalloc(&h->d_whier_chk, hier_elems * sizeof(uint16_t)));
            const long long rows = (long long)h->R * h->d0;
            CU(cudaMemsetAsync(h->d_whier_chk, 0, hier_elems * sizeof(uint16_t), h->stream));
            kmc_wide_hier_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->d_planes, h->d_whier_chk, h->R, h->d0, h->d1, d0p, nct);
            if (h->mm) kmc_wide_hier_mm_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(h->d_planes, h->d_whier_chk, h->R, h->d0, h->d1, d0p);
            kmc_wide_compare_kernel<<<(unsigned)((hier_elems + 255) / 256), 256, 0, h->stream>>>(
                h->d_whier, h->d_whier_chk, (long long)nct * d0p, h->R, h->d_flags);
            CU(cudaGetLastError());
            h->launches += h->mm ? 3 : 2;
        }
        h->swept = false; h->hier_valid = false;
        h->bytes_stale = true;
        if (rs) CU(cudaMemcpyAsync(host_events, h->d_events, rs * (size_t)h->R * (size_t)nsteps, cudaMemcpyDeviceToHost, h->stream));
        return KMC_OK;
    }
    { int rc = ensure_bytes(h); if (rc) return rc; }
    long long hw_min = 4096;
    if (const char *e = getenv("KMC_HW_MIN_REPLICAS")) hw_min = atoll(e);              // tuning knob
    const bool halfwarp = h->variant == 4 || (h->variant == 0 && hw_mm_ok && h->R >= hw_min) ||
                          (!h->mm && h->variant == 0 && h->d1 <= 64 && h->R >= hw_min &&
                           2 * replica_hw_bytes(h->d0) <= (size_t)h->max_smem_optin);
    if (halfwarp) {
        if (h->d1 > 64) return fail(KMC_ERR_ARG, "the bit-plane replica kernels need dim1 <= 64");
        if (nsteps > (1ll << 30)) return fail(KMC_ERR_ARG, "at most 2^30 events per launch of the half-warp kernel: split the run");
        const size_t hbytes = replica_hw_bytes(h->d0, h->mm);
        if (2 * hbytes > (size_t)h->max_smem_optin) return fail(KMC_ERR_ARG, "lattice too large for the half-warp kernel");
        // replicas (half-warps) per block: a multiple of 8 (4 warps, one per scheduler), at most 56 (28 warps of
        // 72 registers fill an SM); small batches get smaller blocks so that every SM has one
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        int hpb = (int)(((long long)h->R + sms - 1) / sms);
        hpb = (hpb + 7) & ~7;
        if (h->warps_per_block > 0) hpb = 2 * h->warps_per_block;
        if (const char *e = getenv("KMC_HW_REPLICAS_PER_BLOCK")) hpb = atoi(e) & ~1;     // tuning knob
        if (hpb > HW_MAX_THREADS / 16) hpb = HW_MAX_THREADS / 16;
        if (hpb < 2) hpb = 2;
        while (hpb > 8 && hpb * hbytes > (size_t)h->max_smem_optin) hpb -= 8;
        while (hpb > 2 && hpb * hbytes > (size_t)h->max_smem_optin) hpb -= 2;
        const size_t smem = hpb * hbytes;
        const size_t rs = record_size(log_mode);
        if (rs) { int rc = ensure_events(h, rs * (size_t)h->R * (size_t)nsteps); if (rc) return rc; }
        ReplicaParams p;
        p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R;
        p.status = h->d_status;
        for (int c = 0; c < NCA; c++) { p.rate[c] = h->rate[c]; p.order_pos[c] = h->order_pos[c]; }
        p.nsteps = nsteps; p.seed = seed; p.replica0 = replica0;
        p.draws = d_draws; p.log_mode = log_mode; p.events = rs ? h->d_events : nullptr;
        p.steps_done = h->d_steps; p.hist = h->d_hist; p.hops = h->d_hops; p.ties = h->d_ties; p.flags = h->d_flags; p.any_flag = h->d_anyflag;
        p.validate = validate;
        p.clock = h->clock_mode ? h->d_clock : nullptr; p.clock_mode = h->clock_mode;
        p.tracer = h->tracer_on ? h->d_tracer : nullptr;
        p.hier = nullptr; p.hier_valid = 0; p.lane16 = lane16;
        const unsigned blocks = (unsigned)((h->R + hpb - 1) / hpb);
        const bool plain = log_mode == KMC_LOG_NONE && !d_draws && !p.clock && !p.tracer;
        if (h->mm) {
            if (h->d0 == 64 && h->d1 == 64 && plain) {
                CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<64, 64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kmc_replica_hw_kernel<64, 64, true, true><<<blocks, hpb * 16, smem, h->stream>>>(p);
            } else if (h->d0 == 64 && h->d1 == 64) {
                CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<64, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kmc_replica_hw_kernel<64, 64, false, true><<<blocks, hpb * 16, smem, h->stream>>>(p);
            } else {
                CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<0, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kmc_replica_hw_kernel<0, 0, false, true><<<blocks, hpb * 16, smem, h->stream>>>(p);
            }
        } else if (h->d0 == 64 && h->d1 == 64 && plain) {
            CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<64, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kmc_replica_hw_kernel<64, 64, true><<<blocks, hpb * 16, smem, h->stream>>>(p);
        } else if (h->d0 == 64 && h->d1 == 64) {
            CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kmc_replica_hw_kernel<64, 64><<<blocks, hpb * 16, smem, h->stream>>>(p);
        } else {
            CU(cudaFuncSetAttribute(kmc_replica_hw_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kmc_replica_hw_kernel<0, 0><<<blocks, hpb * 16, smem, h->stream>>>(p);
        }
        CU(cudaGetLastError());
        h->launches += 1;
        h->swept = false; h->hier_valid = false; h->planes_valid = false;
        if (rs) CU(cudaMemcpyAsync(host_events, h->d_events, rs * (size_t)h->R * (size_t)nsteps, cudaMemcpyDeviceToHost, h->stream));
        return KMC_OK;
    }
    const bool bitplanes = h->variant == 2 || (h->variant == 0 && h->d1 <= 64 &&
                                               replica_bp_warp_bytes(h->d0, h->mm) <= (size_t)h->max_smem_optin);
    if (bitplanes && h->d1 > 64) return fail(KMC_ERR_ARG, "the bit-plane replica kernel needs dim1 <= 64");
    size_t wbytes = bitplanes ? replica_bp_warp_bytes(h->d0, h->mm) : replica_warp_bytes(h->d0, h->d1, h->mm);
    // lattices that do not fit shared memory stay in HBM/L2; only the row-count hierarchy is staged
    const bool global_state = h->variant == 3 || (!bitplanes && wbytes > (size_t)h->max_smem_optin);
    if (global_state) {
        if (bitplanes) return fail(KMC_ERR_ARG, "variant 3 (HBM-resident lattice) uses the byte-lattice kernel");
        wbytes = replica_rc_bytes(h->d0, h->mm);
        if (h->d1 > 10000) return fail(KMC_ERR_ARG, "rows longer than 10000 sites overflow the 16-bit row counts: use kmc_run_noinc");
    }
    if (wbytes > (size_t)h->max_smem_optin) {
        char b[256];
        snprintf(b, sizeof b, "lattice %dx%d needs %zu B of shared memory per replica (limit %d): use kmc_run_noinc",
                 h->d0, h->d1, wbytes, h->max_smem_optin);
        return fail(KMC_ERR_ARG, b);
    }
    // warps (replicas) per block: spread small batches over the SMs, at most 8
    int W = h->warps_per_block > 0 ? h->warps_per_block : (h->R + 147) / 148;
    if (W > 8) W = 8;
    if (W < 1) W = 1;
    while (W > 1 && W * wbytes > (size_t)h->max_smem_optin) W--;
    const size_t smem = W * wbytes;
    const size_t rs = record_size(log_mode);
    if (rs) { int rc = ensure_events(h, rs * (size_t)h->R * (size_t)nsteps); if (rc) return rc; }

    ReplicaParams p;
    p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R;
    p.status = h->d_status;
    for (int c = 0; c < NCA; c++) { p.rate[c] = h->rate[c]; p.order_pos[c] = h->order_pos[c]; }
    p.nsteps = nsteps; p.seed = seed; p.replica0 = replica0;
    p.draws = d_draws; p.log_mode = log_mode; p.events = rs ? h->d_events : nullptr;
    p.steps_done = h->d_steps; p.hist = h->d_hist; p.hops = h->d_hops; p.ties = h->d_ties; p.flags = h->d_flags; p.any_flag = h->d_anyflag;
    p.validate = validate;
    p.clock = h->clock_mode ? h->d_clock : nullptr; p.clock_mode = h->clock_mode;
        p.tracer = h->tracer_on ? h->d_tracer : nullptr;
    p.hier = nullptr; p.hier_valid = 0; p.lane16 = -1;
    if (global_state) {
        if (!h->d_hier) CU(cudaMalloc(&h->d_hier, (size_t)h->R * h->d0 * (h->mm ? 9 : 7) * sizeof(uint32_t)));
        p.hier = h->d_hier; p.hier_valid = h->hier_valid ? 1 : 0;
    }
    const unsigned blocks = (unsigned)((h->R + W - 1) / W);
#define KMC_LAUNCH(KERNEL)                                                                               \
    do {                                                                                                 \
        CU(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        CU(cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, 100));            \
        KERNEL<<<blocks, W * 32, smem, h->stream>>>(p);                                                   \
    } while (0)
    if (bitplanes) {
        // small batches cannot use the occupancy the 64-register build is compiled for: latency build (96 registers)
        const bool lat = getenv("KMC_BP_LAT") ? atoi(getenv("KMC_BP_LAT")) != 0 : h->R <= 16 * 148;
        if (h->mm) {
            if (h->d0 == 64 && h->d1 == 64) {
                if (lat) KMC_LAUNCH((kmc_replica_bp_kernel<64, 64, true, true>)); else KMC_LAUNCH((kmc_replica_bp_kernel<64, 64, false, true>));
            } else KMC_LAUNCH((kmc_replica_bp_kernel<0, 0, false, true>));
        } else if (h->d0 == 64 && h->d1 == 64) { if (lat) KMC_LAUNCH((kmc_replica_bp_kernel<64, 64, true>)); else KMC_LAUNCH((kmc_replica_bp_kernel<64, 64>)); }
        else KMC_LAUNCH((kmc_replica_bp_kernel<0, 0>));
    } else if (global_state) {
        if (h->mm) KMC_LAUNCH((kmc_replica_kernel<0, 0, true, true>)); else KMC_LAUNCH((kmc_replica_kernel<0, 0, true>));
    } else if (h->mm) {
        if (h->d0 == 64 && h->d1 == 64) KMC_LAUNCH((kmc_replica_kernel<64, 64, false, true>));
        else KMC_LAUNCH((kmc_replica_kernel<0, 0, false, true>));
    } else {
        if (h->d0 == 64 && h->d1 == 64) KMC_LAUNCH((kmc_replica_kernel<64, 64>));
        else KMC_LAUNCH((kmc_replica_kernel<0, 0>));
    }
#undef KMC_LAUNCH
    CU(cudaGetLastError());
    h->launches += 1;
    h->swept = false;
    h->hier_valid = global_state;          // only this path keeps it; any other lattice change drops it
    h->planes_valid = false;
    if (rs) {
        CU(cudaMemcpyAsync(host_events, h->d_events, rs * (size_t)h->R * (size_t)nsteps, cudaMemcpyDeviceToHost, h->stream));
    }
    return KMC_OK;
}

extern "C" int kmc_run(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, int log_mode,
                       void *events, int validate)
{
    int rc = bind(h); if (rc) return rc;
    rc = replica_launch(h, nsteps, seed, replica0, nullptr, log_mode, events, validate); if (rc) return rc;
    if (log_mode != KMC_LOG_NONE || validate) return check_flags(h, validate != 0);
    // statistics-only launches are asynchronous: a zero-rate stop raised by this launch is reported by the
    // next kmc_sync / kmc_check_status on the handle (one status word, see include/kmc_b200.h)
    return KMC_OK;
}

// synchronise, then report the handle's status word: KMC_ERR_ZERO_RATE if any replica has stopped since the
// flags were last cleared (kmc_reset_stats, or the lattice being replaced)
static int status_after_sync(kmc_engine *h)
{
    uint32_t any = 0;
    CU(cudaMemcpyAsync(&any, h->d_anyflag, sizeof any, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (any & (FLAG_ZERO_RATE | FLAG_VALIDATE)) return check_flags(h, true);     // names the replica
    return KMC_OK;
}

extern "C" int kmc_check_status(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    return status_after_sync(h);
}

extern "C" int kmc_step_draws(kmc_engine *h, const double *u, int64_t nsteps, int log_mode, void *events,
                              int validate)
{
    int rc = bind(h); if (rc) return rc;
    if (!u) return fail(KMC_ERR_ARG, "kmc_step_draws: null draws");
    if (nsteps < 0) return fail(KMC_ERR_ARG, "nsteps must be >= 0");
    const size_t bytes = (size_t)h->R * (size_t)nsteps * 2 * sizeof(double);
    if (bytes > h->draws_cap) {
        if (h->d_draws) CU(cudaFree(h->d_draws));
        h->d_draws = nullptr; h->draws_cap = 0;
        CU(cudaMalloc(&h->d_draws, bytes ? bytes : 16));
        h->draws_cap = bytes;
    }
    CU(cudaMemcpyAsync(h->d_draws, u, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = replica_launch(h, nsteps, 0, 0, h->d_draws, log_mode, events, validate); if (rc) return rc;
    return check_flags(h, validate != 0);
}

extern "C" int kmc_perform_given(kmc_engine *h, int replica, uint32_t site, int slot)
{
    int rc = bind(h); if (rc) return rc;
    if (replica < 0 || replica >= h->R || site >= (uint32_t)h->n || slot < 0 || slot >= NSA)
        return fail(KMC_ERR_ARG, "kmc_perform_given: bad replica / site / slot");
    rc = ensure_bytes(h); if (rc) return rc;
    kmc_apply_given_kernel<<<1, 1, 0, h->stream>>>(h->d_status + (size_t)replica * h->n, h->d0, h->d1,
                                                   (int)site, slot, h->d_result,
                                                   h->tracer_on ? h->d_tracer + (size_t)replica * h->n : nullptr);
    CU(cudaGetLastError());
    h->launches += 1;
    int res = 0;
    CU(cudaMemcpyAsync(&res, h->d_result, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (res) return fail(KMC_ERR_ARG, "kmc_perform_given: that move does not exist on the current lattice");
    // the lattice was edited in place: derived tables are stale, flags start afresh (tracers carry on)
    h->swept = false; h->hier_valid = false; h->planes_valid = false;
    CU(cudaMemsetAsync(h->d_flags, 0, (size_t)h->R * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->d_anyflag, 0, sizeof(uint32_t), h->stream));
    return KMC_OK;
}

// ---- the --no-incremental event on HBM-resident lattices ---------------------------------------
static int noinc_launch(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, const double *d_draws,
                        int log_mode, void *events);

extern "C" int kmc_run_noinc(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, int log_mode,
                             void *events)
{
    int rc = bind(h); if (rc) return rc;
    return noinc_launch(h, nsteps, seed, replica0, nullptr, log_mode, events);
}

extern "C" int kmc_step_draws_noinc(kmc_engine *h, const double *u, int64_t nsteps, int log_mode, void *events)
{
    int rc = bind(h); if (rc) return rc;
    if (!u) return fail(KMC_ERR_ARG, "kmc_step_draws_noinc: null draws");
    if (nsteps < 0) return fail(KMC_ERR_ARG, "nsteps must be >= 0");
    const size_t bytes = (size_t)h->R * (size_t)nsteps * 2 * sizeof(double);
    if (bytes > h->draws_cap) {
        if (h->d_draws) CU(cudaFree(h->d_draws));
        h->d_draws = nullptr; h->draws_cap = 0;
        CU(cudaMalloc(&h->d_draws, bytes ? bytes : 16));
        h->draws_cap = bytes;
    }
    CU(cudaMemcpyAsync(h->d_draws, u, bytes, cudaMemcpyHostToDevice, h->stream));
    return noinc_launch(h, nsteps, 0, 0, h->d_draws, log_mode, events);
}

static int noinc_launch(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, const double *d_draws,
                        int log_mode, void *events)
{
    int rc;
    if (nsteps < 0) return fail(KMC_ERR_ARG, "nsteps must be >= 0");
    if (nsteps > (1ll << 30)) return fail(KMC_ERR_ARG, "at most 2^30 no-incremental events per call: split the run");
    if (log_mode < 0 || log_mode > 2) return fail(KMC_ERR_ARG, "bad log_mode");
    if (log_mode != KMC_LOG_NONE && !events) return fail(KMC_ERR_ARG, "log_mode set but events buffer is null");
    const size_t rs = record_size(log_mode);
    if (rs) { rc = ensure_events(h, rs * (size_t)h->R * (size_t)nsteps); if (rc) return rc; }
    NoincParams p;
    p.d0 = h->d0; p.d1 = h->d1; p.n = h->n; p.n_replicas = h->R;
    p.status = h->d_status;
    for (int c = 0; c < NCA; c++) { p.rate[c] = h->rate[c]; p.order_pos[c] = h->order_pos[c]; }
    p.mm = h->mm ? 1 : 0; p.rc_stride = h->mm ? RCG_STRIDE_MM : RCG_STRIDE;
    p.seed = seed; p.replica0 = replica0; p.log_mode = log_mode; p.events = rs ? h->d_events : nullptr;
    p.nsteps = nsteps;
    p.steps_done = h->d_steps; p.hist = h->d_hist; p.hops = h->d_hops; p.ties = h->d_ties; p.flags = h->d_flags;
    p.any_flag = h->d_anyflag; p.draws = d_draws;
    p.clock = h->clock_mode ? h->d_clock : nullptr; p.clock_mode = h->clock_mode;
    p.tracer = h->tracer_on ? h->d_tracer : nullptr;
    // The event runs on the bit-sliced lattice (noinc_planes_kernel.cuh) -- 64 sites per operation in the enumeration
    // instead of 8, one kernel per event -- for every lattice with rows of at most 16384 sites: 9.7 us per event from
    // 64 x 64 to 1024 x 1024 (byte path 15-30 us), 14.5 us at 4096 x 4096 (`profiles/probe_noinc_small.py`).
    // (KMC_NOINC_PLANES=0/1 forces the choice.)
    bool on_planes = wide_anchor_words(h->d1) <= 256;
    if (const char *e = getenv("KMC_NOINC_PLANES")) on_planes = atoi(e) != 0 && wide_anchor_words(h->d1) <= 256;
    if (on_planes) {
        const int We = wide_row_words(h->d1);
        const int chunks = planes_chunks(h->d1), rpb = planes_rows_per_cta(h->d1), nblk = (h->d0 + rpb - 1) / rpb;
        if (!h->d_planes) CU(cudaMalloc(&h->d_planes, (size_t)h->R * N_TYPES * h->d0 * We * sizeof(unsigned long long)));
        if (!h->d_prowc) {
            CU(cudaMalloc(&h->d_prowc, (size_t)h->R * PRC_STRIDE * h->d0 * sizeof(uint32_t)));
            CU(cudaMalloc(&h->d_pblk, (size_t)h->R * nblk * PRC_STRIDE * sizeof(uint32_t)));
            CU(cudaMalloc(&h->d_pticket, (size_t)h->R * sizeof(unsigned int)));
            CU(cudaMemsetAsync(h->d_pticket, 0, (size_t)h->R * sizeof(unsigned int), h->stream));
            CU(cudaMalloc(&h->d_pepoch, (size_t)h->R * sizeof(unsigned int)));
        }
        if (!h->planes_valid) {
            rc = ensure_bytes(h); if (rc) return rc;
            const long long cells = (long long)h->R * h->d0 * We;
            kmc_wide_pack_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, h->stream>>>(h->d_status, h->d_planes, h->R, h->d0, h->d1);
            CU(cudaGetLastError());
            h->launches += 1;
            h->planes_valid = true;
        }
        NoincPlanesParams q;
        q.base = p; q.planes = h->d_planes; q.rowc = h->d_prowc; q.blk = h->d_pblk; q.counts = h->d_counts;
        q.ticket = h->d_pticket; q.epoch = h->d_pepoch; q.nblk = nblk; q.rpb = rpb; q.chunks = chunks;
        q.chunk_shift = 0; while ((1 << q.chunk_shift) < chunks) q.chunk_shift++;
        q.base.stage = 0;
        q.skip_select = getenv("KMC_NOINC_DEBUG_SKIP") ? 1 : 0;
        const bool pdl = !getenv("KMC_NOINC_NO_PDL");
        const auto dbg_t0 = std::chrono::steady_clock::now();
        // One cooperative launch for the whole call when every CTA can be resident at once (4096^2: 512 CTAs); else one
        // launch per event, chained by programmatic dependent launch.
        const unsigned grid = (unsigned)(h->R * nblk);
        bool persist = nsteps >= 2 && !getenv("KMC_NOINC_NO_PERSIST") && !q.skip_select;
        if (persist) {
            int per_sm = 0, coop = 0;
            if (h->mm) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmc_noinc_planes_event_kernel<true, true>, 256, 0));
            else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmc_noinc_planes_event_kernel<false, true>, 256, 0));
            CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
            int sms = 0;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
            if (!coop || (long long)per_sm * sms < (long long)grid) persist = false;
        }
        if (persist) {
            CU(cudaMemsetAsync(h->d_pepoch, 0, (size_t)h->R * sizeof(unsigned int), h->stream));
            CU(cudaMemsetAsync(h->d_pticket, 0, (size_t)h->R * sizeof(unsigned int), h->stream));   // (also after an aborted call)
            long long t0 = 0, t1 = nsteps;
            void *args[] = {(void *)&q, (void *)&t0, (void *)&t1};
            const void *fn = h->mm ? (const void *)kmc_noinc_planes_event_kernel<true, true>
                                   : (const void *)kmc_noinc_planes_event_kernel<false, true>;
            CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, h->stream));
            h->launches += 1;
        } else for (long long t = 0; t < nsteps; t++) {
            // sim.py:114-115: initialize() -- a full re-enumeration -- before every event; one launch does both
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = h->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
            const long long t1 = t + 1;
            if (h->mm) CU(cudaLaunchKernelEx(&cfg, kmc_noinc_planes_event_kernel<true, false>, q, t, t1));
            else CU(cudaLaunchKernelEx(&cfg, kmc_noinc_planes_event_kernel<false, false>, q, t, t1));
            h->launches += 1;
        }
#ifdef KMC_NOINC_TRACE
        {
            long long tr[16];
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaMemcpyFromSymbol(tr, g_noinc_trace, sizeof tr));
            fprintf(stderr, "noinc trace (cycles since the start of the event, mean of %lld events):", tr[15]);
            for (int i = 0; i < 14; i++) fprintf(stderr, " %d:%.0f", i, (double)tr[i] / (double)(tr[15] ? tr[15] : 1));
            fprintf(stderr, "\n");
            memset(tr, 0, sizeof tr);
            CU(cudaMemcpyToSymbol(g_noinc_trace, tr, sizeof tr));
        }
#endif
        if (getenv("KMC_NOINC_DEBUG_TIME"))
            fprintf(stderr, "noinc enqueue: %.2f us per event\n", 1e6 * std::chrono::duration<double>(std::chrono::steady_clock::now() - dbg_t0).count() / (double)(nsteps ? nsteps : 1));
        h->swept = false; h->hier_valid = false; h->whier_valid = false;
        h->bytes_stale = true;                     // the status bytes lag behind the planes until somebody asks for them
        if (rs) CU(cudaMemcpyAsync(events, h->d_events, rs * (size_t)h->R * (size_t)nsteps, cudaMemcpyDeviceToHost, h->stream));
        return check_flags(h, false);
    }
    // rows row-1 .. row+2 of the chosen row are staged in shared memory when they fit
    const size_t stage_bytes = 4 * (size_t)h->d1 <= (size_t)h->max_smem_optin - 4096 ? 4 * (size_t)h->d1 : 0;
    p.stage = stage_bytes ? 1 : 0;
    if (stage_bytes > 40 * 1024)
        CU(cudaFuncSetAttribute(kmc_noinc_select_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    for (long long t = 0; t < nsteps; t++) {
        rc = sweep_launch(h, false); if (rc) return rc;        // sim.py:114-115: initialize() every step
        p.rowcnt = h->d_rowcnt; p.counts = h->d_counts; p.t = t;
        p.rowmm = h->sweep_split ? h->d_rowmm : nullptr; p.rc_stride = (h->mm && !h->sweep_split) ? RCG_STRIDE_MM : RCG_STRIDE;
        kmc_noinc_select_apply_kernel<<<h->R, 256, stage_bytes, h->stream>>>(p);
        CU(cudaGetLastError());
        h->launches += 1;
    }
    h->swept = false; h->hier_valid = false; h->planes_valid = false;
    if (rs) CU(cudaMemcpyAsync(events, h->d_events, rs * (size_t)h->R * (size_t)nsteps, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h, false);
}

// ---- statistics -------------------------------------------------------------------------------------
extern "C" int kmc_get_histogram_per_replica(kmc_engine *h, int64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_histogram_per_replica: null buffer");
    CU(cudaMemcpyAsync(out, h->d_hist, (size_t)h->R * NCA * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_zobrist(kmc_engine *h, int bits, uint64_t seed, uint64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (bits < 1 || bits > 64) return fail(KMC_ERR_ARG, "kmc_zobrist: bits must be 1..64");
    if (!out) return fail(KMC_ERR_ARG, "kmc_zobrist: null buffer");
    rc = ensure_bytes(h); if (rc) return rc;
    unsigned long long *d_keys = nullptr;
    CU(cudaMalloc(&d_keys, (size_t)h->R * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_keys, 0, (size_t)h->R * sizeof(unsigned long long), h->stream);
    if (e == cudaSuccess) {
        const unsigned bx = (unsigned)std::min<long long>(((long long)h->n + 255) / 256, 64);
        kmc_zobrist_kernel<<<dim3((unsigned)h->R, bx), 256, 0, h->stream>>>(h->d_status, h->n, bits, seed, d_keys);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_keys, (size_t)h->R * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_keys);
    if (e != cudaSuccess) return fail(KMC_ERR_CUDA, cudaGetErrorString(e));
    h->launches += 1;
    return KMC_OK;
}

extern "C" int kmc_get_hops(kmc_engine *h, int64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_hops: null buffer");
    CU(cudaMemcpyAsync(out, h->d_hops, (size_t)h->R * 6 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_get_histogram(kmc_engine *h, int64_t *out)
{
    if (!h || !out) return fail(KMC_ERR_ARG, "kmc_get_histogram: null argument");
    std::vector<int64_t> tmp((size_t)h->R * NCA);
    int rc = kmc_get_histogram_per_replica(h, tmp.data()); if (rc) return rc;
    for (int c = 0; c < NCA; c++) out[c] = 0;
    for (int r = 0; r < h->R; r++) for (int c = 0; c < NCA; c++) out[c] += tmp[(size_t)r * NCA + c];
    return KMC_OK;
}

extern "C" int kmc_get_steps_done(kmc_engine *h, int64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_steps_done: null buffer");
    CU(cudaMemcpyAsync(out, h->d_steps, (size_t)h->R * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_get_ties(kmc_engine *h, int64_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_ties: null buffer");
    std::vector<int64_t> tmp(h->R);
    CU(cudaMemcpyAsync(tmp.data(), h->d_ties, tmp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = 0;
    for (int r = 0; r < h->R; r++) *out += tmp[r];
    return KMC_OK;
}

extern "C" int kmc_reset_stats(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    const size_t R = (size_t)h->R;
    CU(cudaMemsetAsync(h->d_steps, 0, R * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_hist, 0, R * NCA * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_hops, 0, R * 6 * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_ties, 0, R * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_flags, 0, R * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->d_anyflag, 0, sizeof(uint32_t), h->stream));
    if (h->d_clock) CU(cudaMemsetAsync(h->d_clock, 0, R * sizeof(double), h->stream));
    if (h->d_tracer) CU(cudaMemsetAsync(h->d_tracer, 0, R * h->n * sizeof(uint32_t), h->stream));
    return KMC_OK;
}

extern "C" int kmc_enable_clock(kmc_engine *h, int on)
{
    int rc = bind(h); if (rc) return rc;
    if (on && !h->d_clock) {
        CU(cudaMalloc(&h->d_clock, (size_t)h->R * sizeof(double)));
        CU(cudaMemsetAsync(h->d_clock, 0, (size_t)h->R * sizeof(double), h->stream));
    }
    if (on < 0 || on > 2) return fail(KMC_ERR_ARG, "kmc_enable_clock: 0 off, 1 mean residence time, 2 stochastic");
    h->clock_mode = on;
    return KMC_OK;
}

extern "C" int kmc_enable_tracers(kmc_engine *h, int on)
{
    int rc = bind(h); if (rc) return rc;
    if (on && !h->d_tracer) {
        CU(cudaMalloc(&h->d_tracer, (size_t)h->R * h->n * sizeof(uint32_t)));
        CU(cudaMemsetAsync(h->d_tracer, 0, (size_t)h->R * h->n * sizeof(uint32_t), h->stream));
    }
    h->tracer_on = on != 0;
    return KMC_OK;
}

// sum over the divacancies of a replica of |displacement|^2 = da^2 + da*db + db^2 (axial hex coordinates, unit
// lattice constant), and their number
__global__ void __launch_bounds__(256) kmc_msd_kernel(const uint8_t *status, const uint32_t *tracer, int n,
                                                     long long *sumsq, long long *count)
{
    const int rep = blockIdx.x;
    const uint8_t *st = status + (size_t)rep * n;
    const uint32_t *T = tracer + (size_t)rep * n;
    long long sq = 0, cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (st[i] == S_DIVAC) {
            const uint32_t v = T[i];
            const long long da = (short)(v & 0xFFFFu), db = (short)(v >> 16);
            sq += da * da + da * db + db * db;
            cnt += 1;
        }
    }
    for (int o = 16; o; o >>= 1) { sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o); cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(reinterpret_cast<unsigned long long *>(sumsq + rep), (unsigned long long)sq);
        atomicAdd(reinterpret_cast<unsigned long long *>(count + rep), (unsigned long long)cnt);
    }
}

extern "C" int kmc_get_msd(kmc_engine *h, int64_t *sum_sq, int64_t *n_tracers)
{
    int rc = bind(h); if (rc) return rc;
    if (!sum_sq || !n_tracers) return fail(KMC_ERR_ARG, "kmc_get_msd: null buffer");
    if (!h->d_tracer) return fail(KMC_ERR_ARG, "kmc_get_msd: call kmc_enable_tracers first");
    rc = ensure_bytes(h); if (rc) return rc;
    long long *d_out = nullptr;
    const size_t bytes = 2 * (size_t)h->R * sizeof(long long);
    CU(cudaMalloc(&d_out, bytes));
    cudaError_t e = cudaMemsetAsync(d_out, 0, bytes, h->stream);
    if (e == cudaSuccess) {
        kmc_msd_kernel<<<(unsigned)h->R, 256, 0, h->stream>>>(h->d_status, h->d_tracer, h->n, d_out, d_out + h->R);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sum_sq, d_out, bytes / 2, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_tracers, d_out + h->R, bytes / 2, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(KMC_ERR_CUDA, cudaGetErrorString(e));
    h->launches += 1;
    return KMC_OK;
}

// raw tracer table of one replica (parity / debug): [n] packed int16 pairs
extern "C" int kmc_get_tracers(kmc_engine *h, int replica, uint32_t *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out || replica < 0 || replica >= h->R) return fail(KMC_ERR_ARG, "kmc_get_tracers: bad argument");
    if (!h->d_tracer) return fail(KMC_ERR_ARG, "kmc_get_tracers: call kmc_enable_tracers first");
    CU(cudaMemcpyAsync(out, h->d_tracer + (size_t)replica * h->n, (size_t)h->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

extern "C" int kmc_get_clock(kmc_engine *h, double *out)
{
    int rc = bind(h); if (rc) return rc;
    if (!out) return fail(KMC_ERR_ARG, "kmc_get_clock: null buffer");
    if (!h->d_clock) return fail(KMC_ERR_ARG, "kmc_get_clock: call kmc_enable_clock first");
    CU(cudaMemcpyAsync(out, h->d_clock, (size_t)h->R * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}

// ---- measurement helpers -----------------------------------------------------------------------------
extern "C" int kmc_sync(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    return KMC_OK;
}
// L2 flush by READING a scratch buffer larger than L2: afterwards the cache holds clean lines of the
// scratch (evicting them costs nothing), so the next kernel starts cold without having to write back
// somebody else's dirty data during its timed region (which a memset-based flush would cause).
__global__ void __launch_bounds__(256) kmc_flush_read_kernel(const uint4 *buf, size_t n16, unsigned int *sink)
{
    unsigned int acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcg(buf + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;        // practically never: keeps the loads alive
}
extern "C" int kmc_flush_l2(kmc_engine *h, size_t bytes)
{
    int rc = bind(h); if (rc) return rc;
    if (bytes == 0) bytes = (size_t)256 << 20;
    bytes &= ~(size_t)15;
    if (bytes > h->flush_bytes) {
        if (h->d_flush) CU(cudaFree(h->d_flush));
        h->d_flush = nullptr; h->flush_bytes = 0;
        CU(cudaMalloc(&h->d_flush, bytes));
        CU(cudaMemsetAsync(h->d_flush, 0, bytes, h->stream));
        h->flush_bytes = bytes;
    }
    kmc_flush_read_kernel<<<148 * 8, 256, 0, h->stream>>>(reinterpret_cast<const uint4 *>(h->d_flush), bytes >> 4,
                                                         reinterpret_cast<unsigned int *>(h->d_result));
    CU(cudaGetLastError());
    return KMC_OK;
}
// Measurement helper: the sweep's store pattern alone (17 slot planes, one 128-bit streaming store per
// thread per plane, same grid), no loads and no arithmetic -- the practical ceiling for the sweep's writes.
__global__ void __launch_bounds__(256) kmc_store_pattern_kernel(uint8_t *slots, long long n, int rep_count)
{
    const long long cpl = n >> 4;
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long rep = g / cpl, cl = g - rep * cpl;
    if (rep >= rep_count) return;
    uint4 *O = reinterpret_cast<uint4 *>(slots + (size_t)rep * NS * n) + cl;
    const uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
    for (int j = 0; j < NS; j++) __stcs(O + (size_t)j * cpl, v);
}
extern "C" int kmc_probe_store_pattern(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    if (h->n & 15) return fail(KMC_ERR_ARG, "kmc_probe_store_pattern: needs dim0*dim1 % 16 == 0");
    if (h->mm) return fail(KMC_ERR_ARG, "kmc_probe_store_pattern: measures the 17-plane sweep (MoveMonovacancy disabled)");
    if (!h->d_slots) CU(cudaMalloc(&h->d_slots, (size_t)h->R * NS * h->n));
    const long long chunks = (long long)h->R * (h->n >> 4);
    kmc_store_pattern_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, h->stream>>>(h->d_slots, h->n, h->R);
    CU(cudaGetLastError());
    h->launches += 1;
    h->swept = false;
    return KMC_OK;
}
extern "C" int kmc_timer_start(kmc_engine *h)
{
    int rc = bind(h); if (rc) return rc;
    CU(cudaEventRecord(h->ev0, h->stream));
    return KMC_OK;
}
extern "C" int kmc_timer_stop(kmc_engine *h, float *elapsed_ms)
{
    int rc = bind(h); if (rc) return rc;
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return KMC_OK;
}
extern "C" int kmc_get_launch_count(kmc_engine *h, int64_t *out)
{
    if (!h || !out) return fail(KMC_ERR_ARG, "null argument");
    *out = h->launches;
    return KMC_OK;
}
extern "C" int kmc_get_device_pointers(kmc_engine *h, void **status, void **slots, void **row_counts, void **histogram)
{
    if (!h) return fail(KMC_ERR_ARG, "null engine handle");
    if (status) { int rc = bind(h); if (rc) return rc; rc = ensure_bytes(h); if (rc) return rc; *status = h->d_status; }
    if (slots) *slots = h->d_slots;
    if (row_counts) *row_counts = h->d_rowcnt;
    if (histogram) *histogram = h->d_hist;
    return KMC_OK;
}
extern "C" int kmc_set_warps_per_block(kmc_engine *h, int warps)
{
    if (!h || warps < 0 || warps > 8) return fail(KMC_ERR_ARG, "warps per block must be 0..8");
    h->warps_per_block = warps;
    return KMC_OK;
}
extern "C" int kmc_set_sweep_kernel(kmc_engine *h, int variant)
{
    if (!h || variant < 0 || variant > 2)
        return fail(KMC_ERR_ARG, "variant must be 0 (auto), 1 (nibble/PRMT, one row per CTA pass) or 2 (byte-parallel)");
    h->sweep_variant = variant;
    return KMC_OK;
}
extern "C" int kmc_set_sweep_double_buffer(kmc_engine *h, int on)
{
    if (!h) return fail(KMC_ERR_ARG, "null engine handle");
    h->slots_double = on != 0;
    return KMC_OK;
}
extern "C" int kmc_set_replica_kernel(kmc_engine *h, int variant)
{
    if (!h || variant < 0 || variant > 5)
        return fail(KMC_ERR_ARG, "variant must be 0 (auto), 1 (byte lattice), 2 (bit planes, warp per replica), "
                                 "3 (byte lattice in HBM), 4 (bit planes, half-warp per replica) or 5 (wide bit planes in HBM)");
    h->variant = variant;
    return KMC_OK;
}

#include "hostcheck.cuh"
