// replica_wide_kernel.cuh -- K3 + K4 for lattices too large for shared memory (e.g. 1024 x 1024,
// BASELINE config 5): the bit-sliced algorithm of replica_bp_kernel.cuh with the six predicate planes held
// in HBM/L2 (rows of W = d1/64 words) and only the class-major row-count table in shared memory.
//
// One warp owns one replica.  What keeps an event cheap on a 1 Mi-site lattice is that everything is local
// to a handful of 64-site words:
//   * site search inside the chosen row: one lane per word builds the six MoveDivacancy direction masks of
//     its word from a 3x3 neighbourhood of plane words (one round of independent loads), per-direction
//     totals come from six warp reductions, one warp scan finds the word, a popcount-select the bit;
//   * refresh: only the (row, word) cells within 3 rows and 2 columns of the touched nodes can change;
//     one lane per cell re-derives the cell's counts of the five neighbourhood classes before and after the
//     move and adds the difference to the row table (packed 16-bit shared-memory atomics) -- independent of
//     the lattice size.
// Pack / unpack (status bytes <-> planes) and the initial row table are separate grid-wide kernels, run
// only when the lattice was replaced; planes and table persist between launches.
#pragma once
#include "replica_bp_kernel.cuh"

namespace kmc {

struct WideParams {
    ReplicaParams base;
    u64 *planes;                     // [R][N_TYPES][d0][W]
    uint16_t *hier;                  // [R][NC][d0p] row counts, persists between launches
    int W, d0p;
};

__host__ __device__ inline size_t wide_warp_bytes(int d0) {
    const size_t d0p = ((size_t)d0 + 1) & ~(size_t)1;
    return ((NC * d0p * 2 + 15) & ~(size_t)15) + NC * 32 * sizeof(uint32_t);      // row table + 32 block sums per class
}

struct WGeom {
    const u64 *P; int d0, W;
    __device__ __forceinline__ u64 word(int pl, int a, int w) const { return P[((size_t)pl * d0 + a) * W + w]; }
};

// plane words of rows a-1, a, a+1 (Z and D) and a+2 (D), words w-1, w, w+1 (periodic)
struct Nbhd { u64 z[3][3], d[4][3]; };
template <bool WITH_ROW2 = true>
__device__ __forceinline__ void load_nbhd(const WGeom &g, int a, int w, Nbhd &n) {
    const int wm = w == 0 ? g.W - 1 : w - 1, wp = w + 1 == g.W ? 0 : w + 1;
    const int ws[3] = {wm, w, wp};
#pragma unroll
    for (int r = 0; r < (WITH_ROW2 ? 4 : 3); r++) {
        const int ar = wrap1(wrap1(a + r - 1, g.d0), g.d0);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            n.d[r][c] = g.word(T_D, ar, ws[c]);
            if (r < 3) n.z[r][c] = g.word(T_Z, ar, ws[c]);
        }
    }
}
// word whose bit i is column 64*w + i + db of a row given as its words (w-1, w, w+1); |db| <= 2
__device__ __forceinline__ u64 shifted(const u64 x[3], int db) {
    if (db == 0) return x[1];
    if (db > 0) return (x[1] >> db) | (x[2] << (64 - db));
    return (x[1] << -db) | (x[0] >> (64 + db));
}
// MoveDivacancy direction k of the cell: present / near / far masks
__device__ __forceinline__ void cell_dir_masks(const Nbhd &n, int k, u64 &present, u64 &near, u64 &far) {
    present = n.d[1][1] & shifted(n.z[nib(NBR_DA, k) + 1], nib(NBR_DB, k));
    near = shifted(n.d[nib(NEAR_DA, k) + 1], nib(NEAR_DB, k));
    far = shifted(n.d[nib(FAR_DA, k) + 1], nib(FAR_DB, k));
}
// the same with a run-time direction (keeps the neighbourhood in registers)
__device__ __forceinline__ void cell_dir_masks_dyn(const Nbhd &n, int k, u64 &present, u64 &near, u64 &far) {
    switch (k) {
        case 0: cell_dir_masks(n, 0, present, near, far); break;
        case 1: cell_dir_masks(n, 1, present, near, far); break;
        case 2: cell_dir_masks(n, 2, present, near, far); break;
        case 3: cell_dir_masks(n, 3, present, near, far); break;
        case 4: cell_dir_masks(n, 4, present, near, far); break;
        default: cell_dir_masks(n, 5, present, near, far); break;
    }
}
__device__ __forceinline__ void cell_trefoil_masks(const Nbhd &n, u64 &up, u64 &dn) {
    const u64 both = n.d[1][1] & n.d[3][1];
    up = both & shifted(n.d[1], 2);
    dn = both & shifted(n.d[3], -2);
}
// counts of classes 2..5 (16-bit pairs) and 6 of the sites of one cell
__device__ __forceinline__ void cell_counts(const Nbhd &n, uint32_t &k01, uint32_t &k23, uint32_t &tre) {
    k01 = 0; k23 = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        u64 p, ne, fa;
        cell_dir_masks(n, k, p, ne, fa);
        uint32_t a01, a23;
        kind_counts(p, ne, fa, a01, a23);
        k01 += a01; k23 += a23;
    }
    u64 up, dn;
    cell_trefoil_masks(n, up, dn);
    tre = __popcll(up) + __popcll(dn);
}
// position of the r-th set bit (0-based) of a 64-bit word, r < popc(m): popcount bisection
__device__ __forceinline__ int nth_bit64(u64 m, uint32_t r) {
    uint32_t x = (uint32_t)m;
    int base = 0;
    uint32_t c = __popc(x);
    if (r >= c) { r -= c; x = (uint32_t)(m >> 32); base = 32; }
    c = __popc(x & 0xFFFFu); if (r >= c) { r -= c; x >>= 16; base += 16; }
    c = __popc(x & 0xFFu);   if (r >= c) { r -= c; x >>= 8;  base += 8; }
    c = __popc(x & 0xFu);    if (r >= c) { r -= c; x >>= 4;  base += 4; }
    c = __popc(x & 0x3u);    if (r >= c) { r -= c; x >>= 2;  base += 2; }
    c = x & 1u;              if (r >= c) base += 1;
    return base;
}

// ---- grid-wide helpers ----------------------------------------------------------------------------------
// status bytes -> planes: one thread per (replica, row, word)
__global__ void kmc_wide_pack_kernel(const uint8_t *status, u64 *planes, int R, int d0, int d1)
{
    const int W = d1 >> 6;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0 * W) return;
    const int w = (int)(g % W);
    const int a = (int)((g / W) % d0);
    const long long rep = g / ((long long)W * d0);
    const uint4 *src = reinterpret_cast<const uint4 *>(status + (rep * d0 + a) * (long long)d1 + 64 * w);
    u64 out[N_TYPES] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
        const uint4 v = src[q4];
        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int pl = status_plane((wds[j] >> (8 * i)) & 0xF);
                const int bit = 16 * q4 + 4 * j + i;
#pragma unroll
                for (int q = 0; q < N_TYPES; q++) out[q] |= (u64)(pl == q) << bit;
            }
    }
#pragma unroll
    for (int q = 0; q < N_TYPES; q++) planes[((rep * N_TYPES + q) * d0 + a) * W + w] = out[q];
}

// planes -> status bytes: one thread per (replica, row, word); trefoil members are recognised from the
// anchor planes of the rows / columns they sit relative to
__global__ void kmc_wide_unpack_kernel(const u64 *planes, uint8_t *status, int R, int d0, int d1)
{
    const int W = d1 >> 6;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0 * W) return;
    const int w = (int)(g % W);
    const int a = (int)((g / W) % d0);
    const long long rep = g / ((long long)W * d0);
    WGeom G{planes + rep * N_TYPES * d0 * W, d0, W};
    const int wm = w == 0 ? W - 1 : w - 1, wp = w + 1 == W ? 0 : w + 1;
    const int am2 = wrap1(a - 2, d0);
    const u64 z = G.word(T_Z, a, w), dv = G.word(T_D, a, w), m1 = G.word(T_M1, a, w), m2 = G.word(T_M2, a, w);
    const u64 a4 = G.word(T_A4, a, w), a7 = G.word(T_A7, a, w);
    // members: Up anchor q=(x,y): (x+2,y) role 1 -> code 5, (x,y+2) role 2 -> code 6;
    //          Down anchor:        (x+2,y) role 1 -> code 8, (x+2,y-2) role 2 -> code 9
    const u64 c5 = G.word(T_A4, am2, w);
    const u64 x4[3] = {G.word(T_A4, a, wm), a4, G.word(T_A4, a, wp)};
    const u64 c6 = shifted(x4, -2);                                      // anchor two columns to the left
    const u64 c8 = G.word(T_A7, am2, w);
    const u64 x7[3] = {G.word(T_A7, am2, wm), c8, G.word(T_A7, am2, wp)};
    const u64 c9 = shifted(x7, 2);                                       // anchor at (x-2, y+2)
    uint4 *dst = reinterpret_cast<uint4 *>(status + (rep * d0 + a) * (long long)d1 + 64 * w);
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t word = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int b = 16 * q4 + 4 * j + i;
                const uint32_t s = ((z >> b) & 1) ? 0u : ((dv >> b) & 1) ? 3u : ((m1 >> b) & 1) ? 1u : ((m2 >> b) & 1) ? 2u
                                 : ((a4 >> b) & 1) ? 4u : ((a7 >> b) & 1) ? 7u : ((c5 >> b) & 1) ? 5u
                                 : ((c6 >> b) & 1) ? 6u : ((c8 >> b) & 1) ? 8u : 9u;
                word |= s << (8 * i);
            }
            o[j] = word;
        }
        dst[q4] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    (void)c9;
}

// row table from the planes: one thread per (replica, row)
__global__ void kmc_wide_hier_kernel(const u64 *planes, uint16_t *hier, int R, int d0, int d1, int d0p)
{
    const int W = d1 >> 6;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0) return;
    const int a = (int)(g % d0);
    const long long rep = g / d0;
    WGeom G{planes + rep * N_TYPES * d0 * W, d0, W};
    uint32_t pz = 0, pd = 0, pm = 0, pa = 0, k01 = 0, k23 = 0, tre = 0;
    for (int w = 0; w < W; w++) {
        pz += __popcll(G.word(T_Z, a, w)); pd += __popcll(G.word(T_D, a, w));
        pm += __popcll(G.word(T_M1, a, w) | G.word(T_M2, a, w));
        pa += __popcll(G.word(T_A4, a, w) | G.word(T_A7, a, w));
        Nbhd n;
        load_nbhd(G, a, w, n);
        uint32_t c01, c23, ct;
        cell_counts(n, c01, c23, ct);
        k01 += c01; k23 += c23; tre += ct;
    }
    uint16_t *H = hier + rep * NC * d0p;
    H[C_CREATE_DIV * d0p + a] = (uint16_t)pz; H[C_FILL_DIV * d0p + a] = (uint16_t)pd;
    H[C_MOVE_DIV * d0p + a] = (uint16_t)k01; H[(C_MOVE_DIV + 1) * d0p + a] = (uint16_t)(k01 >> 16);
    H[(C_MOVE_DIV + 2) * d0p + a] = (uint16_t)k23; H[(C_MOVE_DIV + 3) * d0p + a] = (uint16_t)(k23 >> 16);
    H[C_CREATE_TREF * d0p + a] = (uint16_t)tre; H[C_DESTROY_TREF * d0p + a] = (uint16_t)pa;
    H[C_CMONO_FROM_DOUBLE * d0p + a] = (uint16_t)(2 * pz); H[C_CMONO_MAKE_EMPTY * d0p + a] = (uint16_t)pm;
    H[C_FMONO_MAKE_DOUBLE * d0p + a] = (uint16_t)pm; H[C_FMONO_FROM_EMPTY * d0p + a] = (uint16_t)(2 * pd);
    H[C_FLIP * d0p + a] = (uint16_t)pm;
}

// compare two row tables (validate): flags[rep] |= FLAG_VALIDATE on any difference
__global__ void kmc_wide_compare_kernel(const uint16_t *a, const uint16_t *b, long long per_rep, int R, uint32_t *flags)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= per_rep * R) return;
    if (a[g] != b[g]) atomicOr(&flags[g / per_rep], FLAG_VALIDATE);
}

// ---- the fused event kernel ------------------------------------------------------------------------------
// Refresh tasks are data-driven so that lanes working on different directions run the same instructions.
// An operand is "plane pl of row a + da, columns shifted by db": pl | (da + 2) << 1 | (db + 2) << 4 (7 bits);
// a task has three operands (7 bits each):
//   k < 6  (MoveDivacancy direction k): Z at the target, D at the near and far bystanders
//   k == 6 (trefoil shapes):            D at (+2, 0), D at (0, +2), D at (+2, -2)
__device__ __forceinline__ uint32_t operand_desc(int pl, int da, int db) {
    return (uint32_t)pl | ((uint32_t)(da + 2) << 1) | ((uint32_t)(db + 2) << 4);
}
__device__ __forceinline__ uint32_t task_desc(int k) {
    if (k < 6)
        return operand_desc(T_Z, nib(NBR_DA, k), nib(NBR_DB, k)) | (operand_desc(T_D, nib(NEAR_DA, k), nib(NEAR_DB, k)) << 7) |
               (operand_desc(T_D, nib(FAR_DA, k), nib(FAR_DB, k)) << 14);
    return operand_desc(T_D, 2, 0) | (operand_desc(T_D, 0, 2) << 7) | (operand_desc(T_D, 2, -2) << 14);
}
// the operand's word for anchor word w of anchor row a: one 64-bit load, one 32-bit load of the half word that
// shifts in from the neighbouring word, two funnel shifts
__device__ __forceinline__ u64 operand_word(const u64 *P, int d0, int W, int a, int w, uint32_t od) {
    const int pl = (int)(od & 1u), da = (int)((od >> 1) & 7u) - 2, db = (int)((od >> 4) & 7u) - 2;
    int ar = a + da;
    ar = ar < 0 ? ar + d0 : (ar >= d0 ? ar - d0 : ar);
    const uint32_t base = (uint32_t)(pl * d0 + ar) * (uint32_t)W;
    const uint2 x1 = *reinterpret_cast<const uint2 *>(P + base + w);
    const int wn = db < 0 ? (w == 0 ? W - 1 : w - 1) : (w + 1 == W ? 0 : w + 1);
    const uint32_t nbh = reinterpret_cast<const uint32_t *>(P + base + wn)[db < 0 ? 1 : 0];
    const uint32_t h0 = db < 0 ? nbh : x1.x, h1 = db < 0 ? x1.x : x1.y, h2 = db < 0 ? x1.y : nbh;
    const uint32_t sh = (uint32_t)(db < 0 ? 32 + db : db);
    return (u64)__funnelshift_r(h0, h1, sh) | ((u64)__funnelshift_r(h1, h2, sh) << 32);
}
// packed counts of a task: classes (2,3) in c01, (4,5) in c23 for k < 6; class 6 in c45 for k == 6
__device__ __forceinline__ void task_counts(const u64 *P, int d0, int W, int a, int w, int k, uint32_t desc,
                                            uint32_t &c01, uint32_t &c23, uint32_t &c45) {
    const u64 D0 = P[(uint32_t)(T_D * d0 + a) * (uint32_t)W + w];
    const u64 A = operand_word(P, d0, W, a, w, desc & 0x7Fu);
    const u64 Bv = operand_word(P, d0, W, a, w, (desc >> 7) & 0x7Fu);
    const u64 C = operand_word(P, d0, W, a, w, (desc >> 14) & 0x7Fu);
    const u64 first = D0 & A;
    c01 = 0; c23 = 0; c45 = 0;
    if (k < 6) kind_counts(first, Bv, C, c01, c23);
    else c45 = __popcll(first & Bv) + __popcll(first & C);
}
// sum over the warp of packed pairs of signed 16-bit differences (x - y per half); returns the two sums
__device__ __forceinline__ void reduce_packed_diff(uint32_t x, uint32_t y, int &lo, int &hi) {
    const uint32_t s = __reduce_add_sync(FULL, x - y);
    lo = (int)(int16_t)(s & 0xFFFFu);
    hi = (int)(int16_t)((s - (uint32_t)lo) >> 16);
}

__global__ void __launch_bounds__(256, 1) kmc_replica_wide_kernel(const WideParams wp)
{
    const ReplicaParams &p = wp.base;
    const int d0 = p.d0, d1 = p.d1, W = wp.W, d0p = wp.d0p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * (blockDim.x >> 5) + warp;
    if (rep >= p.n_replicas) return;

    extern __shared__ uint4 smem_raw[];
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)warp * wide_warp_bytes(d0));
    uint32_t *rc32 = reinterpret_cast<uint32_t *>(rc16);
    uint32_t *bs = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(rc16) + ((NC * (size_t)d0p * 2 + 15) & ~(size_t)15));
    u64 *P = wp.planes + (size_t)rep * N_TYPES * d0 * W;
    WGeom G{P, d0, W};
    uint16_t *ghier = wp.hier + (size_t)rep * NC * d0p;
    for (int i = lane; i < (NC * d0p) / 2; i += 32) rc32[i] = reinterpret_cast<const uint32_t *>(ghier)[i];
    __syncwarp();
    // rows are grouped in 32 blocks of B = 2^bshift rows; bs[c][i] = events of class c in block i
    int bshift = 0;
    while ((32 << bshift) < d0) bshift++;
    const int B = 1 << bshift;
    for (int c = 0; c < NC; c++) {
        uint32_t sum = 0;
        for (int q = 0; q < B; q++) { const int a = lane * B + q; if (a < d0) sum += rc16[c * d0p + a]; }
        bs[c * 32 + lane] = sum;
    }
    __syncwarp();

    const int myc = lane < NC ? lane : 0;
    const double rate = (lane < NC && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NC && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    int tot = 0;
    if (lane < NC)
        for (int i = 0; i < 32; i++) tot += (int)bs[lane * 32 + i];
    const uint32_t my_g1 = (lane < NC && !(lane >= C_MOVE_DIV && lane <= C_CREATE_TREF)) ? g1_const(lane) : 0u;
    unsigned long long hist = 0, hops = 0, n_ties = 0;
    PickState pstate = pick_state_init(lane);
    double tclock = 0.0;
    const uint32_t my_desc = task_desc(lane < 7 ? lane : 6);          // lane k < 7 holds the descriptor of task kind k
    const int r2 = B > 32 ? B >> 5 : 1;                 // rows per lane in the second level of the row search

    const unsigned long long step_base = p.steps_done[rep];
    double mu1 = 0.0, mu2 = 0.0;
    long long t = 0;
    uint32_t flag = 0;

    for (; t < p.nsteps; t++) {
        if ((t & 31) == 0) {
            const long long tt = t + lane;
            if (tt < p.nsteps) {
                if (p.draws) {
                    const double2 u = *reinterpret_cast<const double2 *>(
                        p.draws + ((size_t)rep * (size_t)p.nsteps + (size_t)tt) * 2);
                    mu1 = u.x; mu2 = u.y;
                } else {
                    kmc_draw(p.seed, p.replica0 + (uint32_t)rep, step_base + (unsigned long long)tt, mu1, mu2);
                }
            }
        }
        const double u1 = __shfl_sync(FULL, mu1, (int)(t & 31));
        const double u2 = __shfl_sync(FULL, mu2, (int)(t & 31));
        const double w = __dmul_rn((double)tot, rate);
        const ClassPick pick = pick_class(w, pos, pstate, u1, lane);
        if (pick.zero) {
            flag |= FLAG_ZERO_RATE;
            if (p.log_mode != KMC_LOG_NONE && lane == 0) {
                const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)t;
                if (p.log_mode == KMC_LOG_FULL) {
                    kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                    rec->site = 0xFFFFFFFFu; rec->cls = KMC_CLASS_NONE; rec->slot = KMC_SLOT_NONE;
                    rec->flags = 0; rec->total_rate = 0.0; rec->pad = 0;
                } else {
                    kmc_event_compact e;
                    e.site = 0xFFFFFFFFu; e.cls = KMC_CLASS_NONE; e.slot = KMC_SLOT_NONE; e.flags = 0;
                    e.total_rate = 0.0;
                    reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
                }
            }
            break;
        }
        const int csel = pick.csel;
        if (p.clock) tclock = __dadd_rn(tclock, __ddiv_rn(1.0, pick.total));
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) +
                                  ((size_t)rep * (size_t)p.nsteps + (size_t)t);
            if (lane < NC) rec->counts[lane] = (uint32_t)tot;
        }
        const int cnt = __shfl_sync(FULL, tot, csel);
        uint32_t r;
        {
            long long rr = (long long)__dmul_rn(u2, (double)cnt);
            if (rr > (long long)cnt - 1) rr = (long long)cnt - 1;
            r = (uint32_t)rr;
        }
        // ---- row: block of rows, then the row inside it -------------------------------------------------
        int row;
        {
            int L = 0;
            warp_rank_search(bs[csel * 32 + lane], r, L, lane);
            const int a0 = (L << bshift) + lane * r2;
            uint32_t sum = 0;
            for (int q = 0; q < r2; q++)
                if (lane * r2 + q < B && a0 + q < d0) sum += rc16[csel * d0p + a0 + q];
            int L2 = 0;
            warp_rank_search(sum, r, L2, lane);
            row = (L << bshift) + L2 * r2;
            for (int q = 0; q < r2 - 1; q++) {
                const uint32_t v = rc16[csel * d0p + row];
                if (r < v) break;
                r -= v; row++;
            }
        }
        // the refresh below reads Z and D of rows row-3 .. row+3: start those lines moving now, under the site search
        {
            const int lines = (W * 8 + 127) >> 7;              // 128-byte lines per plane row (<= 2)
            const int j = lane >> 2, pl = (lane >> 1) & 1, li = lane & 1;
            if (j < 7 && li < lines) {
                const int a = wrap1(wrap1(row - 3 + j, d0), d0);
                asm volatile("prefetch.global.L1 [%0];" :: "l"(P + ((size_t)pl * d0 + a) * W + li * 16));
            }
        }
        // ---- slot and site inside the row: one lane per word (W <= 32), slots in class order ------------------
        int col, slot, s0;
        {
            u64 m[6] = {0, 0, 0, 0, 0, 0};                 // this lane's word: mask of the class per slot
            const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 : 0);
            if (lane < W) {
                if (group == 1) {
                    Nbhd nb;
                    load_nbhd<false>(G, row, lane, nb);
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        u64 pr, ne, fa;
                        cell_dir_masks(nb, k, pr, ne, fa);
                        m[k] = kind_select(pr, ne, fa, csel - C_MOVE_DIV);
                    }
                } else if (group == 2) {
                    const int a2 = wrap1(row + 2, d0), wm = lane == 0 ? W - 1 : lane - 1, wq = lane + 1 == W ? 0 : lane + 1;
                    const u64 da[3] = {0, G.word(T_D, row, lane), G.word(T_D, row, wq)};
                    const u64 db[3] = {G.word(T_D, a2, wm), G.word(T_D, a2, lane), 0};
                    const u64 both = da[1] & db[1];
                    m[0] = both & shifted(da, 2);
                    m[1] = both & shifted(db, -2);
                } else {
                    int pa = T_M1, pb = -1;
                    switch (csel) {
                        case C_CREATE_DIV: pa = T_Z; break;
                        case C_FILL_DIV: pa = T_D; break;
                        case C_DESTROY_TREF: pa = T_A4; pb = T_A7; break;
                        case C_CMONO_FROM_DOUBLE: pa = T_Z; pb = T_Z; break;
                        case C_CMONO_MAKE_EMPTY: pa = T_M2; pb = T_M1; break;
                        case C_FMONO_MAKE_DOUBLE: pa = T_M1; pb = T_M2; break;
                        case C_FMONO_FROM_EMPTY: pa = T_D; pb = T_D; break;
                        default: break;
                    }
                    m[0] = G.word(pa, row, lane);
                    if (csel == C_FLIP) m[0] |= G.word(T_M2, row, lane);
                    if (pb >= 0) m[1] = G.word(pb, row, lane);
                }
            }
            const uint32_t c01 = __popcll(m[0]) | (__popcll(m[1]) << 16);
            int j = 0;
            if (group == 1) {
                const uint32_t c23 = __popcll(m[2]) | (__popcll(m[3]) << 16), c45 = __popcll(m[4]) | (__popcll(m[5]) << 16);
                const uint32_t t01 = __reduce_add_sync(FULL, c01), t23 = __reduce_add_sync(FULL, c23);
                const uint32_t t45 = __reduce_add_sync(FULL, c45);
                const uint32_t tc[5] = {t01 & 0xFFFFu, t01 >> 16, t23 & 0xFFFFu, t23 >> 16, t45 & 0xFFFFu};
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (j == q && r >= tc[q]) { r -= tc[q]; j = q + 1; }
            } else {
                const uint32_t t0 = __reduce_add_sync(FULL, c01) & 0xFFFFu;
                if (r >= t0) { r -= t0; j = 1; }
            }
            u64 mj = m[0];
#pragma unroll
            for (int q = 1; q < 6; q++) mj = j == q ? m[q] : mj;
            int L = 0;
            warp_rank_search(__popcll(mj), r, L, lane);    // r: rank inside word L of slot j
            mj = __shfl_sync(FULL, mj, L);
            const int bit = nth_bit64(mj, r);
            col = 64 * L + bit;
            slot = class_slot_base(csel) + j;
            if (group) s0 = S_DIVAC;
            else switch (csel) {
                case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0 = 0; break;
                case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0 = 3; break;
                case C_DESTROY_TREF: s0 = j ? 7 : 4; break;
                case C_CMONO_MAKE_EMPTY: s0 = j ? 1 : 2; break;
                case C_FMONO_MAKE_DOUBLE: s0 = j ? 2 : 1; break;
                default: s0 = ((G.word(T_M1, row, L) >> bit) & 1ull) ? 1 : 2; break;
            }
        }
        const int site = row * d1 + col;
        if (p.log_mode != KMC_LOG_NONE && lane == 0) {
            const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)t;
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
                rec->flags = pick.tie ? 1 : 0; rec->total_rate = pick.total; rec->pad = 0;
            } else {
                kmc_event_compact e;
                e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
                e.flags = pick.tie ? 1 : 0; e.total_rate = pick.total;
                reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
            }
        }
        if (lane == csel) hist++;
        if (lane == slot - 7) hops++;                          // lanes 0..5: divacancy hops per direction
        if (pick.tie && lane == 0) n_ties++;

        // ---- the move -----------------------------------------------------------------------------------
        MoveNodes mv;
        int os1, os2, da2;
        {
            const MoveEntry e = MOVE_TABLE[slot];
            mv.ns0 = (int)((((uint32_t)s0 | (e.w0 & 0xF)) & ((e.w0 >> 4) & 0xF)) ^ ((e.w0 >> 8) & 0xF));
            mv.na = (int)((e.w0 >> 12) & 3);
            mv.da1 = (int)(e.w1 & 0xF) - 2;
            da2 = (int)((e.w1 >> 8) & 0xF) - 2;
            mv.a1 = wrap1(row + mv.da1, d0);
            mv.b1 = wrap1(col + (int)((e.w1 >> 4) & 0xF) - 2, d1);
            mv.a2 = wrap1(row + da2, d0);
            mv.b2 = wrap1(col + (int)((e.w1 >> 12) & 0xF) - 2, d1);
            mv.ns1 = (int)((e.w1 >> 16) & 0xF); mv.ns2 = (int)((e.w1 >> 20) & 0xF);
            os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
        }
        const int na = mv.na;
        // lanes 0..5 own a plane each: start fetching the words they will rewrite
        const int as[3] = {row, mv.a1, mv.a2}, bcol[3] = {col, mv.b1, mv.b2}, ns[3] = {mv.ns0, mv.ns1, mv.ns2};
        const int os[3] = {s0, os1, os2};
        bool mine[3];                                          // does node i leave or enter this lane's plane?
#pragma unroll
        for (int i = 0; i < 3; i++) {
            mine[i] = i < na && lane < N_TYPES && (status_plane(os[i]) == lane || status_plane(ns[i]) == lane);
            if (mine[i]) asm volatile("prefetch.global.L1 [%0];" :: "l"(P + ((size_t)lane * d0 + as[i]) * W + (bcol[i] >> 6)));
        }
        // Refresh of the five neighbourhood classes.  Every dependency of a site is inside its ring of six
        // neighbours (rows +-1, columns +-1) or, for trefoils, at (+2, 0), (0, +2), (+2, -2).  So the tasks
        // whose counts can change are: MoveDivacancy directions of the rows within 1 of a touched row, and the
        // trefoil shapes of the touched rows and the rows 2 above them -- in the words that hold a column
        // within 2 of a touched column (at most two words).  Counted before and after the move; the
        // difference goes to the row table.
        u64 wmask = 0;
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (i < na) {
                wmask |= 1ull << (wrap1(bcol[i] - 2, d1) >> 6);
                wmask |= 1ull << (bcol[i] >> 6);
                wmask |= 1ull << (wrap1(bcol[i] + 2, d1) >> 6);
            }
        const int nw = __popcll(wmask);
        const int wA = __ffsll((long long)wmask) - 1, wB = 63 - __clzll((long long)wmask);
        // rows of the direction tasks: lo2 .. lo2 + n2 - 1 (relative to `row`); trefoil rows: g3 (nibbles, +3)
        int lo2, n2, n3;
        uint32_t g3;
        if (na == 1) { lo2 = -1; n2 = 3; n3 = 2; g3 = 0x13u; }                       // {0, -2}
        else if (na == 2) {
            lo2 = (mv.da1 < 0 ? mv.da1 : 0) - 1; n2 = (mv.da1 ? 4 : 3);
            n3 = mv.da1 ? 4 : 2;
            g3 = 0x13u | ((uint32_t)(mv.da1 + 3) << 8) | ((uint32_t)(mv.da1 + 1) << 12);   // {0, -2, da1, da1 - 2}
        } else { lo2 = -1; n2 = 5; n3 = 3; g3 = 0x531u; }                            // {-2, 0, 2}
        if (n2 > d0) n2 = d0;
        const int ntask = (6 * n2 + n3) * nw;
        uint32_t b01[3] = {0, 0, 0}, b23[3] = {0, 0, 0}, b45[3] = {0, 0, 0};
        int ta[3] = {0, 0, 0};                                 // row | word << 16 | k << 24 of the lane's tasks
        uint32_t td[3] = {0, 0, 0};
#pragma unroll
        for (int it = 0; it < 3; it++) {
            const int tk = lane + 32 * it;
            if (32 * it < ntask) {                             // warp-uniform
                const int u = nw == 2 ? tk >> 1 : tk;
                const int wd = (nw == 2 && (tk & 1)) ? wB : wA;
                int k = 6, a;
                if (u < 6 * n2) { const int j = (u * 43) >> 8; k = u - 6 * j; a = row + lo2 + j; }       // u / 6 for u < 48
                else a = row + (int)((g3 >> (4 * ((u - 6 * n2) & 7))) & 0xF) - 3;
                a = wrap1(wrap1(a, d0), d0);
                ta[it] = a | (wd << 16) | (k << 24);
                td[it] = __shfl_sync(FULL, my_desc, k);
                if (tk < ntask) task_counts(P, d0, W, a, wd, k, td[it], b01[it], b23[it], b45[it]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (mine[i]) {
                u64 *wptr = P + ((size_t)lane * d0 + as[i]) * W + (bcol[i] >> 6);
                const u64 bit = 1ull << (bcol[i] & 63);
                *wptr = (*wptr & ~bit) | (status_plane(ns[i]) == lane ? bit : 0ull);
            }
        // own-status classes (owner lanes, shared-memory row table)
        if (my_g1) {
            int d = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
            if (d) { rc16[lane * d0p + row] = (uint16_t)(rc16[lane * d0p + row] + d); bs[lane * 32 + (row >> bshift)] += d; }
            int dsum = d;
            if (na > 1) {
                d = (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
                if (d) { rc16[lane * d0p + mv.a1] = (uint16_t)(rc16[lane * d0p + mv.a1] + d); bs[lane * 32 + (mv.a1 >> bshift)] += d; }
                dsum += d;
            }
            if (na > 2) {
                d = (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
                if (d) { rc16[lane * d0p + mv.a2] = (uint16_t)(rc16[lane * d0p + mv.a2] + d); bs[lane * 32 + (mv.a2 >> bshift)] += d; }
                dsum += d;
            }
            tot += dsum;
        }
        __threadfence_block();
        __syncwarp();
        uint32_t x01 = 0, y01 = 0, x23 = 0, y23 = 0, x45 = 0, y45 = 0;       // after / before, summed over the lane's tasks
#pragma unroll
        for (int it = 0; it < 3; it++) {
            const int tk = lane + 32 * it;
            if (tk < ntask) {
                const int a = ta[it] & 0xFFFF, wd = (ta[it] >> 16) & 0xFF, k = ta[it] >> 24;
                uint32_t a01, a23, a45;
                task_counts(P, d0, W, a, wd, k, td[it], a01, a23, a45);
                const int dd[5] = {(int)(a01 & 0xFFFF) - (int)(b01[it] & 0xFFFF), (int)(a01 >> 16) - (int)(b01[it] >> 16),
                                   (int)(a23 & 0xFFFF) - (int)(b23[it] & 0xFFFF), (int)(a23 >> 16) - (int)(b23[it] >> 16),
                                   (int)a45 - (int)b45[it]};
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (dd[q]) {
                        atomicAdd(&rc32[((C_MOVE_DIV + q) * d0p + a) >> 1], (uint32_t)dd[q] << (16 * (a & 1)));
                        atomicAdd(&bs[(C_MOVE_DIV + q) * 32 + (a >> bshift)], (uint32_t)dd[q]);
                    }
                x01 += a01; y01 += b01[it]; x23 += a23; y23 += b23[it]; x45 += a45; y45 += b45[it];
            }
        }
        {
            int d2, d3, d4, d5, d6, unused;
            reduce_packed_diff(x01, y01, d2, d3);
            reduce_packed_diff(x23, y23, d4, d5);
            reduce_packed_diff(x45, y45, d6, unused);
            tot += lane == 2 ? d2 : lane == 3 ? d3 : lane == 4 ? d4 : lane == 5 ? d5 : lane == 6 ? d6 : 0;
        }
        __syncwarp();
    }

    const long long done = t;
    for (int i = lane; i < (NC * d0p) / 2; i += 32) reinterpret_cast<uint32_t *>(ghier)[i] = rc32[i];
    if (lane < NC) {
        int fresh = 0, blocks = 0;
        for (int a = 0; a < d0; a++) fresh += rc16[lane * d0p + a];
        for (int i = 0; i < 32; i++) blocks += (int)bs[lane * 32 + i];
        if (fresh != tot || blocks != tot) flag |= FLAG_VALIDATE;      // totals and block sums vs the row table
    }
    flag = __reduce_or_sync(FULL, flag);
    if (lane < NC) p.hist[(size_t)rep * NC + lane] += hist;
    if (lane < 6) p.hops[(size_t)rep * 6 + lane] += hops;
    if (lane == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (p.clock) p.clock[rep] = __dadd_rn(p.clock[rep], tclock);
        p.ties[rep] += n_ties;
        if (flag) p.flags[rep] |= flag;
    }
}

}  // namespace kmc
