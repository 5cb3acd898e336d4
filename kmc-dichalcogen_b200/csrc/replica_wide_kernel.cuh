// replica_wide_kernel.cuh -- K3 + K4 for lattices with rows wider than 64 sites that are too large, or too
// many, for shared memory (e.g. 1024 x 1024, BASELINE config 5): the bit-sliced algorithm of
// replica_bp_kernel.cuh with the six predicate planes held in HBM/L2 and only the class-major row-count table
// in shared memory.
//
// Plane rows carry a two-column halo on each side so that neighbour columns never need wrap logic: a row is a
// bit array of d1 + 4 positions, position q holding column (q - 2) mod d1 (positions 0, 1 repeat the last two
// columns, positions d1 + 2, d1 + 3 the first two).  The window of anchor word w (columns 64w .. 64w + 63)
// shifted by db in [-2, 2] columns starts at position 64w + db + 2, i.e. always inside words w and w + 1:
//     window = (X[w] >> (db + 2)) | (X[w + 1] << (62 - db))
// Rows have Wa = ceil(d1 / 64) anchor words and We = Wa + 1 stored words; d1 need not be a multiple of 64 (the
// last anchor word is masked).  A site update writes its main position and, within 2 columns of a row end, its
// halo copy.
//
// One warp owns one replica.  What keeps an event cheap on a 1 Mi-site lattice is that everything is local
// to a handful of 64-site words:
//   * site search inside the chosen row: one lane per anchor word builds the six MoveDivacancy direction masks
//     of its word from the plane words of rows a-1 .. a+1 (one round of independent loads), per-direction
//     totals come from three packed warp reductions, one warp scan finds the word, a popcount bisection the bit;
//   * refresh: only tasks (row, word, direction | trefoil) within 1 row (trefoils: the touched rows and 2
//     above) and 2 columns of the touched nodes can change; one lane per task re-derives its counts before and
//     after the move and adds the difference to the row table (packed 16-bit shared-memory atomics) --
//     independent of the lattice size.
// Pack / unpack (status bytes <-> planes) and the initial row table are separate grid-wide kernels, run
// only when the lattice was replaced; planes and table persist between launches.
#pragma once
#include "replica_bp_kernel.cuh"

namespace kmc {

struct WideParams {
    ReplicaParams base;
    u64 *planes;                     // [R][N_TYPES][d0][We]
    uint16_t *hier;                  // [R][NC][d0p] row counts, persists between launches
    int Wa, d0p;
};

__host__ __device__ inline int wide_anchor_words(int d1) { return (d1 + 63) >> 6; }
__host__ __device__ inline int wide_row_words(int d1) { return wide_anchor_words(d1) + 1; }
__host__ __device__ inline size_t wide_warp_bytes(int d0) {
    const size_t d0p = ((size_t)d0 + 1) & ~(size_t)1;
    // row table + 32 block sums per class + the before-counts of up to 4 x 32 refresh tasks (5 words each)
    return ((NC * d0p * 2 + 15) & ~(size_t)15) + NC * 32 * sizeof(uint32_t) + 4 * 5 * 32 * sizeof(uint32_t);
}

struct WGeom {
    const u64 *P; int d0, d1, Wa, We;
    // (a replica's planes hold fewer than 2^32 words, so the offset is taken in 32 bits)
    __device__ __forceinline__ const u64 *rowp(int pl, int a) const { return P + (uint32_t)((pl * d0 + a) * We); }
    // anchor columns of word w that exist (all ones except in a ragged last word)
    __device__ __forceinline__ u64 valid(int w) const {
        const int left = d1 - 64 * w;
        return left >= 64 ? ~0ull : ((1ull << left) - 1ull);
    }
};
// the 64 columns 64w + db .. 64w + db + 63 of a plane row (halo layout), |db| <= 2
__device__ __forceinline__ u64 window(u64 x0, u64 x1, int db) {
    const uint32_t o = (uint32_t)(db + 2);                 // 0 .. 4: only the low half of x1 shifts in
    const uint32_t a = (uint32_t)x0, b = (uint32_t)(x0 >> 32), c = (uint32_t)x1;
    return (u64)__funnelshift_r(a, b, o) | ((u64)__funnelshift_r(b, c, o) << 32);
}
__device__ __forceinline__ u64 window_at(const u64 *rowp, int w, int db) { return window(rowp[w], rowp[w + 1], db); }

// words w, w + 1 of planes Z and D in rows a-1, a, a+1 and of plane D in row a+2
struct Nbhd { u64 z[3][2], d[4][2]; u64 valid; };
template <bool WITH_ROW2 = true>
__device__ __forceinline__ void load_nbhd(const WGeom &g, int a, int w, Nbhd &n) {
    const uint32_t plane = (uint32_t)(g.d0 * g.We);          // words per plane: Z is plane 0, D plane 1
#pragma unroll
    for (int r = 0; r < (WITH_ROW2 ? 4 : 3); r++) {
        const int ar = wrap1(a + r - 1, g.d0);                 // offsets -1 .. +2 wrap at most once (d0 >= 5)
        const u64 *zp = g.P + (uint32_t)(ar * g.We + w);
        n.d[r][0] = zp[plane]; n.d[r][1] = zp[plane + 1];
        if (r < 3) { n.z[r][0] = zp[0]; n.z[r][1] = zp[1]; }
    }
    n.valid = g.valid(w);
}
// MoveDivacancy direction k of the cell: present / near / far masks
__device__ __forceinline__ void cell_dir_masks(const Nbhd &n, int k, u64 &present, u64 &near, u64 &far) {
    const int rn = nib(NBR_DA, k) + 1, re = nib(NEAR_DA, k) + 1, rf = nib(FAR_DA, k) + 1;
    present = window(n.d[1][0], n.d[1][1], 0) & n.valid & window(n.z[rn][0], n.z[rn][1], nib(NBR_DB, k));
    near = window(n.d[re][0], n.d[re][1], nib(NEAR_DB, k));
    far = window(n.d[rf][0], n.d[rf][1], nib(FAR_DB, k));
}
__device__ __forceinline__ void cell_trefoil_masks(const Nbhd &n, u64 &up, u64 &dn) {
    const u64 both = window(n.d[1][0], n.d[1][1], 0) & n.valid & window(n.d[3][0], n.d[3][1], 0);
    up = both & window(n.d[1][0], n.d[1][1], 2);
    dn = both & window(n.d[3][0], n.d[3][1], -2);
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
// status bytes -> planes: one thread per (replica, row, stored word)
__global__ void kmc_wide_pack_kernel(const uint8_t *status, u64 *planes, int R, int d0, int d1)
{
    const int We = wide_row_words(d1);
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0 * We) return;
    const int w = (int)(g % We);
    const int a = (int)((g / We) % d0);
    const long long rep = g / ((long long)We * d0);
    const uint8_t *src = status + (rep * d0 + a) * (long long)d1;
    u64 out[N_TYPES] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 64; i++) {
        const int q = 64 * w + i;
        if (q >= d1 + 4) break;
        int c = q - 2;
        c = c < 0 ? c + d1 : (c >= d1 ? c - d1 : c);
        const int pl = status_plane(src[c] & 0xF);
#pragma unroll
        for (int t = 0; t < N_TYPES; t++) out[t] |= (u64)(pl == t) << i;
    }
#pragma unroll
    for (int t = 0; t < N_TYPES; t++) planes[((rep * N_TYPES + t) * d0 + a) * We + w] = out[t];
}

// planes -> status bytes: one thread per (replica, row, anchor word); trefoil members are recognised from the
// anchor planes of the rows / columns they sit relative to
__global__ void kmc_wide_unpack_kernel(const u64 *planes, uint8_t *status, int R, int d0, int d1)
{
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0 * Wa) return;
    const int w = (int)(g % Wa);
    const int a = (int)((g / Wa) % d0);
    const long long rep = g / ((long long)Wa * d0);
    WGeom G{planes + rep * N_TYPES * d0 * We, d0, d1, Wa, We};
    const int am2 = wrap1(a - 2, d0);
    const u64 z = window_at(G.rowp(T_Z, a), w, 0), dv = window_at(G.rowp(T_D, a), w, 0);
    const u64 m1 = window_at(G.rowp(T_M1, a), w, 0), m2 = window_at(G.rowp(T_M2, a), w, 0);
    const u64 a4 = window_at(G.rowp(T_A4, a), w, 0), a7 = window_at(G.rowp(T_A7, a), w, 0);
    // members: Up anchor q=(x,y): (x+2,y) role 1 -> code 5, (x,y+2) role 2 -> code 6;
    //          Down anchor:        (x+2,y) role 1 -> code 8, (x+2,y-2) role 2 -> code 9
    const u64 c5 = window_at(G.rowp(T_A4, am2), w, 0);
    const u64 c6 = window_at(G.rowp(T_A4, a), w, -2);                    // anchor two columns to the left
    const u64 c8 = window_at(G.rowp(T_A7, am2), w, 0);
    uint8_t *dst = status + (rep * d0 + a) * (long long)d1 + 64 * w;
    const int ncols = min(64, d1 - 64 * w);
    const bool words = (d1 & 3) == 0;                                     // row starts are 4-byte aligned
    for (int b4 = 0; b4 < 16; b4++) {
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = 4 * b4 + i;
            const uint32_t s = ((z >> b) & 1) ? 0u : ((dv >> b) & 1) ? 3u : ((m1 >> b) & 1) ? 1u : ((m2 >> b) & 1) ? 2u
                             : ((a4 >> b) & 1) ? 4u : ((a7 >> b) & 1) ? 7u : ((c5 >> b) & 1) ? 5u
                             : ((c6 >> b) & 1) ? 6u : ((c8 >> b) & 1) ? 8u : 9u;
            word |= s << (8 * i);
        }
        if (words && 4 * b4 + 4 <= ncols) *reinterpret_cast<uint32_t *>(dst + 4 * b4) = word;
        else
            for (int i = 0; i < 4; i++)
                if (4 * b4 + i < ncols) dst[4 * b4 + i] = (uint8_t)(word >> (8 * i));
    }
}

// row table from the planes: one thread per (replica, row)
// nct = tables per replica in `hier` (13; 16 when the MoveMonovacancy tables follow, see replica_wide2_kernel.cuh)
__global__ void kmc_wide_hier_kernel(const u64 *planes, uint16_t *hier, int R, int d0, int d1, int d0p, int nct = NC)
{
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0) return;
    const int a = (int)(g % d0);
    const long long rep = g / d0;
    WGeom G{planes + rep * N_TYPES * d0 * We, d0, d1, Wa, We};
    uint32_t pz = 0, pd = 0, pm = 0, pa = 0, k01 = 0, k23 = 0, tre = 0;
    for (int w = 0; w < Wa; w++) {
        const u64 v = G.valid(w);
        pz += __popcll(window_at(G.rowp(T_Z, a), w, 0) & v); pd += __popcll(window_at(G.rowp(T_D, a), w, 0) & v);
        pm += __popcll((window_at(G.rowp(T_M1, a), w, 0) | window_at(G.rowp(T_M2, a), w, 0)) & v);
        pa += __popcll((window_at(G.rowp(T_A4, a), w, 0) | window_at(G.rowp(T_A7, a), w, 0)) & v);
        Nbhd n;
        load_nbhd(G, a, w, n);
        uint32_t c01, c23, ct;
        cell_counts(n, c01, c23, ct);
        k01 += c01; k23 += c23; tre += ct;
    }
    uint16_t *H = hier + rep * nct * d0p;
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
// the operand's window for anchor word w of anchor row a: two 64-bit loads and a funnel shift by db + 2
__device__ __forceinline__ u64 operand_word(const u64 *P, int d0, int We, int a, int w, uint32_t od) {
    const int pl = (int)(od & 1u), da = (int)((od >> 1) & 7u) - 2, o = (int)((od >> 4) & 7u);
    int ar = a + da;
    ar = ar < 0 ? ar + d0 : (ar >= d0 ? ar - d0 : ar);
    const u64 *rp = P + (uint32_t)(pl * d0 + ar) * (uint32_t)We + w;
    const uint2 x0 = *reinterpret_cast<const uint2 *>(rp);
    const uint32_t x1lo = *reinterpret_cast<const uint32_t *>(rp + 1);      // o <= 4: only the low half shifts in
    return (u64)__funnelshift_r(x0.x, x0.y, o) | ((u64)__funnelshift_r(x0.y, x1lo, o) << 32);
}
// packed counts of a task: classes (2,3) in c01, (4,5) in c23 for k < 6; class 6 in c45 for k == 6
__device__ __forceinline__ void task_counts(const u64 *P, int d0, int We, int a, int w, int k, uint32_t desc, u64 valid,
                                            uint32_t &c01, uint32_t &c23, uint32_t &c45) {
    const u64 *dp = P + (uint32_t)(T_D * d0 + a) * (uint32_t)We + w;
    const u64 D0 = window(dp[0], dp[1], 0) & valid;
    const u64 A = operand_word(P, d0, We, a, w, desc & 0x7Fu);
    const u64 Bv = operand_word(P, d0, We, a, w, (desc >> 7) & 0x7Fu);
    const u64 C = operand_word(P, d0, We, a, w, (desc >> 14) & 0x7Fu);
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
    const int d0 = p.d0, d1 = p.d1, Wa = wp.Wa, We = wp.Wa + 1, d0p = wp.d0p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * (blockDim.x >> 5) + warp;
    if (rep >= p.n_replicas) return;

    extern __shared__ uint4 smem_raw[];
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)warp * wide_warp_bytes(d0));
    uint32_t *rc32 = reinterpret_cast<uint32_t *>(rc16);
    uint32_t *bs = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(rc16) + ((NC * (size_t)d0p * 2 + 15) & ~(size_t)15));
    uint32_t *scr = bs + NC * 32;                          // [4 rounds][5 fields][32 lanes]
    u64 *P = wp.planes + (size_t)rep * N_TYPES * d0 * We;
    WGeom G{P, d0, d1, Wa, We};
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
    double tclock = p.clock ? p.clock[rep] : 0.0;          // continues across launches in event order
    uint32_t *const tracer = p.tracer ? p.tracer + (size_t)rep * p.n : nullptr;
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
        if (p.clock) tclock = __dadd_rn(tclock, kmc_time_step(p.clock_mode, pick.total, p.seed, p.replica0 + (uint32_t)rep,
                                                            step_base + (unsigned long long)t));
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) +
                                  ((size_t)rep * (size_t)p.nsteps + (size_t)t);
            if (lane < NCA) rec->counts[lane] = lane < NC ? (uint32_t)tot : 0u;
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
            const int lines = (We * 8 + 127) >> 7;             // 128-byte lines per plane row (<= 3)
            const int j = lane >> 2, pl = (lane >> 1) & 1, li = lane & 1;
            if (j < 7 && li < lines) {
                const int a = wrap1(wrap1(row - 3 + j, d0), d0);
                asm volatile("prefetch.global.L1 [%0];" :: "l"(P + ((size_t)pl * d0 + a) * We + min(li * 16, We - 1)));
            }
        }
        // ---- slot and site inside the row: one lane per anchor word (Wa <= 32), slots in class order ---------
        int col, slot, s0;
        {
            u64 m[6] = {0, 0, 0, 0, 0, 0};                 // this lane's word: mask of the class per slot
            const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 : 0);
            if (lane < Wa) {
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
                    const u64 *ra = G.rowp(T_D, row), *rb2 = G.rowp(T_D, wrap1(row + 2, d0));
                    const u64 a0 = ra[lane], a1 = ra[lane + 1], b0 = rb2[lane], b1 = rb2[lane + 1];
                    const u64 both = window(a0, a1, 0) & window(b0, b1, 0) & G.valid(lane);
                    m[0] = both & window(a0, a1, 2);
                    m[1] = both & window(b0, b1, -2);
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
                    const u64 v = G.valid(lane);
                    m[0] = window_at(G.rowp(pa, row), lane, 0) & v;
                    if (csel == C_FLIP) m[0] |= window_at(G.rowp(T_M2, row), lane, 0) & v;
                    if (pb >= 0) m[1] = window_at(G.rowp(pb, row), lane, 0) & v;
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
                default: s0 = ((window_at(G.rowp(T_M1, row), L, 0) >> bit) & 1ull) ? 1 : 2; break;
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
        int os1, os2, da2, db1, db2;
        {
            const MoveEntry e = MOVE_TABLE[slot];
            mv.ns0 = (int)((((uint32_t)s0 | (e.w0 & 0xF)) & ((e.w0 >> 4) & 0xF)) ^ ((e.w0 >> 8) & 0xF));
            mv.na = (int)((e.w0 >> 12) & 3);
            mv.da1 = (int)(e.w1 & 0xF) - 2;
            da2 = (int)((e.w1 >> 8) & 0xF) - 2;
            mv.a1 = wrap1(row + mv.da1, d0);
            db1 = (int)((e.w1 >> 4) & 0xF) - 2;
            db2 = (int)((e.w1 >> 12) & 0xF) - 2;
            mv.b1 = wrap1(col + db1, d1);
            mv.a2 = wrap1(row + da2, d0);
            mv.b2 = wrap1(col + db2, d1);
            mv.ns1 = (int)((e.w1 >> 16) & 0xF); mv.ns2 = (int)((e.w1 >> 20) & 0xF);
            os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
        }
        const int na = mv.na;
        if (tracer && lane == 31)
            tracer_update(tracer, slot, na, row * d1 + col, mv.ns0, mv.a1 * d1 + mv.b1, mv.ns1, mv.a2 * d1 + mv.b2, mv.ns2);
        // lanes 0..5 own a plane each (the Z and D rows they rewrite were prefetched with the refresh rows)
        const int as[3] = {row, mv.a1, mv.a2}, bcol[3] = {col, mv.b1, mv.b2}, ns[3] = {mv.ns0, mv.ns1, mv.ns2};
        const int os[3] = {s0, os1, os2};
        bool mine[3];                                          // does node i leave or enter this lane's plane?
#pragma unroll
        for (int i = 0; i < 3; i++) {
            mine[i] = i < na && lane < N_TYPES && (status_plane(os[i]) == lane || status_plane(ns[i]) == lane);
        }
        // Refresh of the five neighbourhood classes.  Every dependency of a site is inside its ring of six
        // neighbours (rows +-1, columns +-1) or, for trefoils, at (+2, 0), (0, +2), (+2, -2).  So the tasks
        // whose counts can change are: MoveDivacancy directions of the rows within 1 of a touched row, and the
        // trefoil shapes of the touched rows and the rows 2 above them -- in the words that hold a column
        // within 2 of a touched column (at most two words).  Counted before and after the move; the
        // difference goes to the row table.
        // words holding a column within 2 of a touched column: the touched columns are col + {0, db1, db2} (the
        // table holds 0 for nodes a move does not have), so the span is one unwrapped column range of at most 9
        // columns: one or two words, or -- across the row end -- the words of its two ends plus the last and the
        // first word (a superset is harmless: unchanged tasks contribute a zero difference)
        uint32_t wmask;
        {
            const int lo = col + min(0, min(db1, db2)) - 2, hi = col + max(0, max(db1, db2)) + 2;
            if (lo >= 0 && hi < d1) wmask = (1u << (lo >> 6)) | (1u << (hi >> 6));
            else wmask = (1u << ((lo < 0 ? lo + d1 : lo) >> 6)) | (1u << ((hi >= d1 ? hi - d1 : hi) >> 6)) | 1u | (1u << (Wa - 1));
        }
        const int nw = __popc(wmask);
        // at most three words: two in the interior, three when the span wraps over a ragged last word
        const int wA = __ffs((int)wmask) - 1, wB = 31 - __clz((int)wmask);
        const int wM = nw == 3 ? __ffs((int)(wmask & (wmask - 1))) - 1 : wA;
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
        // one task per lane and round (a single round unless the span crosses a word boundary); the task and its
        // before-counts wait in shared memory for the after-pass, so the task code exists once
#pragma unroll 1
        for (int it = 0; 32 * it < ntask; it++) {
            const int tk = lane + 32 * it;
            const int u = nw == 1 ? tk : (nw == 2 ? tk >> 1 : (tk * 171) >> 9);      // tk / nw
            const int iw = tk - u * nw;
            const int wd = iw == 0 ? wA : (iw == nw - 1 ? wB : wM);
            int k = 6, a;
            if (u < 6 * n2) { const int j = (u * 43) >> 8; k = u - 6 * j; a = row + lo2 + j; }       // u / 6 for u < 48
            else a = row + (int)((g3 >> (4 * ((u - 6 * n2) & 7))) & 0xF) - 3;
            a = wrap1(wrap1(a, d0), d0);
            const uint32_t desc = __shfl_sync(FULL, my_desc, k);
            uint32_t b01 = 0, b23 = 0, b45 = 0;
            if (tk < ntask) task_counts(P, d0, We, a, wd, k, desc, G.valid(wd), b01, b23, b45);
            uint32_t *sc = scr + it * 160 + lane;
            sc[0] = b01; sc[32] = b23; sc[64] = b45; sc[96] = (uint32_t)(a | (wd << 16) | (k << 24)); sc[128] = desc;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (mine[i]) {
                // the site's main position and, within two columns of a row end, its halo copy
                u64 *rowp = P + (uint32_t)((lane * d0 + as[i]) * We);
                const bool set = status_plane(ns[i]) == lane;
                const int q0 = bcol[i] + 2;
                const int q1 = bcol[i] < 2 ? bcol[i] + d1 + 2 : (bcol[i] >= d1 - 2 ? bcol[i] + 2 - d1 : -1);
                {
                    u64 *wptr = rowp + (q0 >> 6);
                    const u64 bit = 1ull << (q0 & 63);
                    *wptr = (*wptr & ~bit) | (set ? bit : 0ull);
                }
                if (q1 >= 0) {
                    u64 *wptr = rowp + (q1 >> 6);
                    const u64 bit = 1ull << (q1 & 63);
                    *wptr = (*wptr & ~bit) | (set ? bit : 0ull);
                }
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
#pragma unroll 1
        for (int it = 0; 32 * it < ntask; it++) {
            const int tk = lane + 32 * it;
            if (tk < ntask) {
                const uint32_t *sc = scr + it * 160 + lane;
                const uint32_t b01 = sc[0], b23 = sc[32], b45 = sc[64], tak = sc[96], desc = sc[128];
                const int a = tak & 0xFFFF, wd = (tak >> 16) & 0xFF, k = tak >> 24;
                uint32_t a01, a23, a45;
                task_counts(P, d0, We, a, wd, k, desc, G.valid(wd), a01, a23, a45);
                const int dd[5] = {(int)(a01 & 0xFFFF) - (int)(b01 & 0xFFFF), (int)(a01 >> 16) - (int)(b01 >> 16),
                                   (int)(a23 & 0xFFFF) - (int)(b23 & 0xFFFF), (int)(a23 >> 16) - (int)(b23 >> 16),
                                   (int)a45 - (int)b45};
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (dd[q]) {
                        atomicAdd(&rc32[((C_MOVE_DIV + q) * d0p + a) >> 1], (uint32_t)dd[q] << (16 * (a & 1)));
                        atomicAdd(&bs[(C_MOVE_DIV + q) * 32 + (a >> bshift)], (uint32_t)dd[q]);
                    }
                x01 += a01; y01 += b01; x23 += a23; y23 += b23; x45 += a45; y45 += b45;
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
    if (lane < NC) p.hist[(size_t)rep * NCA + lane] += hist;
    if (lane < 6) p.hops[(size_t)rep * 6 + lane] += hops;
    if (lane == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (p.clock) p.clock[rep] = tclock;
        p.ties[rep] += n_ties;
        if (flag) { p.flags[rep] |= flag; atomicOr(p.any_flag, flag); }
    }
}

}  // namespace kmc
