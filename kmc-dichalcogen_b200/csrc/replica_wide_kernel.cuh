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
    return (NC * d0p * 2 + 15) & ~(size_t)15;
}

struct WGeom {
    const u64 *P; int d0, W;
    __device__ __forceinline__ u64 word(int pl, int a, int w) const { return P[((size_t)pl * d0 + a) * W + w]; }
};

// plane words of rows a-1, a, a+1 (Z and D) and a+2 (D), words w-1, w, w+1 (periodic)
struct Nbhd { u64 z[3][3], d[4][3]; };
__device__ __forceinline__ void load_nbhd(const WGeom &g, int a, int w, Nbhd &n) {
    const int wm = w == 0 ? g.W - 1 : w - 1, wp = w + 1 == g.W ? 0 : w + 1;
    const int ws[3] = {wm, w, wp};
#pragma unroll
    for (int r = 0; r < 4; r++) {
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
// position of the r-th set bit (0-based) of a 64-bit word, r < popc(m)
__device__ __forceinline__ int nth_bit64(u64 m, uint32_t r) {
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    int base = 0;
    const uint32_t cl = __popc(lo);
    uint32_t x = lo;
    if (r >= cl) { r -= cl; x = hi; base = 32; }
    return base + (int)__fns(x, 0, (int)r + 1);
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
    u64 *P = wp.planes + (size_t)rep * N_TYPES * d0 * W;
    WGeom G{P, d0, W};
    uint16_t *ghier = wp.hier + (size_t)rep * NC * d0p;
    for (int i = lane; i < (NC * d0p) / 2; i += 32) rc32[i] = reinterpret_cast<const uint32_t *>(ghier)[i];
    __syncwarp();

    const int myc = lane < NC ? lane : 0;
    const double rate = (lane < NC && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NC && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    int tot = 0;
    if (lane < NC)
        for (int a = 0; a < d0; a++) tot += rc16[lane * d0p + a];
    const uint32_t my_g1 = (lane < NC && !(lane >= C_MOVE_DIV && lane <= C_CREATE_TREF)) ? g1_const(lane) : 0u;
    unsigned long long hist = 0, n_ties = 0;
    PickState pstate = pick_state_init(lane);
    double tclock = 0.0;
    const int rpl = (d0 + 31) >> 5;                     // rows per lane in the row search
    const int wpl = (W + 31) >> 5;                      // words per lane in the site search (1 for d1 <= 2048)

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
        // ---- row: each lane owns rpl consecutive rows ---------------------------------------------------
        int row;
        {
            const int r0 = lane * rpl;
            uint32_t sum = 0;
            for (int q = 0; q < rpl; q++)
                if (r0 + q < d0) sum += rc16[csel * d0p + r0 + q];
            int L = 0;
            warp_rank_search(sum, r, L, lane);
            row = L * rpl;
            for (int q = 0; q < rpl - 1; q++) {
                const uint32_t v = rc16[csel * d0p + row];
                if (r < v) break;
                r -= v; row++;
            }
        }
        // ---- slot and site inside the row: one lane per word -------------------------------------------
        int col = 0, slot = 0, s0 = 0;
        {
            // up to six slot masks of this lane's word(s); slots in class order (row, slot, column)
            u64 m[6] = {0, 0, 0, 0, 0, 0};
            const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 : 0);
            int pa = 0, pb = -1;
            if (group == 0) {
                switch (csel) {
                    case C_CREATE_DIV: pa = T_Z; break;
                    case C_FILL_DIV: pa = T_D; break;
                    case C_DESTROY_TREF: pa = T_A4; pb = T_A7; break;
                    case C_CMONO_FROM_DOUBLE: pa = T_Z; pb = T_Z; break;
                    case C_CMONO_MAKE_EMPTY: pa = T_M2; pb = T_M1; break;
                    case C_FMONO_MAKE_DOUBLE: pa = T_M1; pb = T_M2; break;
                    case C_FMONO_FROM_EMPTY: pa = T_D; pb = T_D; break;
                    default: pa = T_M1; break;
                }
            }
            uint32_t c01 = 0, c23 = 0, c45 = 0;            // per-slot counts of this lane, 16-bit pairs
            for (int i = 0; i < wpl; i++) {
                const int wd = lane * wpl + i;
                if (wd < W) {
                    if (group == 1) {
                        Nbhd nb;
                        load_nbhd(G, row, wd, nb);
#pragma unroll
                        for (int k = 0; k < 6; k++) {
                            u64 pr, ne, fa;
                            cell_dir_masks(nb, k, pr, ne, fa);
                            const u64 km = kind_select(pr, ne, fa, csel - C_MOVE_DIV);
                            if (wpl == 1) m[k] = km;
                            const uint32_t c = __popcll(km);
                            if (k < 2) c01 += c << (16 * k); else if (k < 4) c23 += c << (16 * (k - 2)); else c45 += c << (16 * (k - 4));
                        }
                    } else if (group == 2) {
                        Nbhd nb;
                        load_nbhd(G, row, wd, nb);
                        u64 up, dn;
                        cell_trefoil_masks(nb, up, dn);
                        if (wpl == 1) { m[0] = up; m[1] = dn; }
                        c01 += __popcll(up) | (__popcll(dn) << 16);
                    } else {
                        u64 ma = G.word(pa, row, wd);
                        if (csel == C_FLIP) ma |= G.word(T_M2, row, wd);
                        const u64 mb = pb >= 0 ? G.word(pb, row, wd) : 0ull;
                        if (wpl == 1) { m[0] = ma; m[1] = mb; }
                        c01 += __popcll(ma) | (__popcll(mb) << 16);
                    }
                }
            }
            const uint32_t t01 = __reduce_add_sync(FULL, c01), t23 = __reduce_add_sync(FULL, c23);
            const uint32_t t45 = __reduce_add_sync(FULL, c45);
            int j = 0;
            {
                const uint32_t tc[6] = {t01 & 0xFFFFu, t01 >> 16, t23 & 0xFFFFu, t23 >> 16, t45 & 0xFFFFu, t45 >> 16};
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (j == q && r >= tc[q]) { r -= tc[q]; j = q + 1; }
            }
            const uint32_t cj = j < 2 ? (c01 >> (16 * j)) & 0xFFFFu : j < 4 ? (c23 >> (16 * (j - 2))) & 0xFFFFu
                                                                           : (c45 >> (16 * (j - 4))) & 0xFFFFu;
            int L = 0;
            warp_rank_search(cj, r, L, lane);              // r: rank inside lane L's word(s) of slot j
            // the owner lane's mask of slot j (recomputed uniformly when a lane holds several words)
            u64 mj = m[0];
#pragma unroll
            for (int q = 1; q < 6; q++) mj = j == q ? m[q] : mj;
            int wd = L * wpl;
            if (wpl == 1) {
                mj = __shfl_sync(FULL, mj, L);
            } else {
                for (int i = 0; i < wpl; i++, wd++) {
                    u64 x;
                    if (group == 0) {
                        x = j == 0 ? G.word(pa, row, wd) | (csel == C_FLIP ? G.word(T_M2, row, wd) : 0ull) : G.word(pb, row, wd);
                    } else {
                        Nbhd nb;
                        load_nbhd(G, row, wd, nb);
                        if (group == 1) { u64 pr, ne, fa; cell_dir_masks(nb, j, pr, ne, fa); x = kind_select(pr, ne, fa, csel - C_MOVE_DIV); }
                        else { u64 up, dn; cell_trefoil_masks(nb, up, dn); x = j ? dn : up; }
                    }
                    const uint32_t c = __popcll(x);
                    if (r < c) { mj = x; break; }
                    r -= c;
                }
            }
            col = 64 * wd + nth_bit64(mj, r);
            slot = class_slot_base(csel) + j;
            if (group) s0 = S_DIVAC;
            else switch (csel) {
                case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0 = 0; break;
                case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0 = 3; break;
                case C_DESTROY_TREF: s0 = j ? 7 : 4; break;
                case C_CMONO_MAKE_EMPTY: s0 = j ? 1 : 2; break;
                case C_FMONO_MAKE_DOUBLE: s0 = j ? 2 : 1; break;
                default: s0 = ((G.word(T_M1, row, col >> 6) >> (col & 63)) & 1ull) ? 1 : 2; break;
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
        if (pick.tie && lane == 0) n_ties++;

        // ---- the move -----------------------------------------------------------------------------------
        MoveNodes mv;
        int os1, os2;
        {
            const MoveEntry e = MOVE_TABLE[slot];
            mv.ns0 = (int)((((uint32_t)s0 | (e.w0 & 0xF)) & ((e.w0 >> 4) & 0xF)) ^ ((e.w0 >> 8) & 0xF));
            mv.na = (int)((e.w0 >> 12) & 3);
            mv.da1 = (int)(e.w1 & 0xF) - 2;
            mv.a1 = wrap1(row + mv.da1, d0);
            mv.b1 = wrap1(col + (int)((e.w1 >> 4) & 0xF) - 2, d1);
            mv.a2 = wrap1(row + (int)((e.w1 >> 8) & 0xF) - 2, d0);
            mv.b2 = wrap1(col + (int)((e.w1 >> 12) & 0xF) - 2, d1);
            mv.ns1 = (int)((e.w1 >> 16) & 0xF); mv.ns2 = (int)((e.w1 >> 20) & 0xF);
            os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
        }
        const int na = mv.na;
        // cells whose neighbourhood-class counts can change: rows row-3 .. row+3, words holding a column
        // within 2 of a touched column.  One lane per cell; ncell <= 7 * (number of such words).
        u64 wmask = 0;
        {
            const int cols[3] = {col, mv.b1, mv.b2};
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < na) {
                    wmask |= 1ull << (wrap1(cols[i] - 2, d1) >> 6);
                    wmask |= 1ull << (cols[i] >> 6);
                    wmask |= 1ull << (wrap1(cols[i] + 2, d1) >> 6);
                }
        }
        const int nw = __popcll(wmask);
        const int nrow_cells = d0 < 7 ? d0 : 7;
        const int ncell = nrow_cells * nw;
        uint32_t b01[2] = {0, 0}, b23[2] = {0, 0}, btr[2] = {0, 0};
        int ca[2] = {0, 0};
        for (int pass = 0, c = lane; pass < 2; pass++, c += 32) {
            if (c < ncell) {
                const int j = c / nw, iw = c - j * nw;
                ca[pass] = wrap1(wrap1(row - 3 + j, d0), d0);
                const int wd = nth_bit64(wmask, iw);
                Nbhd nb;
                load_nbhd(G, ca[pass], wd, nb);
                cell_counts(nb, b01[pass], b23[pass], btr[pass]);
            }
        }
        __syncwarp();
        // apply: lanes 0..5 own a plane each
        if (lane < N_TYPES) {
            const int as[3] = {row, mv.a1, mv.a2}, bs[3] = {col, mv.b1, mv.b2}, ns[3] = {mv.ns0, mv.ns1, mv.ns2};
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < na) {
                    u64 *wptr = P + ((size_t)lane * d0 + as[i]) * W + (bs[i] >> 6);
                    const u64 bit = 1ull << (bs[i] & 63);
                    *wptr = (*wptr & ~bit) | (status_plane(ns[i]) == lane ? bit : 0ull);
                }
        }
        // own-status classes (owner lanes, shared-memory row table)
        if (my_g1) {
            int d = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
            if (d) rc16[lane * d0p + row] = (uint16_t)(rc16[lane * d0p + row] + d);
            int dsum = d;
            if (na > 1) {
                d = (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
                if (d) rc16[lane * d0p + mv.a1] = (uint16_t)(rc16[lane * d0p + mv.a1] + d);
                dsum += d;
            }
            if (na > 2) {
                d = (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
                if (d) rc16[lane * d0p + mv.a2] = (uint16_t)(rc16[lane * d0p + mv.a2] + d);
                dsum += d;
            }
            tot += dsum;
        }
        __threadfence_block();
        __syncwarp();
        // after: same cells; the difference goes to the row table (packed pairs of rows) and the totals
        int dk[5] = {0, 0, 0, 0, 0};
        for (int pass = 0, c = lane; pass < 2; pass++, c += 32) {
            if (c < ncell) {
                const int j = c / nw, iw = c - j * nw;
                const int wd = nth_bit64(wmask, iw);
                Nbhd nb;
                load_nbhd(G, ca[pass], wd, nb);
                uint32_t a01, a23, atr;
                cell_counts(nb, a01, a23, atr);
                const int dd[5] = {(int)(a01 & 0xFFFF) - (int)(b01[pass] & 0xFFFF), (int)(a01 >> 16) - (int)(b01[pass] >> 16),
                                   (int)(a23 & 0xFFFF) - (int)(b23[pass] & 0xFFFF), (int)(a23 >> 16) - (int)(b23[pass] >> 16),
                                   (int)atr - (int)btr[pass]};
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (dd[q]) {
                        const int a = ca[pass];
                        atomicAdd(&rc32[((C_MOVE_DIV + q) * d0p + a) >> 1], (uint32_t)dd[q] << (16 * (a & 1)));
                        dk[q] += dd[q];
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int s = (int)__reduce_add_sync(FULL, (uint32_t)dk[q]);
            if (lane == C_MOVE_DIV + q) tot += s;
        }
        __syncwarp();
    }

    const long long done = t;
    for (int i = lane; i < (NC * d0p) / 2; i += 32) reinterpret_cast<uint32_t *>(ghier)[i] = rc32[i];
    if (lane < NC) {
        int fresh = 0;
        for (int a = 0; a < d0; a++) fresh += rc16[lane * d0p + a];
        if (__any_sync(__activemask(), fresh != tot)) flag |= FLAG_VALIDATE;      // totals vs row table
    }
    flag = __reduce_or_sync(FULL, flag);
    if (lane < NC) p.hist[(size_t)rep * NC + lane] += hist;
    if (lane == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (p.clock) p.clock[rep] = __dadd_rn(p.clock[rep], tclock);
        p.ties[rep] += n_ties;
        if (flag) p.flags[rep] |= flag;
    }
}

}  // namespace kmc
