// noinc_planes_kernel.cuh -- the `--no-incremental` event (reference demo1/sim.py:114-143: re-enumerate everything,
// choose, perform) on the BIT-SLICED lattice, for lattices held in HBM (BASELINE config 4: 4096 x 4096).
//
// The byte-lattice version (sweep_kernel.cuh count-only + noinc_kernel.cuh) spends ~1.4 warp-instructions per site
// on the enumeration (nibble-parallel, 8 sites per operation) and selects on one CTA from a strided row table:
// 23 + 18 us per event at 4096^2.  Here the lattice lives as the six predicate planes of replica_wide_kernel.cuh
// (Z, D, M1, M2, A4, A7; 0.75 B per site, two-column halo per row end), so
//   * the full enumeration evaluates 64 sites per operation: one thread per 64-site cell builds the six
//     MoveDivacancy direction masks and the trefoil masks of its cell from a 3 x 2 neighbourhood of plane words and
//     popcounts them (~0.1 warp-instructions per site; the planes are 12.6 MB at 4096^2: HBM/L2-bound);
//   * row counts leave the kernel class-major (coalesced for the row search) together with per-CTA partial sums
//     (a two-level hierarchy: CTA -> row), no atomics, nothing to zero;
//   * the selection kernel scans the CTA sums, walks the <= 8 rows of the chosen CTA, searches the chosen row with
//     one thread per cell and flips the touched bits of the planes (and their halo copies).
// Status bytes are only materialised (kmc_wide_unpack_kernel) when somebody asks for them.
// Results are bit-identical to the byte-lattice path (tests: both against the oracle).
#pragma once
#include "noinc_kernel.cuh"
#include "replica_wide2_kernel.cuh"

namespace kmc {

constexpr int PRC_STRIDE = 16;             // words per CTA partial-sum record / classes per row-count plane

__host__ __device__ inline int planes_chunks(int d1) {       // warps per row: power of two >= ceil(Wa / 32), <= 8
    const int need = (wide_anchor_words(d1) + 31) >> 5;
    int c = 1;
    while (c < need) c <<= 1;
    return c;
}

// Full enumeration from the planes.  CTA = 8 warps = 8 / chunks rows; warp = 32 consecutive cells of one row.
//   rowc [R][16][d0]   class-major row counts (classes 0..12)
//   blk  [R][nblk][16] per-CTA sums of the same, nblk = ceil(d0 / rows_per_cta)
// MM: slots 13, 14, 15 of a record hold MoveMonovacancy natural, merge-divacancy, swap-divacancy (class 16); the
// split-divacancy count (class 15) is twice the sum of slots 2..5 and is not stored.
template <bool MM = false>
__global__ void __launch_bounds__(256) kmc_planes_rowcount_kernel(const u64 *planes, uint32_t *rowc, uint32_t *blk,
                                                                 int R, int d0, int d1, int chunks, int nblk)
{
    __shared__ uint32_t part[8][7];
    __shared__ uint32_t rowval[8][PRC_STRIDE];
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpb = 8 / chunks;
    const int rep = blockIdx.x / nblk, b = blockIdx.x - rep * nblk;
    const int rl = warp / chunks;                                  // row slot inside the CTA
    const int a = b * rpb + rl, chunk = warp % chunks;
    const int w = chunk * 32 + lane;
    WGeom G{planes + (size_t)rep * N_TYPES * d0 * We, d0, d1, Wa, We};
    uint32_t zd = 0, ma = 0, k01 = 0, k23 = 0, tre = 0, nm = 0, sw = 0;
    if (a < d0 && w < Wa) {
        if (MM)
            for (int k = 0; k < 6; k++) { uint32_t x, y; mm_cell_counts(G, a, w, k, x, y); nm += x; sw += y; }
        const u64 v = G.valid(w);
        zd = (uint32_t)__popcll(window_at(G.rowp(T_Z, a), w, 0) & v) |
             ((uint32_t)__popcll(window_at(G.rowp(T_D, a), w, 0) & v) << 16);
        ma = (uint32_t)__popcll((window_at(G.rowp(T_M1, a), w, 0) | window_at(G.rowp(T_M2, a), w, 0)) & v) |
             ((uint32_t)__popcll((window_at(G.rowp(T_A4, a), w, 0) | window_at(G.rowp(T_A7, a), w, 0)) & v) << 16);
        Nbhd n;
        load_nbhd(G, a, w, n);
        cell_counts(n, k01, k23, tre);
    }
    // 16-bit fields: a warp covers 2048 sites with <= 6 moves per site and class (<= 12288 per field)
    zd = __reduce_add_sync(FULL, zd); ma = __reduce_add_sync(FULL, ma);
    k01 = __reduce_add_sync(FULL, k01); k23 = __reduce_add_sync(FULL, k23); tre = __reduce_add_sync(FULL, tre);
    if (MM) { nm = __reduce_add_sync(FULL, nm); sw = __reduce_add_sync(FULL, sw); }
    if (lane == 0) {
        part[warp][0] = zd; part[warp][1] = ma; part[warp][2] = k01; part[warp][3] = k23; part[warp][4] = tre;
        part[warp][5] = nm; part[warp][6] = sw;
    }
    __syncthreads();
    // the leader warp of each row: lane c < 13 combines the row's chunks (fields unpacked first) and writes class c
    if (chunk == 0 && lane < PRC_STRIDE) {
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, tr = 0, nat = 0, mer = 0, swp = 0;
        for (int q = 0; q < chunks; q++) {
#pragma unroll
            for (int f = 0; f < 4; f++) { const uint32_t x = part[warp + q][f]; lo[f] += x & 0xFFFFu; hi[f] += x >> 16; }
            tr += part[warp + q][4];
            nat += part[warp + q][5] & 0xFFFFu; mer += part[warp + q][5] >> 16; swp += part[warp + q][6];
        }
        const uint32_t pz = lo[0], pd = hi[0], pm = lo[1], pa = hi[1];
        uint32_t val = 0;
        switch (lane) {
            case C_CREATE_DIV: val = pz; break;
            case C_FILL_DIV: val = pd; break;
            case C_MOVE_DIV: val = lo[2]; break;
            case C_MOVE_DIV + 1: val = hi[2]; break;
            case C_MOVE_DIV + 2: val = lo[3]; break;
            case C_MOVE_DIV + 3: val = hi[3]; break;
            case C_CREATE_TREF: val = tr; break;
            case C_DESTROY_TREF: val = pa; break;
            case C_CMONO_FROM_DOUBLE: val = 2 * pz; break;
            case C_FMONO_FROM_EMPTY: val = 2 * pd; break;
            case C_CMONO_MAKE_EMPTY: case C_FMONO_MAKE_DOUBLE: case C_FLIP: val = pm; break;   // one per monovacancy
            case 13: val = MM ? nat : 0; break;
            case 14: val = MM ? mer : 0; break;
            case 15: val = MM ? swp : 0; break;
            default: val = 0; break;
        }
        if (a >= d0) val = 0;
        if (a < d0 && lane < (MM ? PRC_STRIDE : NC)) rowc[((size_t)rep * PRC_STRIDE + lane) * d0 + a] = val;
        rowval[rl][lane] = val;
    }
    __syncthreads();
    if (warp == 0 && lane < PRC_STRIDE) {
        uint32_t s = 0;
        for (int q = 0; q < rpb; q++) s += rowval[q][lane];
        blk[((size_t)rep * nblk + b) * PRC_STRIDE + lane] = s;          // (lanes 13..15: MoveMonovacancy tables, 0 without MM)
    }
}

struct NoincPlanesParams {
    NoincParams base;                    // rate / order / draws / statistics pointers; status, rowcnt unused
    u64 *planes;                         // [R][6][d0][We]
    const uint32_t *rowc;                // [R][16][d0]
    const uint32_t *blk;                 // [R][nblk][16]
    unsigned long long *counts;          // [R][NCA] class totals of this event's enumeration (out)
    int nblk, rpb;
};

// One CTA (256 threads) per replica.
template <bool MM = false>
__global__ void __launch_bounds__(256) kmc_noinc_planes_select_kernel(const NoincPlanesParams q)
{
    constexpr int NCX = MM ? NCA : NC;
    constexpr int NL = MM ? 32 : 16;
    const NoincParams &p = q.base;
    __shared__ unsigned long long wsum[8];
    __shared__ uint32_t s_tot[8][PRC_STRIDE];               // per-warp class sums of the CTA records
    __shared__ double s_u1, s_u2;
    __shared__ unsigned long long s_rank;
    __shared__ double s_total;
    __shared__ int s_csel, s_zero, s_tie, s_row, s_word, s_bit;
    __shared__ uint32_t s_rin, s_slotcnt[12];
    const int rep = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int d0 = p.d0, d1 = p.d1;
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)p.t;
    if (p.flags[rep] & FLAG_ZERO_RATE) {                // this replica already stopped (uniform): no event
        if (tid == 0) {
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = 0xFFFFFFFFu; rec->cls = KMC_CLASS_NONE; rec->slot = KMC_SLOT_NONE;
                rec->flags = 0; rec->total_rate = 0.0; rec->pad = 0;
            } else if (p.log_mode == KMC_LOG_COMPACT) {
                kmc_event_compact e; e.site = 0xFFFFFFFFu; e.cls = KMC_CLASS_NONE; e.slot = KMC_SLOT_NONE;
                e.flags = 0; e.total_rate = 0.0;
                reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
            }
        }
        return;
    }
    u64 *P = q.planes + (size_t)rep * N_TYPES * d0 * We;
    WGeom G{P, d0, d1, Wa, We};
    const uint32_t *blk = q.blk + (size_t)rep * q.nblk * PRC_STRIDE;
    const uint32_t *rowc = q.rowc + (size_t)rep * PRC_STRIDE * d0;
    const unsigned long long step = p.steps_done[rep];

    // ---- class totals: every thread owns KB consecutive CTA records and keeps all 13 class sums of them in
    // registers (one round of independent 16-byte loads); they serve the totals now and the row search below
    constexpr int KB_MAX = 8;                               // nblk <= 2048 CTA records (d0 <= 16384 rows)
    const int kb = (q.nblk + 255) / 256;
    uint32_t mine_c[PRC_STRIDE];
#pragma unroll
    for (int c = 0; c < PRC_STRIDE; c++) mine_c[c] = 0;
    for (int i = 0; i < kb; i++) {
        const int b = tid * kb + i;
        if (b < q.nblk) {
            const uint4 *rec = reinterpret_cast<const uint4 *>(blk + (size_t)b * PRC_STRIDE);
            const uint4 x0 = rec[0], x1 = rec[1], x2 = rec[2], x3 = rec[3];
            mine_c[0] += x0.x; mine_c[1] += x0.y; mine_c[2] += x0.z; mine_c[3] += x0.w;
            mine_c[4] += x1.x; mine_c[5] += x1.y; mine_c[6] += x1.z; mine_c[7] += x1.w;
            mine_c[8] += x2.x; mine_c[9] += x2.y; mine_c[10] += x2.z; mine_c[11] += x2.w;
            mine_c[12] += x3.x; mine_c[13] += x3.y; mine_c[14] += x3.z; mine_c[15] += x3.w;
        }
    }
    (void)KB_MAX;
    {
        uint32_t mysum = 0;                                 // lane c < 13 of each warp collects the warp's sum of class c
#pragma unroll
        for (int c = 0; c < (MM ? PRC_STRIDE : NC); c++) {
            const uint32_t sc = __reduce_add_sync(FULL, mine_c[c]);
            if (lane == c) mysum = sc;
        }
        if (lane < PRC_STRIDE) s_tot[tid >> 5][lane] = lane < (MM ? PRC_STRIDE : NC) ? mysum : 0u;
    }
    // the draws do not depend on the lattice: warp 1 makes them while the sums settle
    if (tid == 32) {
        double u1, u2;
        if (p.draws) {
            const double2 u = *reinterpret_cast<const double2 *>(p.draws + ((size_t)rep * (size_t)p.nsteps + (size_t)p.t) * 2);
            u1 = u.x; u2 = u.y;
        } else {
            kmc_draw(p.seed, p.replica0 + (uint32_t)rep, step, u1, u2);
        }
        s_u1 = u1; s_u2 = u2;
    }
    if (tid < 12) s_slotcnt[tid] = 0;
    __syncthreads();
    if (tid < 32) {
        // class totals: lane c reads the table of class c (MM: 13, 14 -> 13, 14; 16 -> 15; 15 = 2 x classes 2..5)
        long long tot = 0;
        if (lane < NCX) {
            if (MM && lane == C_MOVE_MONO + 2) {
                for (int c = C_MOVE_DIV; c < C_MOVE_DIV + 4; c++) for (int i = 0; i < 8; i++) tot += 2 * (long long)s_tot[i][c];
            } else {
                const int tb = (MM && lane == C_MOVE_MONO + 3) ? 15 : lane;
                for (int i = 0; i < 8; i++) tot += (long long)s_tot[i][tb];
            }
        }
        if (lane < NCA) q.counts[(size_t)rep * NCA + lane] = lane < NCX ? (unsigned long long)tot : 0ull;
        const int myc = lane < NCX ? lane : 0;
        const double rate = (lane < NCX && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
        const int pos = (lane < NCX && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
        const double u1 = s_u1, u2 = s_u2;
        const double w = __dmul_rn((double)tot, rate);
        PickState pstate = pick_state_init_n<NL>(lane);
        const ClassPick pick = pick_class<NL>(w, pos, pstate, u1, lane);
        if (p.log_mode == KMC_LOG_FULL && lane < NCA)
            (reinterpret_cast<kmc_event_full *>(p.events) + idx)->counts[lane] = lane < NCX ? (uint32_t)tot : 0u;
        const long long cnt = __shfl_sync(FULL, tot, pick.zero ? 0 : pick.csel);
        long long rr = (long long)__dmul_rn(u2, (double)cnt);
        if (rr > cnt - 1) rr = cnt - 1;
        if (lane == 0) {
            s_csel = pick.csel; s_zero = pick.zero; s_tie = pick.tie != 0; s_total = pick.total;
            s_rank = (unsigned long long)(rr < 0 ? 0 : rr);
        }
    }
    __syncthreads();
    if (s_zero) {
        if (tid == 0) {
            p.flags[rep] |= FLAG_ZERO_RATE;
            atomicOr(p.any_flag, FLAG_ZERO_RATE);
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = 0xFFFFFFFFu; rec->cls = KMC_CLASS_NONE; rec->slot = KMC_SLOT_NONE;
                rec->flags = 0; rec->total_rate = 0.0; rec->pad = 0;
            } else if (p.log_mode == KMC_LOG_COMPACT) {
                kmc_event_compact e; e.site = 0xFFFFFFFFu; e.cls = KMC_CLASS_NONE; e.slot = KMC_SLOT_NONE;
                e.flags = 0; e.total_rate = 0.0;
                reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
            }
        }
        return;
    }
    const int csel = s_csel;

    // ---- row: CTA record, then the rows of that CTA --------------------------------------------------------
    {
        const int b0 = tid * kb, b1 = min(b0 + kb, q.nblk);
        // table of the chosen class (MM: class 16 -> 15); split-divacancy is twice tables 2..5
        const bool split = MM && csel == C_MOVE_MONO + 2;
        const int tb = (MM && csel == C_MOVE_MONO + 3) ? 15 : csel;
        uint32_t mysel = mine_c[0];                        // this thread's records' sum of the chosen class (registers)
#pragma unroll
        for (int c = 1; c < (MM ? PRC_STRIDE : NC); c++) mysel = tb == c ? mine_c[c] : mysel;
        if (split) mysel = 2u * (mine_c[2] + mine_c[3] + mine_c[4] + mine_c[5]);
        auto blkval = [&](int b) -> unsigned long long {
            const uint32_t *rec = blk + (size_t)b * PRC_STRIDE;
            return split ? 2ull * ((unsigned long long)rec[2] + rec[3] + rec[4] + rec[5]) : rec[tb];
        };
        auto rowval_of = [&](int a) -> uint32_t {
            return split ? 2u * (rowc[(size_t)2 * d0 + a] + rowc[(size_t)3 * d0 + a] + rowc[(size_t)4 * d0 + a] + rowc[(size_t)5 * d0 + a])
                         : rowc[(size_t)tb * d0 + a];
        };
        unsigned long long local = mysel, total;
        const unsigned long long excl = block_exclusive_scan(local, wsum, tid, total);
        const unsigned long long rank = s_rank;
        if (rank >= excl && rank < excl + local) {
            // the one thread that holds the rank: its <= kb records, then the <= 8 rows of the record
            unsigned long long r = rank - excl;
            int b = b0;
            for (; b < b1 - 1; b++) {
                const unsigned long long v = blkval(b);
                if (r < v) break;
                r -= v;
            }
            const int a0 = b * q.rpb;
            uint32_t rv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) rv[i] = (i < q.rpb && a0 + i < d0) ? rowval_of(a0 + i) : 0u;
            int a = a0;
#pragma unroll
            for (int i = 0; i < 7; i++)
                if (a == a0 + i && i < q.rpb - 1 && r >= rv[i]) { r -= rv[i]; a = a0 + i + 1; }
            s_row = a; s_rin = (uint32_t)r;
        }
    }
    __syncthreads();
    const int row = s_row;

    // ---- slot and site inside the row: one thread per 64-site cell (rows of <= 16384 sites) --------------------
    constexpr int NSL = MM ? 12 : 6;
    u64 m[NSL];                                             // this thread's cell: sites of the class per slot
#pragma unroll
    for (int i = 0; i < NSL; i++) m[i] = 0;
    const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 :
                      ((MM && csel >= C_MOVE_MONO) ? 3 : 0));
    if (tid < Wa) {
        if (MM && group == 3) {
#pragma unroll
            for (int i = 0; i < NSL; i++) m[i] = mm_cell_mask(G, row, tid, i >> 1, (i & 1) + 1, csel - C_MOVE_MONO);
        } else if (group == 1) {
            Nbhd nb;
            load_nbhd<false>(G, row, tid, nb);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                u64 pr, ne, fa;
                cell_dir_masks(nb, k, pr, ne, fa);
                m[k] = kind_select(pr, ne, fa, csel - C_MOVE_DIV);
            }
        } else if (group == 2) {
            const u64 *ra = G.rowp(T_D, row), *rb2 = G.rowp(T_D, wrap1(row + 2, d0));
            const u64 a0 = ra[tid], a1 = ra[tid + 1], b0 = rb2[tid], b1 = rb2[tid + 1];
            const u64 both = window(a0, a1, 0) & window(b0, b1, 0) & G.valid(tid);
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
            const u64 v = G.valid(tid);
            m[0] = window_at(G.rowp(pa, row), tid, 0) & v;
            if (csel == C_FLIP) m[0] |= window_at(G.rowp(T_M2, row), tid, 0) & v;
            if (pb >= 0) m[1] = window_at(G.rowp(pb, row), tid, 0) & v;
        }
    }
    {
#pragma unroll
        for (int i = 0; i < NSL; i += 2) {
            const uint32_t c = __reduce_add_sync(FULL, (uint32_t)__popcll(m[i]) | ((uint32_t)__popcll(m[i + 1]) << 16));
            if (lane == 0) {
                if (c & 0xFFFFu) atomicAdd(&s_slotcnt[i], c & 0xFFFFu);
                if (c >> 16) atomicAdd(&s_slotcnt[i + 1], c >> 16);
            }
        }
    }
    __syncthreads();
    uint32_t r = s_rin;
    int j = 0;
#pragma unroll
    for (int i = 0; i < NSL - 1; i++)
        if (j == i && r >= s_slotcnt[i]) { r -= s_slotcnt[i]; j = i + 1; }
    u64 mj = m[0];
#pragma unroll
    for (int i = 1; i < NSL; i++) mj = j == i ? m[i] : mj;
    {
        const unsigned long long local = (unsigned long long)__popcll(mj);
        unsigned long long total;
        const unsigned long long excl = block_exclusive_scan(local, wsum, tid, total);
        if ((unsigned long long)r >= excl && (unsigned long long)r < excl + local) {
            s_word = tid; s_bit = nth_bit64(mj, r - (uint32_t)excl);
        }
    }
    __syncthreads();
    const int col = 64 * s_word + s_bit;
    const int slot = class_slot_base(csel) + j;
    int s0;
    if (MM && group == 3) s0 = csel < C_MOVE_MONO + 2 ? (j & 1) + 1 : S_DIVAC;
    else if (group) s0 = S_DIVAC;
    else switch (csel) {
        case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0 = 0; break;
        case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0 = 3; break;
        case C_DESTROY_TREF: s0 = j ? 7 : 4; break;
        case C_CMONO_MAKE_EMPTY: s0 = j ? 1 : 2; break;
        case C_FMONO_MAKE_DOUBLE: s0 = j ? 2 : 1; break;
        default: s0 = ((window_at(G.rowp(T_M1, row), s_word, 0) >> s_bit) & 1ull) ? 1 : 2; break;
    }
    const int site = row * d1 + col;
    // ---- the move: thread (node i, plane pl) = tid < 18 rewrites its plane bit (+ halo copy) ---------------------
    MoveNodes mv = move_nodes(row, col, slot, s0, d0, d1);
    int os1, os2;
    {
        const MoveEntry e = MOVE_TABLE[slot];
        os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
        if (MM && slot >= MM_SLOT0) {
            const int layer = ((slot - MM_SLOT0) & 1) + 1;
            const bool onto_mono = (csel - C_MOVE_MONO) & 1;
            os1 = onto_mono ? 3 - layer : 0;
            mv.ns1 = onto_mono ? 3 : layer;
        }
    }
    __syncthreads();                                       // every read of the planes above is done
    if (tid < 18) {
        const int i = tid / 6, pl = tid % 6;
        const int as = i == 0 ? row : (i == 1 ? mv.a1 : mv.a2), bc = i == 0 ? col : (i == 1 ? mv.b1 : mv.b2);
        const int ns = i == 0 ? mv.ns0 : (i == 1 ? mv.ns1 : mv.ns2), os = i == 0 ? s0 : (i == 1 ? os1 : os2);
        if (i < mv.na && (status_plane(os) == pl || status_plane(ns) == pl)) {
            u64 *rowp = P + (size_t)(pl * d0 + as) * We;
            const bool set = status_plane(ns) == pl;
            const int q0 = bc + 2;
            const int q1 = bc < 2 ? bc + d1 + 2 : (bc >= d1 - 2 ? bc + 2 - d1 : -1);
            // (two nodes of a move can share a word of the same plane: atomic bit operations)
            {
                unsigned long long *wp = rowp + (q0 >> 6); const u64 bit = 1ull << (q0 & 63);
                if (set) atomicOr(wp, bit); else atomicAnd(wp, ~bit);
            }
            if (q1 >= 0) {
                unsigned long long *wp = rowp + (q1 >> 6); const u64 bit = 1ull << (q1 & 63);
                if (set) atomicOr(wp, bit); else atomicAnd(wp, ~bit);
            }
        }
    }
    if (tid == 0) {
        if (p.tracer)
            tracer_update(p.tracer + (size_t)rep * p.n, slot, mv.na, site, mv.ns0, mv.a1 * d1 + mv.b1, mv.ns1,
                          mv.a2 * d1 + mv.b2, mv.ns2, s0);
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
            rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
            rec->flags = s_tie ? 1 : 0; rec->total_rate = s_total; rec->pad = 0;
        } else if (p.log_mode == KMC_LOG_COMPACT) {
            kmc_event_compact e; e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
            e.flags = s_tie ? 1 : 0; e.total_rate = s_total;
            reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
        }
        p.steps_done[rep] = step + 1;
        if (p.clock) p.clock[rep] = __dadd_rn(p.clock[rep], kmc_time_step(p.clock_mode, s_total, p.seed,
                                                                          p.replica0 + (uint32_t)rep, step));
        p.hist[(size_t)rep * NCA + csel] += 1;
        if (slot >= 7 && slot < 13) p.hops[(size_t)rep * 6 + slot - 7] += 1;
        if (s_tie) p.ties[rep] += 1;
    }
}

}  // namespace kmc
