// noinc_planes_kernel.cuh -- the `--no-incremental` event (reference demo1/sim.py:114-143: re-enumerate everything,
// choose, perform) on the BIT-SLICED lattice, for lattices held in HBM (BASELINE config 4: 4096 x 4096).
//
// The byte-lattice version (sweep_kernel.cuh count-only + noinc_kernel.cuh) spends ~1.4 warp-instructions per site
// on the enumeration (nibble-parallel, 8 sites per operation) and selects on one CTA from a strided row table:
// 23 + 18 us per event at 4096^2.  Here the lattice lives as the six predicate planes of replica_wide_kernel.cuh
// (Z, D, M1, M2, A4, A7; 0.75 B per site, two-column halo per row end), and ONE launch does the whole event:
//   * enumeration: a warp owns PNR consecutive rows x 32 cells (a cell = 64 sites); a lane loads the plane words of
//     rows a-1 .. a+PNR+1 of its cell once, builds every shifted window once (they are shared by the anchor rows
//     and the six directions) and popcounts the MoveDivacancy / trefoil / MoveMonovacancy masks;
//   * row counts leave the CTA class-major (coalesced for the row search) together with one record of per-CTA sums
//     (a two-level hierarchy: CTA -> row), no atomics, nothing to zero;
//   * the CTA that finishes last (a ticket per replica) is the selector: it sums the records, chooses the class
//     (same fp64 sequence as every other kernel), finds record -> row -> slot -> cell -> bit with three dependent
//     rounds of loads, and flips the touched bits of the planes (and their halo copies) with atomics.
// Tried and dropped (profiles/README.md): a fixed selector CTA per replica; relaxed polling with back-off; a rolled
// row loop in the enumeration (smaller code, no faster); an unrolled one-shot
// class choice (1200 straight-line instructions run once by a lone warp cost ~20 cycles of instruction fetch each).
// (History: two launches -- 1024 CTAs of one row x 64 cells, then a one-CTA selection kernel with ~10 dependent L2
// round trips -- took 10.8 + 11.4 us.)
// Status bytes are only materialised (kmc_wide_unpack_kernel) when somebody asks for them.
// Results are bit-identical to the byte-lattice path (tests: both against the oracle).
#pragma once
#include "noinc_kernel.cuh"
#include "replica_wide2_kernel.cuh"

namespace kmc {

#ifdef KMC_NOINC_TRACE
__device__ long long g_noinc_trace[16];
#define NTRACE(i) do { if (threadIdx.x == 0) atomicAdd((unsigned long long *)&g_noinc_trace[i], (unsigned long long)(clock64() - tr0)); } while (0)
#else
#define NTRACE(i) do { } while (0)
#endif

constexpr int PRC_STRIDE = 16;             // words per CTA record / classes per row-count plane
constexpr int PNR = 2;                     // consecutive rows per warp in the enumeration

__host__ __device__ inline int planes_chunks(int d1) {       // warps per row group: power of two >= ceil(Wa / 32), <= 8
    const int need = (wide_anchor_words(d1) + 31) >> 5;
    int c = 1;
    while (c < need) c <<= 1;
    return c;
}
__host__ __device__ inline int planes_rows_per_cta(int d1) { return (8 / planes_chunks(d1)) * PNR; }

struct NoincPlanesParams {
    NoincParams base;                    // rate / order / draws / statistics pointers; status, rowcnt unused
    u64 *planes;                         // [R][6][d0][We]
    uint32_t *rowc;                      // [R][16][d0]   class-major row counts (classes 0..12; MM: 13, 14, 15)
    uint32_t *blk;                       // [R][nblk][16] per-CTA sums of the same
    unsigned long long *counts;          // [R][NCA] class totals of this event's enumeration (out)
    unsigned int *ticket;                // [R] arrivals; zero between launches
    unsigned int *epoch;                 // [R] persistent form: events of this call completed so far (zero on entry)
    int nblk, rpb, chunks, chunk_shift;
    int skip_select;                     // measurement only (KMC_NOINC_DEBUG_SKIP): enumerate, elect, do not select
};

// Word wc of a plane row and the low half of word wc + 1 (all that ever shifts into a window): two independent loads.
// (Taking the second from the neighbour lane by shuffle saves the load but makes every window wait for a shuffle:
// 14.5 instead of 14.0 us per 4096^2 event.)
__device__ __forceinline__ u64 plane_word_pair(const u64 *rp, int wc, uint32_t &x1) {
    x1 = reinterpret_cast<const uint32_t *>(rp + wc + 1)[0];
    return rp[wc];
}

// the count of class c (record slot c) from the unpacked field sums: lo/hi[0] = pristine / divacancy sites,
// lo/hi[1] = monovacancies / trefoil anchors, lo/hi[2], lo/hi[3] = MoveDivacancy kinds, tr = trefoil sites
template <bool MM>
__device__ __forceinline__ uint32_t planes_class_value(int c, const uint32_t lo[4], const uint32_t hi[4], uint32_t tr,
                                                       uint32_t nat, uint32_t mer, uint32_t swp) {
    const uint32_t pz = lo[0], pd = hi[0], pm = lo[1], pa = hi[1];
    switch (c) {
        case C_CREATE_DIV: return pz;
        case C_FILL_DIV: return pd;
        case C_MOVE_DIV: return lo[2];
        case C_MOVE_DIV + 1: return hi[2];
        case C_MOVE_DIV + 2: return lo[3];
        case C_MOVE_DIV + 3: return hi[3];
        case C_CREATE_TREF: return tr;
        case C_DESTROY_TREF: return pa;
        case C_CMONO_FROM_DOUBLE: return 2 * pz;
        case C_FMONO_FROM_EMPTY: return 2 * pd;
        case C_CMONO_MAKE_EMPTY: case C_FMONO_MAKE_DOUBLE: case C_FLIP: return pm;      // one per monovacancy
        case 13: return MM ? nat : 0;
        case 14: return MM ? mer : 0;
        case 15: return MM ? swp : 0;
        default: return 0;
    }
}

// ---- enumeration: the CTA's rows ---------------------------------------------------------------------------------
// MM: slots 13, 14, 15 of a record hold MoveMonovacancy natural, merge-divacancy, swap-divacancy (class 16); the
// split-divacancy count (class 15) is twice the sum of slots 2..5 and is not stored.
template <bool MM>
__device__ __forceinline__ void planes_enumerate(const NoincPlanesParams &q, int rep, int b,
                                                 uint32_t (*part)[PNR][8])
{
    const int d0 = q.base.d0, d1 = q.base.d1;
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = q.chunks, groups = 8 / chunks;
    const int g = warp >> q.chunk_shift, chunk = warp & (chunks - 1);     // chunks is a power of two
    const int a0 = (b * groups + g) * PNR;                         // first row of this warp
    const int w = chunk * 32 + lane;
    uint32_t acc[PNR][7];
#pragma unroll
    for (int i = 0; i < PNR; i++)
#pragma unroll
        for (int f = 0; f < 7; f++) acc[i][f] = 0;
    if (a0 < d0) {                                                 // (warp-uniform)
        const u64 *P = q.planes + (size_t)rep * N_TYPES * d0 * We;
        WGeom G{P, d0, d1, Wa, We};
        // lanes beyond the row's last cell run along on its last cell with an empty `valid` mask
        const int wc = min(w, Wa - 1);
        const u64 valid = w < Wa ? G.valid(w) : 0ull;
        // windows of rows a0-1 .. a0+PNR+1 (index r = offset + 1): Z and the monovacancy layers at db -1, 0, +1
        // (index db + 1), D at db -2 .. +2 (index db + 2)
        u64 Zw[PNR + 2][3], Dw[PNR + 3][5], M1w[MM ? PNR + 2 : 1][3], M2w[MM ? PNR + 2 : 1][3];
#pragma unroll
        for (int r = 0; r < PNR + 3; r++) {
            const int ar = wrap1(a0 + r - 1, d0);                  // a0 < d0 and d0 >= 5: wraps at most once
            uint32_t x1;
            const u64 x0 = plane_word_pair(G.rowp(T_D, ar), wc, x1);
#pragma unroll
            for (int j = 0; j < 5; j++) Dw[r][j] = window(x0, (u64)x1, j - 2);
            if (r < PNR + 2) {
                uint32_t y1;
                const u64 y0 = plane_word_pair(G.rowp(T_Z, ar), wc, y1);
#pragma unroll
                for (int j = 0; j < 3; j++) Zw[r][j] = window(y0, (u64)y1, j - 1);
                if (MM) {
                    uint32_t u1, v1;
                    const u64 u0 = plane_word_pair(G.rowp(T_M1, ar), wc, u1), v0 = plane_word_pair(G.rowp(T_M2, ar), wc, v1);
#pragma unroll
                    for (int j = 0; j < 3; j++) { M1w[r][j] = window(u0, (u64)u1, j - 1); M2w[r][j] = window(v0, (u64)v1, j - 1); }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < PNR; i++) {
            if (a0 + i < d0) {
                const int a = a0 + i;
                const u64 dv = Dw[i + 1][2] & valid, zv = Zw[i + 1][1] & valid;
                u64 mono, anch, m1 = 0, m2 = 0;
                {
                    uint32_t s1, t1;
                    const u64 s0 = plane_word_pair(G.rowp(T_A4, a), wc, s1), t0 = plane_word_pair(G.rowp(T_A7, a), wc, t1);
                    anch = window(s0 | t0, (u64)(s1 | t1), 0) & valid;
                    if (MM) { m1 = M1w[i + 1][1] & valid; m2 = M2w[i + 1][1] & valid; mono = m1 | m2; }
                    else {
                        uint32_t u1, v1;
                        const u64 u0 = plane_word_pair(G.rowp(T_M1, a), wc, u1), v0 = plane_word_pair(G.rowp(T_M2, a), wc, v1);
                        mono = window(u0 | v0, (u64)(u1 | v1), 0) & valid;
                    }
                }
                uint32_t nP = 0, nPN = 0, nPF = 0, nPNF = 0, nm = 0, sw = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const int rn = i + 1 + nib(NBR_DA, k), re = i + 1 + nib(NEAR_DA, k), rf = i + 1 + nib(FAR_DA, k);
                    const u64 zn = Zw[rn][nib(NBR_DB, k) + 1];
                    // present, present & near, present & far, all three: the four kinds follow by inclusion-exclusion
                    const u64 pr = dv & zn, pn = pr & Dw[re][nib(NEAR_DB, k) + 2], pf = pr & Dw[rf][nib(FAR_DB, k) + 2];
                    nP += __popcll(pr); nPN += __popcll(pn); nPF += __popcll(pf);
                    nPNF += __popcll(pn & pf);
                    if (MM) {
                        const u64 m1n = M1w[rn][nib(NBR_DB, k) + 1], m2n = M2w[rn][nib(NBR_DB, k) + 1];
                        nm += (uint32_t)__popcll((m1 | m2) & zn) | ((uint32_t)__popcll((m1 & m2n) | (m2 & m1n)) << 16);
                        sw += (uint32_t)__popcll(dv & (m1n | m2n));
                    }
                }
                const u64 both = dv & Dw[i + 3][2];
                acc[i][0] = (uint32_t)__popcll(zv) | ((uint32_t)__popcll(dv) << 16);
                acc[i][1] = (uint32_t)__popcll(mono) | ((uint32_t)__popcll(anch) << 16);
                acc[i][2] = (nP - nPN - nPF + nPNF) | ((nPN - nPNF) << 16);        // natural | missing-metal
                acc[i][3] = (nPF - nPNF) | (nPNF << 16);                            // missing-comb | missing-both
                acc[i][4] = (uint32_t)(__popcll(both & Dw[i + 1][4]) + __popcll(both & Dw[i + 3][0]));
                acc[i][5] = nm; acc[i][6] = sw;
            }
        }
    }
    // 16-bit fields: a warp covers 2048 sites of a row with <= 6 moves per site and class (<= 12288 per field)
#pragma unroll
    for (int i = 0; i < PNR; i++)
#pragma unroll
        for (int f = 0; f < (MM ? 7 : 5); f++) {
            const uint32_t s = __reduce_add_sync(FULL, acc[i][f]);
            if (lane == 0) part[warp][i][f] = s;
        }
    __syncthreads();
    // thread (row slot rs, class c) combines the row's chunks (fields unpacked first) and writes the row count
    const int rpb = groups * PNR;
    if (tid < rpb * PRC_STRIDE) {
        const int rs = tid >> 4, c = tid & 15;
        const int gg = rs / PNR, i = rs - gg * PNR;                   // (PNR is a constant power of two)
        const int a = (b * groups + gg) * PNR + i;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, tr = 0, nat = 0, mer = 0, swp = 0;
        for (int ch = 0; ch < chunks; ch++) {
            const uint32_t *pw = part[gg * chunks + ch][i];
#pragma unroll
            for (int f = 0; f < 4; f++) { const uint32_t x = pw[f]; lo[f] += x & 0xFFFFu; hi[f] += x >> 16; }
            tr += pw[4];
            if (MM) { nat += pw[5] & 0xFFFFu; mer += pw[5] >> 16; swp += pw[6]; }
        }
        const uint32_t val = planes_class_value<MM>(c, lo, hi, tr, nat, mer, swp);
        if (a < d0 && c < (MM ? PRC_STRIDE : NC)) q.rowc[((size_t)rep * PRC_STRIDE + c) * d0 + a] = val;
    }
    // the CTA record: lane c of the last warp sums class c over all of the CTA's warps and rows (rows beyond the
    // lattice contributed zeros)
    if (tid >= 224 && tid < 224 + PRC_STRIDE) {
        const int c = tid - 224;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, tr = 0, nat = 0, mer = 0, swp = 0;
        for (int wq = 0; wq < 8; wq++)
#pragma unroll
            for (int i = 0; i < PNR; i++) {
                const uint32_t *pw = part[wq][i];
#pragma unroll
                for (int f = 0; f < 4; f++) { const uint32_t x = pw[f]; lo[f] += x & 0xFFFFu; hi[f] += x >> 16; }
                tr += pw[4];
                if (MM) { nat += pw[5] & 0xFFFFu; mer += pw[5] >> 16; swp += pw[6]; }
            }
        q.blk[((size_t)rep * q.nblk + b) * PRC_STRIDE + c] = planes_class_value<MM>(c, lo, hi, tr, nat, mer, swp);
    }
}

__device__ __forceinline__ void noinc_record_none(const NoincParams &p, size_t idx) {
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

// ---- selection + move: the 256 threads of the replica's last CTA ---------------------------------------------------
template <bool MM>
__device__ __forceinline__ void planes_select(const NoincPlanesParams &q, int rep, long long t, unsigned long long step,
                                              const double *s_u, long long tr0)
{
    (void)tr0;
    constexpr int NCX = MM ? NCA : NC;
    constexpr int NL = MM ? 32 : 16;
    const NoincParams &p = q.base;
    __shared__ uint32_t s_tot[8][PRC_STRIDE];               // per-warp class sums of the CTA records
    __shared__ unsigned long long s_rank;
    __shared__ double s_total;
    __shared__ int s_csel, s_zero, s_tie, s_row, s_word, s_bit, s_layer1;
    __shared__ uint32_t s_rin, s_wslot[8][12];
    const int tid = threadIdx.x, lane = tid & 31;
    const int d0 = p.d0, d1 = p.d1;
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)t;
    // ---- round 1 of loads: everything that does not depend on the enumeration's outcome, and the CTA records -------
    const uint32_t flag0 = p.flags[rep];
    double clk = 0.0;
    if (tid == 0 && p.clock) clk = p.clock[rep];
    u64 *P = q.planes + (size_t)rep * N_TYPES * d0 * We;
    WGeom G{P, d0, d1, Wa, We};
    const uint32_t *blk = q.blk + (size_t)rep * q.nblk * PRC_STRIDE;
    const uint32_t *rowc = q.rowc + (size_t)rep * PRC_STRIDE * d0;
    // every thread owns kb consecutive CTA records and keeps the class sums of them in registers: they serve the
    // totals now and the search for the thread that holds the rank below
    const int kb = (q.nblk + 255) / 256;
    uint32_t mine_c[PRC_STRIDE];
#pragma unroll
    for (int c = 0; c < PRC_STRIDE; c++) mine_c[c] = 0;
    for (int i = 0; i < kb; i++) {
        const int b = tid * kb + i;
        if (b < q.nblk) {
            const uint4 *rec = reinterpret_cast<const uint4 *>(blk + (size_t)b * PRC_STRIDE);
            const uint4 x0 = __ldcg(rec), x1 = __ldcg(rec + 1), x2 = __ldcg(rec + 2), x3 = __ldcg(rec + 3);
            mine_c[0] += x0.x; mine_c[1] += x0.y; mine_c[2] += x0.z; mine_c[3] += x0.w;
            mine_c[4] += x1.x; mine_c[5] += x1.y; mine_c[6] += x1.z; mine_c[7] += x1.w;
            mine_c[8] += x2.x; mine_c[9] += x2.y; mine_c[10] += x2.z; mine_c[11] += x2.w;
            mine_c[12] += x3.x; mine_c[13] += x3.y; mine_c[14] += x3.z; mine_c[15] += x3.w;
        }
    }
    if (flag0 & FLAG_ZERO_RATE) {                        // this replica already stopped (uniform): no event
        if (tid == 0) noinc_record_none(p, idx);
        return;
    }
    {
        uint32_t mysum = 0;                                 // lane c of each warp collects the warp's sum of class c
#pragma unroll
        for (int c = 0; c < (MM ? PRC_STRIDE : NC); c++) {
            const uint32_t sc = __reduce_add_sync(FULL, mine_c[c]);
            if (lane == c) mysum = sc;
        }
        if (lane < PRC_STRIDE) s_tot[tid >> 5][lane] = lane < (MM ? PRC_STRIDE : NC) ? mysum : 0u;
    }
    __syncthreads();
    NTRACE(2);
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
        NTRACE(8);
        const int myc = lane < NCX ? lane : 0;
        const double rate = (lane < NCX && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
        const int pos = (lane < NCX && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
        const double u1 = s_u[0], u2 = s_u[1];             // (made at the top of the event, see the kernel)
        const double w = __dmul_rn((double)tot, rate);
        PickState pstate = pick_state_init_n<NL>(lane);
        NTRACE(9);
        const ClassPick pick = pick_class<NL>(w, pos, pstate, u1, lane);
        NTRACE(10);
        if (p.log_mode == KMC_LOG_FULL && lane < NCA)
            (reinterpret_cast<kmc_event_full *>(p.events) + idx)->counts[lane] = lane < NCX ? (uint32_t)tot : 0u;
        const long long cnt = __shfl_sync(FULL, tot, pick.zero ? 0 : pick.csel);
        long long rr = (long long)__dmul_rn(u2, (double)cnt);
        if (rr > cnt - 1) rr = cnt - 1;
        NTRACE(11);
        if (lane == 0) {
            s_csel = pick.csel; s_zero = pick.zero; s_tie = pick.tie != 0; s_total = pick.total;
            s_rank = (unsigned long long)(rr < 0 ? 0 : rr);
        }
    }
    __syncthreads();
    NTRACE(3);
    if (s_zero) {
        if (tid == 0) {
            p.flags[rep] = flag0 | FLAG_ZERO_RATE;
            atomicOr(p.any_flag, FLAG_ZERO_RATE);
            noinc_record_none(p, idx);
        }
        return;
    }
    const int csel = s_csel;

    // ---- row: the thread whose records hold the rank walks their rows (round 2 of loads) ------------------------------
    {
        // table of the chosen class (MM: class 16 -> 15); split-divacancy is twice tables 2..5
        const bool split = MM && csel == C_MOVE_MONO + 2;
        const int tb = (MM && csel == C_MOVE_MONO + 3) ? 15 : csel;
        uint32_t mysel = mine_c[0];                        // this thread's records' sum of the chosen class (registers)
#pragma unroll
        for (int c = 1; c < (MM ? PRC_STRIDE : NC); c++) mysel = tb == c ? mine_c[c] : mysel;
        if (split) mysel = 2u * (mine_c[2] + mine_c[3] + mine_c[4] + mine_c[5]);
        auto rowval_of = [&](int a) -> uint32_t {
            return split ? 2u * (__ldcg(rowc + (size_t)2 * d0 + a) + __ldcg(rowc + (size_t)3 * d0 + a) +
                                 __ldcg(rowc + (size_t)4 * d0 + a) + __ldcg(rowc + (size_t)5 * d0 + a))
                         : __ldcg(rowc + (size_t)tb * d0 + a);
        };
        // exclusive prefix of `mysel` over the CTA: the warps before this one from the per-warp class sums that made
        // the totals (s_tot), the lanes before this one by a warp scan -- no barrier
        unsigned long long excl = 0;
        for (int wq = 0; wq < (tid >> 5); wq++)
            excl += split ? 2ull * ((unsigned long long)s_tot[wq][2] + s_tot[wq][3] + s_tot[wq][4] + s_tot[wq][5])
                          : (unsigned long long)s_tot[wq][tb];
        const unsigned long long local = mysel;
        {
            uint32_t inc = mysel;                          // (a warp's records hold fewer than 2^32 moves of a class)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
            excl += inc - mysel;
        }
        const unsigned long long rank = s_rank;
        if (rank >= excl && rank < excl + local) {
            uint32_t r = (uint32_t)(rank - excl);
            const int a_lo = tid * kb * q.rpb, a_hi = min((tid + 1) * kb * q.rpb, d0);
            int row = a_hi - 1;
            bool found = false;
            for (int a = a_lo; a < a_hi && !found; a += 16) {
                uint32_t rv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) rv[i] = a + i < a_hi ? rowval_of(a + i) : 0u;    // independent loads
#pragma unroll
                for (int i = 0; i < 16; i++)
                    if (!found && a + i < a_hi) {
                        if (r < rv[i]) { found = true; row = a + i; }
                        else r -= rv[i];
                    }
            }
            s_row = row; s_rin = r;
        }
    }
    __syncthreads();
    NTRACE(4);
    const int row = s_row;
    u64 flip_m1 = 0;

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
            flip_m1 = m[0];
            if (csel == C_FLIP) m[0] |= window_at(G.rowp(T_M2, row), tid, 0) & v;
            if (pb >= 0) m[1] = window_at(G.rowp(pb, row), tid, 0) & v;
        }
    }
    {
        // per-warp slot counts (plain stores: every warp writes its own line, nothing to zero)
#pragma unroll
        for (int i = 0; i < NSL; i += 2) {
            const uint32_t c = __reduce_add_sync(FULL, (uint32_t)__popcll(m[i]) | ((uint32_t)__popcll(m[i + 1]) << 16));
            if (lane == 0) { s_wslot[tid >> 5][i] = c & 0xFFFFu; s_wslot[tid >> 5][i + 1] = c >> 16; }
        }
    }
    __syncthreads();
    NTRACE(5);
    uint32_t r = s_rin;
    int j = 0;
#pragma unroll
    for (int i = 0; i < NSL - 1; i++) {
        uint32_t tot = 0;
#pragma unroll
        for (int wq = 0; wq < 8; wq++) tot += s_wslot[wq][i];
        if (j == i && r >= tot) { r -= tot; j = i + 1; }
    }
    u64 mj = m[0];
#pragma unroll
    for (int i = 1; i < NSL; i++) mj = j == i ? m[i] : mj;
    {
        // exclusive prefix of the cells' counts: warps before this one from s_wslot, lanes by a warp scan
        const uint32_t local = (uint32_t)__popcll(mj);
        uint32_t excl = 0;
        for (int wq = 0; wq < (tid >> 5); wq++) excl += s_wslot[wq][j];
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        excl += inc - local;
        if (r >= excl && r < excl + local) {
            s_word = tid; s_bit = nth_bit64(mj, r - excl);
            s_layer1 = (int)((flip_m1 >> s_bit) & 1ull);
        }
    }
    __syncthreads();
    NTRACE(6);
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
        default: s0 = s_layer1 ? 1 : 2; break;             // FlipMonovacancy: the layer the vacancy is in
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
    NTRACE(7);
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
        if (p.clock) p.clock[rep] = __dadd_rn(clk, kmc_time_step(p.clock_mode, s_total, p.seed, p.replica0 + (uint32_t)rep, step));
        // (statistics of one replica are only ever touched by its selector: fire-and-forget adds, no load round trips)
        atomicAdd(p.hist + (size_t)rep * NCA + csel, 1ull);
        if (slot >= 7 && slot < 13) atomicAdd(p.hops + (size_t)rep * 6 + slot - 7, 1ull);
        if (s_tie) atomicAdd(p.ties + rep, 1ull);
    }
}

// grid = R * nblk CTAs of 256 threads; CTA (rep, b) enumerates rows b * rpb .. of replica rep.
//  * PERSIST = false: one launch per event (local step index t0), chained by programmatic dependent launch;
//  * PERSIST = true: ONE cooperative launch runs events t0 .. t1-1.  After its enumeration every CTA waits until the
//    replica's selector has published the event (`epoch`, release / acquire through fences as in a grid barrier), so
//    the ~4 us it costs to start a dependent grid are paid once per call instead of once per event.  All CTAs must be
//    co-resident (the host checks the occupancy and uses the cooperative launch API); a bounded spin turns a
//    missing selector into FLAG_STALL instead of a hang.
constexpr uint32_t FLAG_STALL = 8u;
template <bool MM = false, bool PERSIST = false>
__global__ void __launch_bounds__(256, MM ? 2 : 4) kmc_noinc_planes_event_kernel(const NoincPlanesParams q, const long long t0, const long long t1)
{
    __shared__ uint32_t part[8][PNR][8];
    __shared__ int s_last, s_abort;
    __shared__ double s_u[2];
    __shared__ uint32_t s_touch;
    const int rep = q.base.n_replicas == 1 ? 0 : blockIdx.x / q.nblk, b = blockIdx.x - rep * q.nblk;
    if (!PERSIST) {
        // Programmatic dependent launch: the next event's CTAs may become resident while this grid runs (they hold at
        // their own `wait` until this grid has completed and its writes are visible).  Nothing the previous grid
        // wrote is read above this line.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    // the replica's event counter when this launch started (a stopped replica does not advance it, nor does it use it)
    const unsigned long long step0 = q.base.steps_done[rep];
    if (threadIdx.x == 64) s_touch = MOVE_TABLE[0].w1;      // the selector reads the move table late: have it in the constant cache
    for (long long t = t0; t < t1; t++) {
        // the event's draws depend on nothing the enumeration makes: one thread has them ready for the selector
        const unsigned long long step = step0 + (unsigned long long)(t - t0);
#ifdef KMC_NOINC_TRACE
        const long long tr0 = clock64();
#else
        const long long tr0 = 0;
#endif
        if (threadIdx.x == 32) {
            double u1, u2;
            if (q.base.draws) {
                const double2 u = *reinterpret_cast<const double2 *>(q.base.draws + ((size_t)rep * (size_t)q.base.nsteps + (size_t)t) * 2);
                u1 = u.x; u2 = u.y;
            } else {
                kmc_draw(q.base.seed, q.base.replica0 + (uint32_t)rep, step, u1, u2);
            }
            s_u[0] = u1; s_u[1] = u2;
        }
        planes_enumerate<MM>(q, rep, b, part);
#ifdef KMC_NOINC_TRACE
        const long long c0 = clock64() - tr0;
#endif
        // row counts and the record are written; the barrier orders them before thread 0's release (cumulative),
        // the ticket with it: the CTA that draws the last ticket sees everything every CTA wrote
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tk;                                // release our rows, acquire everybody else's
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(tk) : "l"(q.ticket + rep) : "memory");
            s_last = tk == (unsigned)q.nblk - 1u;
            if (s_last) q.ticket[rep] = 0u;
        }
        __syncthreads();
        // (A fixed selector -- the replica's CTA 0 waiting for all arrivals, to keep the selection code warm in one SM's
        // instruction caches -- was tried: no faster, and it cannot overlap the next grid's start without persistence.)
        const bool last = s_last != 0;
#ifdef KMC_NOINC_TRACE
        if (last && threadIdx.x == 0) {
            atomicAdd((unsigned long long *)&g_noinc_trace[0], (unsigned long long)c0);
            atomicAdd((unsigned long long *)&g_noinc_trace[1], (unsigned long long)(clock64() - tr0));
            atomicAdd((unsigned long long *)&g_noinc_trace[15], 1ull);
        }
#endif
        if (last && !q.skip_select) planes_select<MM>(q, rep, t, step, s_u, tr0);
        if (!PERSIST) return;
        __syncthreads();                                    // the selector's threads are through
        if (last) NTRACE(12);
        if (threadIdx.x == 0) {
            int stalled = 0;
            if (last) {                                     // the move, the statistics and the ticket reset, then the epoch
                asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(q.epoch + rep), "r"((unsigned)(t + 1)) : "memory");
                NTRACE(13);
            } else {
                unsigned int spins = 0, seen;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(q.epoch + rep) : "memory");
                    if (seen >= (unsigned)(t + 1)) break;
                    if (++spins > (1u << 22)) { stalled = 1; break; }
                }
            }
            s_abort = stalled;
        }
        __syncthreads();
        if (s_abort) {
            if (threadIdx.x == 0) { atomicOr(q.base.flags + rep, FLAG_STALL); atomicOr(q.base.any_flag, FLAG_STALL); }
            return;
        }
    }
}

}  // namespace kmc
