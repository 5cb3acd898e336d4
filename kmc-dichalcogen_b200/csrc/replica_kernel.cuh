// replica_kernel.cuh -- K3 + K4: fused select -> apply -> neighbourhood-refresh kernel.
//
// One warp owns one replica (an independent trajectory, i.e. what the reference runs as one
// process: demo1/__main__.py:148).  The lattice (one status byte per site) and the Fenwick-style
// hierarchy used for the in-class rank search (per-row x per-class move counts, uint16) live in
// shared memory for the whole launch; the 13 per-class totals live in registers (lane c holds class
// c).  Many events run per launch; HBM is touched only to load/store the lattice once per launch, to
// fetch 16 B of draws per event in draw-injection mode, and to write event records when logging.
//
// One event = KmcSim.perform_random_move (reference demo1/sim.py:106-143):
//   1. weights w_c = count_c * rate_c                       (sim.py:120-121)
//   2. class choice: weighted_choice bit for bit            (kmc.py:21-54, util.py:32-46,65-68):
//      drop zeros, stable ascending sort, sequential fp64 prefix sums, x = total*u1, bisect_left
//   3. in-class pick: rank floor(u2*count) in (row, slot, column) order -- hierarchical search
//      rows -> slots of the class -> sites of the row   (replaces get_random_key, incremental.py:283)
//   4. apply the move                                        (rules.py perform())
//   5. refresh the neighbourhood: every site whose moves can depend on the touched nodes has its
//      13 class counts re-derived before and after the mutation; the difference goes to the row
//      counts and class totals   (pre/post_status_change, rules.py:22-27,46-51; sim.py:149-157)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <initializer_list>
#include "kmc_common.cuh"
#include "../../include/kmc_b200.h"

namespace kmc {

struct ReplicaParams {
    int d0, d1, n;
    int n_replicas;
    uint8_t *status;                 // [R][n]
    double rate[NCA];                // per-class rate (0 for disabled)
    int order_pos[NCA];              // position in weighted_choice input order, or -1
    long long nsteps;
    unsigned long long seed;
    uint32_t replica0;
    const double *draws;             // [R][nsteps][2] or nullptr (Philox)
    int log_mode;
    void *events;                    // [R][nsteps] records or nullptr
    unsigned long long *steps_done;  // [R] in: step offset; out: += events done
    unsigned long long *hist;        // [R][NCA]
    unsigned long long *hops;        // [R][6] MoveDivacancy events per direction k (slot 7 + k)
    unsigned long long *ties;        // [R]
    uint32_t *flags;                 // [R] bit0 zero-rate stop, bit1 validate mismatch
    uint32_t *any_flag;              // OR of all replicas' flags (one word: the asynchronous status of the handle)
    int validate;
    // HBM-resident lattices only: the row-count hierarchy of each replica, kept between launches so that
    // a launch does not have to re-enumerate a 1 Mi-site lattice first.  hier_valid != 0: load it.
    uint32_t *hier;                  // [R][d0 * RC_WORDS] or nullptr
    int hier_valid;
    // optional KMC clock: t += 1 / total_rate per event (mean residence time; the reference has no clock
    // but logs total_rate for exactly this post-processing, README.md:36-37).  nullptr = off.
    double *clock;                   // [R]
    int clock_mode;                  // 1: t += 1 / total_rate; 2: t += -ln(u3) / total_rate (kmc_common.cuh)
    // optional divacancy tracers (extension): per site the unwrapped displacement
    // (int16 da | int16 db << 16, axial) of the divacancy sitting there since it appeared.  nullptr = off.
    uint32_t *tracer;                // [R][n]
    // half-warp kernel with MoveMonovacancy (17 classes, 16 lanes): the half-lane that owns class 16 (swap-divacancy);
    // it is the lane of a DISABLED own-status class (capi.cu picks it)
    int lane16;
};

constexpr uint32_t FULL = 0xFFFFFFFFu;

// small offset tables packed as nibbles (value + 2), entry 0 in the lowest nibble
constexpr unsigned long long pack_nibbles(std::initializer_list<int> v) {
    unsigned long long r = 0; int i = 0;
    for (int x : v) { r |= (unsigned long long)(x + 2) << (4 * i); i++; }
    return r;
}
// ring neighbours R0..R5 (state.py:392-394)
constexpr uint32_t RING_DA = (uint32_t)pack_nibbles({-1, 0, 1, 1, 0, -1});
constexpr uint32_t RING_DB = (uint32_t)pack_nibbles({0, -1, -1, 0, 1, 1});
// sites whose move counts can depend on a touched node p: p, its six neighbours (MoveDivacancy,
// rules.py:115-117), and the anchors of the triangles containing p (p-(2,0), p-(0,2), p-(2,-2))
constexpr unsigned long long REFRESH_DA = pack_nibbles({0, -1, 0, 1, 1, 0, -1, -2, 0, -2});
constexpr unsigned long long REFRESH_DB = pack_nibbles({0, 0, -1, -1, 0, 1, 1, 0, -2, 2});
constexpr uint32_t FLAG_ZERO_RATE = 1u, FLAG_VALIDATE = 2u;

// Tracer bookkeeping for one event (uniform; executed by one lane).  A MoveDivacancy hop (slot 7 + k) carries
// the displacement from the old site to the new one and adds neighbour k of neighbors_and_mutuals
// (state.py:443-459); any other event that turns a site into a divacancy starts a new tracer at 0 there.
// Entries of sites that hold no divacancy are stale and never read.
__device__ __forceinline__ uint32_t tracer_step(int k) {
    const int ri = kring(k);
    const int da = (int)((RING_DA >> (4 * ri)) & 0xF) - 2, db = (int)((RING_DB >> (4 * ri)) & 0xF) - 2;
    return ((uint32_t)da & 0xFFFFu) | ((uint32_t)db << 16);
}
__device__ __forceinline__ void tracer_update(uint32_t *T, int slot, int na, int site0, int ns0, int site1, int ns1,
                                              int site2, int ns2, int s0 = 0) {
    if (slot >= 7 && slot < 13) {
        T[site1] = __vadd2(T[site0], tracer_step(slot - 7));
    } else if (slot >= MM_SLOT0) {
        // MoveMonovacancy: a divacancy that trades places with a monovacancy (swap) keeps its displacement; a
        // merge makes a new divacancy; a split ends one
        if (ns1 == S_DIVAC) T[site1] = s0 == S_DIVAC ? __vadd2(T[site0], tracer_step((slot - MM_SLOT0) >> 1)) : 0u;
    } else {
        if (ns0 == S_DIVAC) T[site0] = 0u;
        if (na > 1 && ns1 == S_DIVAC) T[site1] = 0u;
        if (na > 2 && ns2 == S_DIVAC) T[site2] = 0u;
    }
}

struct SmemStatus {
    const uint8_t *p;
    __device__ __forceinline__ int operator()(int i) const { return p[i]; }
};

__host__ __device__ inline size_t replica_rc_bytes(int d0, bool mm = false) {
    return ((size_t)d0 * (mm ? 9 : 7) * 4 + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t replica_warp_bytes(int d0, int d1, bool mm = false) {
    size_t n = (size_t)d0 * d1;
    size_t a = (n + 15) & ~(size_t)15;
    return a + replica_rc_bytes(d0, mm);
}

// find the lane whose inclusive prefix first exceeds r; v = per-lane count.
// returns false (and r -= warp total) when r is beyond this chunk.
__device__ __forceinline__ bool warp_rank_search(uint32_t v, uint32_t &r, int &L, int lane) {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    uint32_t tot = __shfl_sync(FULL, inc, 31);
    if (r >= tot) { r -= tot; return false; }
    uint32_t m = __ballot_sync(FULL, inc > r);
    L = __ffs(m) - 1;
    r -= __shfl_sync(FULL, inc - v, L);
    return true;
}

// weighted_choice (reference demo1/kmc.py:21-54) over 16 lanes holding (weight, input position).
// State kept across events (PickState): `sc` (lanes 0..15) = class held at each sorted position,
// `spos` = sorted position of the lane's own class, `wprev` = the lane's weight at the previous event,
// `cum` (lanes 0..15) = cumulative weight at the lane's sorted position.  The order is re-derived only when
// the (weight, position) order actually changed, and the sequential prefix sums are recomputed only from
// the first sorted position whose weight changed -- everything before it is bit-identical by construction.
struct PickState { int sc, spos; double wprev, cum; bool need_sort; };
__device__ __forceinline__ PickState pick_state_init(int lane) {
    PickState s; s.sc = lane & 15; s.spos = lane & 15; s.wprev = -1.0; s.cum = 0.0; s.need_sort = true; return s;
}
struct ClassPick { int csel; double total; uint32_t tie; bool zero; };
// NL = number of lanes that hold a (weight, position) pair: 16 for the 13 classes of the reference's own rules,
// 32 once MoveMonovacancy's four classes are in play (17 classes)
template <int NL = 16>
__device__ __forceinline__ PickState pick_state_init_n(int lane) {
    PickState s; s.sc = lane & (NL - 1); s.spos = lane & (NL - 1); s.wprev = -1.0; s.cum = 0.0; s.need_sort = true; return s;
}
template <int NL = 16>
__device__ __forceinline__ ClassPick pick_class(double w, int pos, PickState &st, double u1, int lane)
{
    constexpr uint32_t LANES = NL == 32 ? 0xFFFFFFFFu : 0xFFFFu;
    double ws = __shfl_sync(FULL, w, st.sc);
    const int ps = __shfl_sync(FULL, pos, st.sc);
    // first sorted position whose weight differs from the previous event
    int first = (int)__reduce_min_sync(FULL, (lane < NL && w != st.wprev) ? (unsigned)st.spos : (unsigned)NL);
    st.wprev = w;
    {
        // is the previous order still (w, pos)-sorted?  (almost always)
        const double wp = __shfl_up_sync(FULL, ws, 1);
        const int pp = __shfl_up_sync(FULL, ps, 1);
        const bool ok = (lane == 0) || (lane > NL - 1) || (wp < ws) || (wp == ws && pp < ps);
        if (st.need_sort || !__all_sync(FULL, ok)) {
            // stable ascending sort (kmc.py:35) as a rank sort of the NL unique (w, pos) keys
            int rank = 0;
            for (int j = 0; j < NL; j++) {
                const double wj = __shfl_sync(FULL, w, j);
                const int pj = __shfl_sync(FULL, pos, j);
                rank += (wj < w) || (wj == w && pj < pos);
            }
            st.spos = rank;
            for (int j = 0; j < NL; j++) {
                const int rj = __shfl_sync(FULL, rank, j);
                if (rj == lane) st.sc = j;
            }
            st.need_sort = false;
            ws = __shfl_sync(FULL, w, st.sc);
            first = 0;
        }
    }
    ClassPick out;
    // zero weights are dropped (kmc.py:30); they sort first and adding them is exact, so the
    // sequential sum simply starts after them
    const int z = __popc(__ballot_sync(FULL, ws == 0.0) & LANES);
    out.zero = (z == NL);
    const int start = first > z ? first : z;
    double acc = __shfl_sync(FULL, st.cum, (start + NL - 1) & (NL - 1));   // cumulative weight just before `start`
    if (start == z) acc = 0.0;
    for (int k = start; k < NL; k++) {
        const double wk = __shfl_sync(FULL, ws, k);
        acc = __dadd_rn(acc, wk);                              // util.py:65-68: sequential sums
        if (lane == k) st.cum = acc;
    }
    if (start == NL) acc = __shfl_sync(FULL, st.cum, NL - 1);  // nothing changed: total is the old one
    out.total = acc;
    const double x = __dmul_rn(acc, u1);                       // random.uniform(0, total), kmc.py:49
    const bool live = (lane >= z) && (lane < NL);
    const uint32_t ge = __ballot_sync(FULL, live && st.cum >= x); // bisect_left, kmc.py:50
    const int isel = ge ? (__ffs(ge) - 1) : NL - 1;
    out.tie = __ballot_sync(FULL, live && st.cum == x);
    out.csel = __shfl_sync(FULL, st.sc, isel);
    return out;
}

// 2 bits per status: number of class-c moves a site of that status carries (own-status classes)
__device__ __forceinline__ uint32_t g1_const(int c) {
    switch (c) {
        case C_CREATE_DIV: return 0x1u;
        case C_FILL_DIV: return 0x40u;
        case C_DESTROY_TREF: return 0x4100u;
        case C_CMONO_FROM_DOUBLE: return 0x2u;
        case C_FMONO_FROM_EMPTY: return 0x80u;
        default: return 0x14u;        // classes 9, 10, 12: monovacancy sites
    }
}

// Rule.perform + nodes_affected_by (reference demo1/rules.py:62-65,79-82,96-101,143-148,178-183,
// 204-214,279-283): the nodes a move touches and the status each one gets.
struct MoveNodes { int na, a1, b1, a2, b2, ns0, ns1, ns2, da1; };
__device__ __forceinline__ MoveNodes move_nodes(int row, int col, int slot, int s0, int d0, int d1)
{
    MoveNodes m;
    m.na = 1; m.a1 = row; m.b1 = col; m.a2 = row; m.b2 = col; m.ns1 = 0; m.ns2 = 0; m.da1 = 0;
    if (slot < 7) {
        m.ns0 = slot == 0 ? S_DIVAC : slot == 1 ? S_PRISTINE : slot <= 3 ? (s0 | (slot - 1))
                : slot <= 5 ? (s0 & ~(slot - 3)) : (s0 ^ 3);
    } else if (slot < 13) {
        const int ri = kring(slot - 7);     // ring position of the target
        const int da = (int)((RING_DA >> (4 * ri)) & 0xF) - 2;
        const int db = (int)((RING_DB >> (4 * ri)) & 0xF) - 2;
        m.a1 = wrap_inc(wrap_dec(row + da, d0), d0);
        m.b1 = wrap_inc(wrap_dec(col + db, d1), d1);
        m.na = 2; m.ns0 = S_PRISTINE; m.ns1 = S_DIVAC; m.da1 = da;
    } else if (slot >= MM_SLOT0) {
        // MoveMonovacancy (direction k, layer L): source loses the layer; the target gains it -- its new status
        // depends on its old one, the caller ORs that in (move_nodes_mm)
        const int ri = kring((slot - MM_SLOT0) >> 1), layer = ((slot - MM_SLOT0) & 1) + 1;
        const int da = (int)((RING_DA >> (4 * ri)) & 0xF) - 2;
        const int db = (int)((RING_DB >> (4 * ri)) & 0xF) - 2;
        m.a1 = wrap_inc(wrap_dec(row + da, d0), d0);
        m.b1 = wrap_inc(wrap_dec(col + db, d1), d1);
        m.na = 2; m.ns0 = s0 & ~layer; m.ns1 = layer; m.da1 = da;
    } else {
        const int o = (slot - 13) & 1;
        const bool create = slot < 15;
        m.a1 = wrap_inc(row + 2, d0); m.b1 = col;                          // role 1: p + (2, 0)
        m.a2 = o ? m.a1 : row;                                              // role 2: p+(0,2) | p+(2,-2)
        m.b2 = o ? wrap_dec(col - 2, d1) : wrap_inc(col + 2, d1);
        m.na = 3;
        m.ns0 = create ? S_TREF + 3 * o : S_DIVAC;
        m.ns1 = create ? S_TREF + 3 * o + 1 : S_DIVAC;
        m.ns2 = create ? S_TREF + 3 * o + 2 : S_DIVAC;
    }
    return m;
}
// the same with the lattice at hand: completes a MoveMonovacancy target (old status | layer)
__device__ __forceinline__ MoveNodes move_nodes_st(const uint8_t *st, int row, int col, int slot, int s0, int d0, int d1)
{
    MoveNodes m = move_nodes(row, col, slot, s0, d0, d1);
    if (slot >= MM_SLOT0) m.ns1 |= st[m.a1 * d1 + m.b1];
    return m;
}
__device__ __forceinline__ void apply_move(uint8_t *st, int d0, int d1, int row, int col, int slot, int s0)
{
    const MoveNodes m = move_nodes_st(st, row, col, slot, s0, d0, d1);
    st[row * d1 + col] = (uint8_t)m.ns0;
    if (m.na > 1) st[m.a1 * d1 + m.b1] = (uint8_t)m.ns1;
    if (m.na > 2) st[m.a2 * d1 + m.b2] = (uint8_t)m.ns2;
}

// ---- in-class pick inside one row, byte lattice ---------------------------------------------------------
// Canonical in-class order: rows ascending; inside a row the class's move slots in slot order; inside a
// slot the sites by column.  A class owns 1, 2 or 6 slots: slot = class_slot_base(c) + bit index.
__device__ __forceinline__ int class_slot_base(int c) {
    //            c: 0  1  2  3  4  5  6   7   8  9  10 11 12     (13..16 MoveMonovacancy: 17)
    return c >= C_MOVE_MONO ? MM_SLOT0 : (int)((0x6442'2FD7'7771'0ull >> (4 * c)) & 0xF);
}
// bit i set: the site carries the class's i-th slot
__device__ __forceinline__ uint32_t site_slot_mask(const uint8_t *st, int a, int b, int d0, int d1, int c) {
    const int s0 = st[a * d1 + b];
    if (c >= C_MOVE_MONO) {
        // MoveMonovacancy kind c - 13: bit 2k + (L - 1) for direction k (neighbors_and_mutuals order), layer L
        if (s0 < 1 || s0 > 3 || ((c - C_MOVE_MONO) >> 1) != (s0 == S_DIVAC)) return 0u;
        const int am = wrap_dec(a - 1, d0), ap = wrap_inc(a + 1, d0);
        const int bm = wrap_dec(b - 1, d1), bp = wrap_inc(b + 1, d1);
        const int ring[6] = {st[am * d1 + b], st[a * d1 + bm], st[ap * d1 + bm], st[ap * d1 + b], st[a * d1 + bp],
                             st[am * d1 + bp]};
        const int want_nonpristine = (c - C_MOVE_MONO) & 1;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int dst = ring[kring(k)];
#pragma unroll
            for (int layer = 1; layer <= 2; layer++)
                if ((s0 & layer) && dst <= 3 && !(dst & layer) && (int)(dst != 0) == want_nonpristine)
                    out |= 1u << (2 * k + layer - 1);
        }
        return out;
    }
    if (c >= C_MOVE_DIV && c < C_MOVE_DIV + 4) {
        if (s0 != S_DIVAC) return 0u;
        const int am = wrap_dec(a - 1, d0), ap = wrap_inc(a + 1, d0);
        const int bm = wrap_dec(b - 1, d1), bp = wrap_inc(b + 1, d1);
        RingMasks m = ring_masks(st[am * d1 + b], st[a * d1 + bm], st[ap * d1 + bm], st[ap * d1 + b],
                                 st[a * d1 + bp], st[am * d1 + bp]);
        const uint32_t km = kind_mask(m, c - C_MOVE_DIV);          // by ring position
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) out |= ((km >> kring(k)) & 1u) << k;   // by direction k = slot - 7
        return out;
    }
    if (c == C_CREATE_TREF) {
        if (s0 != S_DIVAC) return 0u;
        const int ap2 = wrap_inc(a + 2, d0), bp2 = wrap_inc(b + 2, d1), bm2 = wrap_dec(b - 2, d1);
        const uint32_t t20 = st[ap2 * d1 + b] == S_DIVAC;
        return (t20 & (uint32_t)(st[a * d1 + bp2] == S_DIVAC)) | ((t20 & (uint32_t)(st[ap2 * d1 + bm2] == S_DIVAC)) << 1);
    }
    switch (c) {
        case C_CREATE_DIV: return s0 == 0;
        case C_FILL_DIV: return s0 == 3;
        case C_DESTROY_TREF: return (uint32_t)(s0 == 4) | ((uint32_t)(s0 == 7) << 1);       // slots 15, 16
        case C_CMONO_FROM_DOUBLE: return s0 == 0 ? 3u : 0u;                                // slots 2, 3
        case C_CMONO_MAKE_EMPTY: return (uint32_t)(s0 == 2) | ((uint32_t)(s0 == 1) << 1);   // slot 2 (layer 1), 3
        case C_FMONO_MAKE_DOUBLE: return (uint32_t)(s0 == 1) | ((uint32_t)(s0 == 2) << 1);  // slot 4 (layer 1), 5
        case C_FMONO_FROM_EMPTY: return s0 == 3 ? 3u : 0u;                                 // slots 4, 5
        default: return s0 == 1 || s0 == 2;                                                // FlipMonovacancy
    }
}
// the r-th move (0-based) of class c in row `row`: column and slot.  Two passes over the row: per-slot
// totals, then a rank search among the sites of the chosen slot.  NSL = slots a class can own: 6 for the
// reference's own rules, 12 with MoveMonovacancy (direction x layer).
template <int NSL = 6>
__device__ __forceinline__ void find_in_row_bytes(const uint8_t *st, int d0, int d1, int row, int c, uint32_t r,
                                                  int lane, int &col, int &slot) {
    constexpr int NP = NSL / 2;                              // packed pairs of 16-bit counters
    if (d1 > 64 && d1 <= 1024) {
        // Wide rows: each lane owns a contiguous run of <= 32 sites and evaluates every site ONCE, keeping
        // one bit mask per slot (a lane's sites share cache lines, and no cross-lane step sits between the
        // loads, so the latency of an HBM/L2-resident lattice is paid once, not once per 32-site chunk).
        const int per = (d1 + 31) >> 5;
        const int b0 = lane * per;
        uint32_t m[NSL];
#pragma unroll
        for (int q = 0; q < NSL; q++) m[q] = 0u;
        for (int i = 0; i < per; i++) {
            const int b = b0 + i;
            const uint32_t sm = b < d1 ? site_slot_mask(st, row, b, d0, d1, c) : 0u;
#pragma unroll
            for (int q = 0; q < NSL; q++) m[q] |= ((sm >> q) & 1u) << i;
        }
        uint32_t cnt[NSL];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const uint32_t a = __reduce_add_sync(FULL, __popc(m[2 * q]) | (__popc(m[2 * q + 1]) << 16));
            cnt[2 * q] = a & 0xFFFFu; cnt[2 * q + 1] = a >> 16;
        }
        int j = 0;
#pragma unroll
        for (int q = 0; q < NSL - 1; q++)
            if (j == q && r >= cnt[q]) { r -= cnt[q]; j = q + 1; }
        uint32_t mine = m[0];
#pragma unroll
        for (int q = 1; q < NSL; q++) mine = j == q ? m[q] : mine;
        int L = 0;
        warp_rank_search(__popc(mine), r, L, lane);           // r becomes the rank inside lane L's run
        uint32_t mL = __shfl_sync(FULL, mine, L);
        for (uint32_t i = 0; i < r; i++) mL &= mL - 1;
        col = L * per + __ffs(mL) - 1;
        slot = class_slot_base(c) + j;
        return;
    }
    uint32_t acc[NP];                                        // NSL 16-bit per-slot counters
#pragma unroll
    for (int q = 0; q < NP; q++) acc[q] = 0u;
    for (int base = 0; base < d1; base += 32) {
        const int b = base + lane;
        const uint32_t m = b < d1 ? site_slot_mask(st, row, b, d0, d1, c) : 0u;
#pragma unroll
        for (int q = 0; q < NP; q++) acc[q] += ((m >> (2 * q)) & 1u) | (((m >> (2 * q + 1)) & 1u) << 16);
    }
    int j = 0;
    {
        uint32_t cnt[NSL];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const uint32_t a = __reduce_add_sync(FULL, acc[q]);
            cnt[2 * q] = a & 0xFFFFu; cnt[2 * q + 1] = a >> 16;
        }
#pragma unroll
        for (int q = 0; q < NSL - 1; q++)
            if (j == q && r >= cnt[q]) { r -= cnt[q]; j = q + 1; }
    }
    col = 0;
    for (int base = 0; base < d1; base += 32) {
        const int b = base + lane;
        const uint32_t v = b < d1 ? ((site_slot_mask(st, row, b, d0, d1, c) >> j) & 1u) : 0u;
        int L;
        if (warp_rank_search(v, r, L, lane)) { col = base + L; break; }
    }
    slot = class_slot_base(c) + j;
}

// per-row class counts of row `a`, as 7 (9 with MoveMonovacancy) pair-words (two uint16 fields each)
template <bool MM = false>
__device__ __forceinline__ void row_counts(const uint8_t *st, int a, int d0, int d1, uint32_t out[RcWords<MM>::value]) {
    SmemStatus S{st};
    uint32_t A0 = 0, B0 = 0, A1 = 0, B1 = 0, A2 = 0, B2 = 0, A3 = 0, B3 = 0, A4 = 0;
    for (int b = 0; b < d1; b++) {
        F4 f = eval_site<MM>(S, a, b, d0, d1);
        A0 += f.w0 & 0x00FF00FFu; B0 += (f.w0 >> 8) & 0x00FF00FFu;
        A1 += f.w1 & 0x00FF00FFu; B1 += (f.w1 >> 8) & 0x00FF00FFu;
        A2 += f.w2 & 0x00FF00FFu; B2 += (f.w2 >> 8) & 0x00FF00FFu;
        A3 += f.w3 & 0x00FF00FFu;
        if (MM) { B3 += (f.w3 >> 8) & 0x00FF00FFu; A4 += f.w4 & 0x00FF00FFu; }
    }
    out[0] = A0; out[1] = B0; out[2] = A1; out[3] = B1; out[4] = A2; out[5] = B2; out[6] = A3;
    if (MM) { out[7] = B3; out[8] = A4; }
}

// GLOBAL_STATE = true: the lattice stays in HBM/L2 (lattices too large for shared memory, e.g.
// 1024x1024 = 1 MiB per replica, BASELINE config 5); only the row-count hierarchy is in shared memory.
// MM = true: MoveMonovacancy's four classes are maintained as well (17 classes, 9 row-count words, 32 weight lanes).
template <int TD0, int TD1, bool GLOBAL_STATE = false, bool MM = false>
__global__ void __launch_bounds__(256, 4) kmc_replica_kernel(const ReplicaParams p)
{
    constexpr int RCW = RcWords<MM>::value;          // row-count words per row
    constexpr int NCX = MM ? NCA : NC;               // classes this build maintains
    constexpr int NL = MM ? 32 : 16;                 // lanes holding a class weight
    const int d0 = TD0 ? TD0 : p.d0;
    const int d1 = TD1 ? TD1 : p.d1;
    const int n = d0 * d1;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int rep = blockIdx.x * warps_per_block + warp;
    if (rep >= p.n_replicas) return;       // whole warp exits together; no block barriers are used

    extern __shared__ uint4 smem_raw[];
    uint8_t *gst = p.status + (size_t)rep * n;
    uint8_t *st;
    uint32_t *rc;
    if (GLOBAL_STATE) {
        st = gst;
        rc = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)warp * replica_rc_bytes(d0, MM));
    } else {
        st = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)warp * replica_warp_bytes(d0, d1, MM);
        rc = reinterpret_cast<uint32_t *>(st + (((size_t)n + 15) & ~(size_t)15));
    }
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(rc);

    // ---- load the lattice ------------------------------------------------------------------
    if (GLOBAL_STATE) {
        // nothing to stage
    } else if ((n & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(gst);
        uint4 *dst = reinterpret_cast<uint4 *>(st);
        for (int i = lane; i < (n >> 4); i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < n; i += 32) st[i] = gst[i];
    }
    __syncwarp();

    // ---- build the row hierarchy (a full enumeration, as Rule.initialize_moves does) --------
    if (GLOBAL_STATE && p.hier && p.hier_valid) {
        const uint32_t *src = p.hier + (size_t)rep * d0 * RCW;
        for (int i = lane; i < d0 * RCW; i += 32) rc[i] = src[i];
    } else {
        for (int a = lane; a < d0; a += 32) {
            uint32_t w[RCW];
            row_counts<MM>(st, a, d0, d1, w);
#pragma unroll
            for (int j = 0; j < RCW; j++) rc[a * RCW + j] = w[j];
        }
    }
    __syncwarp();

    // per-lane class state: lane c < 13 owns class c; lanes 13..15 are permanent zero weights
    const int myc = lane < NCX ? lane : 0;
    const double rate = (lane < NCX && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NCX && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    int tot = 0;
    if (lane < NCX) {
        const int f = rc_field(lane);
        for (int a = 0; a < d0; a++) tot += rc16[a * (2 * RCW) + f];
    }
    unsigned long long hist = 0, hops = 0;
    unsigned long long n_ties = 0;
    PickState pstate = pick_state_init_n<NL>(lane);
    double tclock = p.clock ? p.clock[rep] : 0.0;          // continues across launches in event order
    uint32_t *const tracer = p.tracer ? p.tracer + (size_t)rep * n : nullptr;

    // per-lane role in the refresh: node index and one of 10 offsets
    const int my_i = lane / 10, my_o = lane % 10;
    //            o:   0   1   2   3   4   5   6   7   8   9
    // da              0  -1   0   1   1   0  -1  -2   0  -2
    // db              0   0  -1  -1   0   1   1   0  -2   2
    const int my_da = (int)((REFRESH_DA >> (4 * my_o)) & 0xF) - 2;
    const int my_db = (int)((REFRESH_DB >> (4 * my_o)) & 0xF) - 2;

    const unsigned long long step_base = p.steps_done[rep];
    double mu1 = 0.0, mu2 = 0.0;           // this lane's prefetched draws
    long long t = 0;
    uint32_t flag = 0;

    for (; t < p.nsteps; t++) {
        // ---- draws: a block of 32 events at a time, one event per lane ----------------------
        if ((t & 31) == 0) {
            long long tt = t + lane;
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

        // ---- 1. weights ------------------------------------------------------------------------
        const double w = __dmul_rn((double)tot, rate);       // lanes >= 13: tot = 0, rate = 0

        // ---- 2. class choice -------------------------------------------------------------------
        const ClassPick pick = pick_class<NL>(w, pos, pstate, u1, lane);
        if (pick.zero) {                                        // kmc.py:31-32: nothing can happen
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
        const double total = pick.total;
        const uint32_t tie = pick.tie;
        if (p.clock) tclock = __dadd_rn(tclock, kmc_time_step(p.clock_mode, total, p.seed, p.replica0 + (uint32_t)rep,
                                                            step_base + (unsigned long long)t));

        // ---- event record, part 1 (counts before the event) -------------------------------------
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) +
                                  ((size_t)rep * (size_t)p.nsteps + (size_t)t);
            if (lane < NCA) rec->counts[lane] = lane < NCX ? (uint32_t)tot : 0u;
        }

        // ---- 3. in-class pick --------------------------------------------------------------------
        const int cnt = __shfl_sync(FULL, tot, csel);
        uint32_t r;
        {
            long long rr = (long long)__dmul_rn(u2, (double)cnt);
            if (rr > (long long)cnt - 1) rr = (long long)cnt - 1;
            r = (uint32_t)rr;
        }
        // 3a. row: each lane owns a contiguous block of rows (2 per pass; 8 when the lattice is tall)
        int row = 0;
        {
            const int f = rc_field(csel);
            const int K = d0 > 256 ? 8 : 2;
            for (int base = 0; base < d0; base += 32 * K) {
                const int r0 = base + K * lane;
                uint32_t sum = 0;
                for (int q = 0; q < K; q++)
                    if (r0 + q < d0) sum += rc16[(r0 + q) * (2 * RCW) + f];
                int L;
                if (warp_rank_search(sum, r, L, lane)) {
                    // walk the owner lane's rows (uniform: every lane reads the same entries)
                    row = base + K * L;
                    for (int q = 0; q < K - 1; q++) {
                        const uint32_t v = rc16[row * (2 * RCW) + f];
                        if (r < v) break;
                        r -= v; row++;
                    }
                    break;
                }
            }
        }
        // 3b/3c. site within the row and slot (row, slot, column order)
        int col, slot;
        find_in_row_bytes<MM ? 12 : 6>(st, d0, d1, row, csel, r, lane, col, slot);
        const int site = row * d1 + col;
        const int s0 = st[site];

        // ---- event record, part 2 ------------------------------------------------------------------
        if (p.log_mode != KMC_LOG_NONE && lane == 0) {
            const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)t;
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
                rec->flags = tie ? 1 : 0; rec->total_rate = total; rec->pad = 0;
            } else {
                kmc_event_compact *rec = reinterpret_cast<kmc_event_compact *>(p.events) + idx;
                kmc_event_compact e;
                e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
                e.flags = tie ? 1 : 0; e.total_rate = total;
                *rec = e;
            }
        }
        if (lane == csel) hist++;
        if (lane == slot - 7) hops++;                          // lanes 0..5: divacancy hops per direction
        if (tie && lane == 0) n_ties++;

        // ---- 4. the move: touched nodes and their new status (uniform) --------------------------------
        const MoveNodes mv = move_nodes_st(st, row, col, slot, s0, d0, d1);
        const int na = mv.na, a1 = mv.a1, b1 = mv.b1, a2 = mv.a2, b2 = mv.b2;
        const int ns0 = mv.ns0, ns1 = mv.ns1, ns2 = mv.ns2;
        if (tracer && lane == 0) tracer_update(tracer, slot, na, site, ns0, a1 * d1 + b1, ns1, a2 * d1 + b2, ns2, s0);

        // ---- 5. neighbourhood refresh -----------------------------------------------------------------
        const bool act = my_i < na;
        const int ba = my_i == 0 ? row : my_i == 1 ? a1 : a2;
        const int bb = my_i == 0 ? col : my_i == 1 ? b1 : b2;
        int xa = ba + my_da; xa = xa < 0 ? xa + d0 : (xa >= d0 ? xa - d0 : xa);
        int xb = bb + my_db; xb = xb < 0 ? xb + d1 : (xb >= d1 ? xb - d1 : xb);
        const int xsite = xa * d1 + xb;
        // the same site can be reached from two touched nodes: keep one lane per distinct site
        const uint32_t same = __match_any_sync(FULL, act ? xsite : (int)(0x40000000 | lane));
        const bool leader = act && ((__ffs(same) - 1) == lane);

        SmemStatus S{st};
        F4 fb = {0u, 0u, 0u, 0u, 0u};
        if (leader) fb = eval_site<MM>(S, xa, xb, d0, d1);
        __syncwarp();
        if (lane == 0) {
            st[site] = (uint8_t)ns0;
            if (na > 1) st[a1 * d1 + b1] = (uint8_t)ns1;
            if (na > 2) st[a2 * d1 + b2] = (uint8_t)ns2;
        }
        __syncwarp();
        F4 fa = {0u, 0u, 0u, 0u, 0u};
        if (leader) fa = eval_site<MM>(S, xa, xb, d0, d1);

        // per-byte (after - before), biased by 0x80 so no borrow crosses bytes; then spread to
        // 16-bit pair-words whose integer value is hi*65536 + lo with signed lo, hi
        uint32_t dw[RCW];
        {
            const uint32_t e0 = (fa.w0 | 0x80808080u) - fb.w0, e1 = (fa.w1 | 0x80808080u) - fb.w1;
            const uint32_t e2 = (fa.w2 | 0x80808080u) - fb.w2, e3 = (fa.w3 | 0x80808080u) - fb.w3;
            dw[0] = (e0 & 0x00FF00FFu) - 0x00800080u; dw[1] = ((e0 >> 8) & 0x00FF00FFu) - 0x00800080u;
            dw[2] = (e1 & 0x00FF00FFu) - 0x00800080u; dw[3] = ((e1 >> 8) & 0x00FF00FFu) - 0x00800080u;
            dw[4] = (e2 & 0x00FF00FFu) - 0x00800080u; dw[5] = ((e2 >> 8) & 0x00FF00FFu) - 0x00800080u;
            dw[6] = (e3 & 0x00FF00FFu) - 0x00800080u;
            if (MM) {
                const uint32_t e4 = (fa.w4 | 0x80808080u) - fb.w4;
                dw[7] = ((e3 >> 8) & 0x00FF00FFu) - 0x00800080u;
                dw[8] = (e4 & 0x00FF00FFu) - 0x00800080u;
            }
        }
        // row counts: the few lanes with a change add their packed deltas
#pragma unroll
        for (int j = 0; j < RCW; j++)
            if (dw[j] != 0u) atomicAdd(&rc[xa * RCW + j], dw[j]);
        // class totals: warp-sum of the packed deltas, then lane c unpacks its own field
        int keep = 0;
#pragma unroll
        for (int j = 0; j < RCW; j++) {
            const int sj = (int)__reduce_add_sync(FULL, dw[j]);
            if (lane == j) keep = sj;
        }
        if (true) {
            const int fld = rc_field(myc);                         // 2*pairword + half
            const int v = __shfl_sync(FULL, keep, fld >> 1);
            const int lo = (int)(short)(v & 0xFFFF);
            const int hi = (v - lo) >> 16;
            if (lane < NCX) tot += (fld & 1) ? hi : lo;
        }
        __syncwarp();
    }

    // ---- epilogue ------------------------------------------------------------------------------------
    const long long done = t;
    if (p.validate) {
        // KmcSim.validate (sim.py:159-168): incremental tables vs a fresh enumeration
        bool bad = false;
        for (int a = lane; a < d0; a += 32) {
            uint32_t w[RCW];
            row_counts<MM>(st, a, d0, d1, w);
#pragma unroll
            for (int j = 0; j < RCW; j++) bad |= (rc[a * RCW + j] != w[j]);
        }
        if (lane < NCX) {
            const int f = rc_field(lane);
            int fresh = 0;
            for (int a = 0; a < d0; a++) fresh += rc16[a * (2 * RCW) + f];
            bad |= (fresh != tot);
        }
        if (__any_sync(FULL, bad)) flag |= FLAG_VALIDATE;
    }
    if (GLOBAL_STATE) {
        // the lattice was updated in place; keep the hierarchy for the next launch
        if (p.hier) {
            uint32_t *dst = p.hier + (size_t)rep * d0 * RCW;
            for (int i = lane; i < d0 * RCW; i += 32) dst[i] = rc[i];
        }
    } else if ((n & 15) == 0) {
        uint4 *dst = reinterpret_cast<uint4 *>(gst);
        const uint4 *src = reinterpret_cast<const uint4 *>(st);
        for (int i = lane; i < (n >> 4); i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < n; i += 32) gst[i] = st[i];
    }
    if (lane < NCX) p.hist[(size_t)rep * NCA + lane] += hist;
    if (lane < 6) p.hops[(size_t)rep * 6 + lane] += hops;
    if (lane == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (p.clock) p.clock[rep] = tclock;
        p.ties[rep] += n_ties;
        if (flag) { p.flags[rep] |= flag; atomicOr(p.any_flag, flag); }
    }
}

}  // namespace kmc
