// replica_hw_kernel.cuh -- K3 + K4 on the bit-sliced lattice with TWO replicas per warp.
//
// Same state and the same per-event algorithm as replica_bp_kernel.cuh (bit-sliced lattice, pre-rotated
// Z/D planes, class-major row-count table, recompute-and-replace refresh); what changes is the mapping:
// a replica is owned by a HALF-warp (16 lanes).  Almost all of an event is warp-uniform work that needs at
// most 16 lanes (class choice over 13+3 weight lanes, rank bookkeeping, move decoding, plane and own-status
// updates on 10 / 13 lanes, six direction lanes), so a full warp per replica spends most issue slots on
// instructions whose upper 16 lanes are idle.  With two replicas per warp those instructions serve two
// events.  Every shuffle / vote / redux names only its own half (mask 0x0000FFFF or 0xFFFF0000, width 16),
// so the halves may diverge freely (different classes, loop trip counts, early stop).
//
// Lane roles inside a half (hl = lane & 15):
//   hl < 13   owns class hl (total count, rate, own-status row counts)      hl < 10  maintains bit plane hl
//   hl < 6    direction hl of the MoveDivacancy masks during the site search, CreateTrefoil anchor rows
//   hl < 12   refresh: direction hl % 6 of band rows 2*pass + hl / 6
#pragma once
#include <type_traits>
#include "replica_bp_kernel.cuh"

namespace kmc {

// Shared memory is what limits how many half-warps fit on an SM, so this kernel keeps only the six
// type planes (3 KiB at 64 rows) and rotates neighbour rows on the fly (two funnel shifts per rotate)
// instead of holding pre-rotated copies: 4.6 KiB per replica -> 48 replicas = 24 warps per SM.
enum { HW_PLANES = N_TYPES };      // plane index == plane type: T_Z, T_D, T_M1, T_M2, T_A4, T_A7

__host__ __device__ inline size_t replica_hw_bytes(int d0) {      // per replica
    const size_t d0p = ((size_t)d0 + 3) & ~(size_t)3;
    const size_t planes = (size_t)HW_PLANES * d0 * 8;
    const size_t rc = (NC * d0p * 2 + 15) & ~(size_t)15;
    return planes + rc;
}

// word whose bit b is bit (b + db) of x, db in {-1, 0, +1}: (sh, sw) = (0, no) / (1, no) / (31, swap halves)
template <int TD1>
__device__ __forceinline__ u64 rot_site(u64 x, int db, int sh, bool sw, const RowBits<TD1> &rb) {
    if (TD1 == 64) {
        const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
        const uint32_t a = sw ? hi : lo, b = sw ? lo : hi;
        return (u64)__funnelshift_l(b, a, sh) | ((u64)__funnelshift_l(a, b, sh) << 32);
    }
    return rb.rotl(x, rb.amount(db));
}
struct HwDir { int da_n, db_n, sh_n, da_e, db_e, sh_e, da_f, db_f, sh_f; };
__device__ __forceinline__ HwDir hw_dir_const(int k) {
    HwDir c;
    c.da_n = nib(NBR_DA, k);  c.db_n = nib(NBR_DB, k);  c.sh_n = c.db_n == 0 ? 0 : (c.db_n < 0 ? 1 : 31);
    c.da_e = nib(NEAR_DA, k); c.db_e = nib(NEAR_DB, k); c.sh_e = c.db_e == 0 ? 0 : (c.db_e < 0 ? 1 : 31);
    c.da_f = nib(FAR_DA, k);  c.db_f = nib(FAR_DB, k);  c.sh_f = c.db_f == 0 ? 0 : (c.db_f < 0 ? 1 : 31);
    return c;
}
template <int TD0, int TD1>
__device__ __forceinline__ void hw_dir_masks(const u64 *PL, int d0, int a, const HwDir &c, const RowBits<TD1> &rb,
                                             u64 &present, u64 &near, u64 &far) {
    present = PL[T_D * d0 + a] & rot_site(PL[T_Z * d0 + wrapT<TD0>(a + c.da_n, d0)], c.db_n, c.sh_n, c.db_n > 0, rb);
    near = rot_site(PL[T_D * d0 + wrapT<TD0>(a + c.da_e, d0)], c.db_e, c.sh_e, c.db_e > 0, rb);
    far = rot_site(PL[T_D * d0 + wrapT<TD0>(a + c.da_f, d0)], c.db_f, c.sh_f, c.db_f > 0, rb);
}
template <int TD0, int TD1>
__device__ __forceinline__ void hw_trefoil_masks(const u64 *PL, int d0, int a, const RowBits<TD1> &rb, u64 &up, u64 &dn) {
    const u64 Da = PL[T_D * d0 + a];
    const u64 Db = PL[T_D * d0 + wrapT<TD0>(a + 2, d0)];
    const u64 both = Da & Db;
    up = both & rb.rotl(Da, rb.amount(2));
    dn = both & rb.rotl(Db, rb.amount(-2));
}
template <int TD0, int TD1>
__device__ __forceinline__ void hw_row_counts(const u64 *PL, int d0, int a, const RowBits<TD1> &rb, uint32_t out[NC]) {
    const uint32_t pz = __popcll(PL[T_Z * d0 + a]), pd = __popcll(PL[T_D * d0 + a]);
    const uint32_t pm = __popcll(PL[T_M1 * d0 + a] | PL[T_M2 * d0 + a]);
    const uint32_t pa = __popcll(PL[T_A4 * d0 + a] | PL[T_A7 * d0 + a]);
    out[C_CREATE_DIV] = pz; out[C_FILL_DIV] = pd; out[C_DESTROY_TREF] = pa;
    out[C_CMONO_FROM_DOUBLE] = 2 * pz; out[C_CMONO_MAKE_EMPTY] = pm; out[C_FMONO_MAKE_DOUBLE] = pm;
    out[C_FMONO_FROM_EMPTY] = 2 * pd; out[C_FLIP] = pm;
    uint32_t k01 = 0, k23 = 0;
    for (int k = 0; k < 6; k++) {
        const HwDir c = hw_dir_const(k);
        u64 p, n, f; hw_dir_masks<TD0>(PL, d0, a, c, rb, p, n, f);
        uint32_t a01, a23; kind_counts(p, n, f, a01, a23);
        k01 += a01; k23 += a23;
    }
    out[C_MOVE_DIV] = k01 & 0xFFFF; out[C_MOVE_DIV + 1] = k01 >> 16;
    out[C_MOVE_DIV + 2] = k23 & 0xFFFF; out[C_MOVE_DIV + 3] = k23 >> 16;
    u64 up, dn; hw_trefoil_masks<TD0>(PL, d0, a, rb, up, dn);
    out[C_CREATE_TREF] = __popcll(up) + __popcll(dn);
}

// Two execution modes of the same event code:
//   LOCK = true   both halves of the warp run in lockstep: every data-dependent branch that contains a
//                 shuffle or vote is taken if EITHER half needs it (its effects are predicated per half), so
//                 all sync intrinsics can name the full warp (plain SHFL / VOTE, no WARPSYNC prologue);
//   LOCK = false  a half runs on its own (the partner stopped on a zero total rate, or does not exist):
//                 intrinsics name only the half's 16 lanes.
struct Half { int hl, shift; uint32_t mask; };
template <bool LOCK> struct HalfSync {
    __device__ __forceinline__ static uint32_t m(const Half &hw) { return LOCK ? FULL : hw.mask; }
    // "does any half that shares my instruction stream need this?"  (pred is uniform inside a half)
    __device__ __forceinline__ static bool any(const Half &hw, bool pred) { return LOCK ? __any_sync(FULL, pred) : pred; }
    __device__ __forceinline__ static uint32_t ballot(const Half &hw, bool pred) {
        return (__ballot_sync(m(hw), pred) >> hw.shift) & 0xFFFFu;
    }
    // min / max over the halves that share the instruction stream (x is uniform inside a half)
    __device__ __forceinline__ static int umin(const Half &hw, int x) { return LOCK ? (int)__reduce_min_sync(FULL, (unsigned)x) : x; }
    __device__ __forceinline__ static int umax(const Half &hw, int x) { return LOCK ? (int)__reduce_max_sync(FULL, (unsigned)x) : x; }
    __device__ __forceinline__ static void sync(const Half &hw) { __syncwarp(m(hw)); }
};
#define HSHFL(v, src) __shfl_sync(HS::m(hw), (v), (src), 16)
#define HSHFL_UP(v, d) __shfl_up_sync(HS::m(hw), (v), (d), 16)
#define HSHFL_DOWN(v, d) __shfl_down_sync(HS::m(hw), (v), (d), 16)
// minimum over the 16 lanes of a half (4 xor steps; REDUX has no segment width)
template <class HS> __device__ __forceinline__ uint32_t half_min(const Half &hw, uint32_t v) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) v = min(v, __shfl_xor_sync(HS::m(hw), v, o, 16));
    return v;
}

// weighted_choice over the 16 lanes of a half (see pick_class in replica_kernel.cuh; same arithmetic)
template <bool LOCK>
__device__ __forceinline__ ClassPick pick_class_hw(double w, int pos, PickState &st, double u1, const Half &hw)
{
    typedef HalfSync<LOCK> HS;
    const int hl = hw.hl;
    double ws = HSHFL(w, st.sc);
    const int ps = HSHFL(pos, st.sc);
    int first = (int)half_min<HS>(hw, (w != st.wprev) ? (unsigned)st.spos : 16u);
    st.wprev = w;
    {
        const double wp = HSHFL_UP(ws, 1);
        const int pp = HSHFL_UP(ps, 1);
        const bool ok = (hl == 0) || (wp < ws) || (wp == ws && pp < ps);
        const bool resort = st.need_sort || HS::ballot(hw, !ok) != 0;
        if (HS::any(hw, resort)) {
            // re-deriving the order of a half that did not need it reproduces its order exactly
            int rank = 0;
            for (int j = 0; j < 16; j++) {
                const double wj = HSHFL(w, j);
                const int pj = HSHFL(pos, j);
                rank += (wj < w) || (wj == w && pj < pos);
            }
            st.spos = rank;
            for (int j = 0; j < 16; j++) {
                const int rj = HSHFL(rank, j);
                if (rj == hl) st.sc = j;
            }
            st.need_sort = false;
            ws = HSHFL(w, st.sc);
            if (resort) first = 0;
        }
    }
    ClassPick out;
    const int z = __popc(HS::ballot(hw, ws == 0.0));
    out.zero = (z == 16);
    const int start = first > z ? first : z;
    double acc = HSHFL(st.cum, (start + 15) & 15);
    if (start == z) acc = 0.0;
    const int k0 = HS::umin(hw, start);                        // lockstep: the earlier of the two halves
    for (int k = k0; k < 16; k++) {
        const double wk = HSHFL(ws, k);
        if (k >= start) {
            acc = __dadd_rn(acc, wk);                          // util.py:65-68: sequential sums
            if (hl == k) st.cum = acc;
        }
    }
    const double last = HSHFL(st.cum, 15);
    if (start == 16) acc = last;
    out.total = acc;
    const double x = __dmul_rn(acc, u1);                       // random.uniform(0, total), kmc.py:49
    const bool live = hl >= z;
    const uint32_t ge = HS::ballot(hw, live && st.cum >= x);   // bisect_left, kmc.py:50
    const int isel = ge ? (__ffs(ge) - 1) : 15;
    out.tie = HS::ballot(hw, live && st.cum == x);
    out.csel = HSHFL(st.sc, isel);
    return out;
}

// column of the (idx)-th set bit of a row mask: lane hl counts the set bits below site 4*hl + 4
template <class HS>
__device__ __forceinline__ int nth_set_bit_hw(u64 m, uint32_t idx, const Half &hw, uint32_t low_lo, uint32_t low_hi) {
    const uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32);
    const uint32_t c = __popc(mlo & low_lo) + __popc(mhi & low_hi);
    const int L = (__ffs(HS::ballot(hw, c > idx)) - 1) & 15;
    const uint32_t below = HSHFL(c, (L + 15) & 15);
    uint32_t nibble = ((L & 8 ? mhi : mlo) >> ((4 * L) & 31)) & 0xFu;
    const uint32_t rem = idx - (L ? below : 0u);
    for (uint32_t i = 0; i < (rem & 3u); i++) nibble &= nibble - 1;    // rem <= 3, uniform inside the half
    return 4 * L + ((__ffs(nibble) - 1) & 3);
}

template <int TD0, int TD1>
__global__ void __launch_bounds__(256, 3) kmc_replica_hw_kernel(const ReplicaParams p)
{
    const int d0 = TD0 ? TD0 : p.d0;
    const int d1 = TD1 ? TD1 : p.d1;
    const int n = d0 * d1;
    const int d0p = (d0 + 3) & ~3;
    const int lane = threadIdx.x & 31;
    Half hw;
    hw.hl = lane & 15; hw.shift = lane & 16; hw.mask = 0xFFFFu << hw.shift;
    const int hl = hw.hl;
    const int slot_in_block = (threadIdx.x >> 4);                       // half-warp index inside the block
    const int rep = blockIdx.x * (blockDim.x >> 4) + slot_in_block;
    // lockstep needs both halves of the warp to own a replica
    const bool pair_ok = __all_sync(FULL, rep < p.n_replicas);
    if (rep >= p.n_replicas) return;       // a whole half exits; the other half then runs on its own
    const RowBits<TD1> rb(d1);

    extern __shared__ uint4 smem_raw[];
    u64 *PL = reinterpret_cast<u64 *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)slot_in_block * replica_hw_bytes(d0));
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(PL + HW_PLANES * d0);      // [NC][d0p]
    const u64 *rc64 = reinterpret_cast<const u64 *>(rc16);

    // ---- load: status bytes -> bit planes (one row per lane at a time) --------------------------
    uint8_t *gst = p.status + (size_t)rep * n;
    for (int a = hl; a < d0; a += 16) {
        u64 w[N_TYPES] = {0, 0, 0, 0, 0, 0};
        if ((d1 & 3) == 0) {
            const uint32_t *g32 = reinterpret_cast<const uint32_t *>(gst + a * d1);
            for (int b4 = 0; b4 < (d1 >> 2); b4++) {
                const uint32_t v = g32[b4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int pl = status_plane((v >> (8 * i)) & 0xF);
#pragma unroll
                    for (int q = 0; q < N_TYPES; q++) w[q] |= (u64)(pl == q) << (4 * b4 + i);
                }
            }
        } else {
            for (int b = 0; b < d1; b++) {
                const int pl = status_plane(gst[a * d1 + b]);
#pragma unroll
                for (int q = 0; q < N_TYPES; q++) w[q] |= (u64)(pl == q) << b;
            }
        }
#pragma unroll
        for (int q = 0; q < N_TYPES; q++) PL[q * d0 + a] = w[q];
    }
    __syncwarp(hw.mask);
    // ---- build the row hierarchy (full enumeration, as Rule.initialize_moves does) ---------------
    for (int a = hl; a < d0p; a += 16) {
        uint32_t cnt[NC];
        if (a < d0) hw_row_counts<TD0>(PL, d0, a, rb, cnt);
#pragma unroll
        for (int c = 0; c < NC; c++) rc16[c * d0p + a] = a < d0 ? (uint16_t)cnt[c] : (uint16_t)0;
    }
    __syncwarp(hw.mask);

    // per-lane class state: half-lane c < 13 owns class c; 13..15 are permanent zero weights
    const int myc = hl < NC ? hl : 0;
    const double rate = (hl < NC && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (hl < NC && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + hl;
    int tot = 0;
    if (hl < NC)
        for (int a = 0; a < d0; a++) tot += rc16[hl * d0p + a];
    const uint32_t my_g1 = (hl < NC && !(hl >= C_MOVE_DIV && hl <= C_CREATE_TREF)) ? g1_const(hl) : 0u;
    unsigned long long hist = 0, n_ties = 0;
    PickState pstate = pick_state_init(hl);
    double tclock = 0.0;
    // refresh roles: half-lanes 0..11 = direction hl % 6 of band rows 2*pass + hl / 6
    const int my_g = hl / 6, my_k = hl % 6;
    HwDir my_dir = hw_dir_const(my_k);
    asm volatile("" : "+r"(my_dir.da_n), "+r"(my_dir.db_n), "+r"(my_dir.sh_n), "+r"(my_dir.da_e), "+r"(my_dir.db_e),
                      "+r"(my_dir.sh_e), "+r"(my_dir.da_f), "+r"(my_dir.db_f), "+r"(my_dir.sh_f));
    // this lane's mask of the sites below column 4*hl + 4 (for popcount-select)
    uint32_t my_low_lo = hl >= 7 ? 0xFFFFFFFFu : ((16u << (4 * hl)) - 1u);
    uint32_t my_low_hi = hl < 8 ? 0u : (hl == 15 ? 0xFFFFFFFFu : ((16u << (4 * (hl - 8))) - 1u));
    asm volatile("" : "+r"(my_low_lo), "+r"(my_low_hi));

    const unsigned long long step_base = p.steps_done[rep];
    double mu1 = 0.0, mu2 = 0.0;
    long long t = 0;
    uint32_t flag = 0;

    // the event loop, in lockstep (both halves, full-warp intrinsics) or alone (half-warp intrinsics)
    auto run = [&](auto lock_tag) {
    constexpr bool LOCK = decltype(lock_tag)::value;
    typedef HalfSync<LOCK> HS;
    for (; t < p.nsteps; t++) {
        // ---- draws: a block of 16 events at a time, one event per half-lane ---------------------
        if ((t & 15) == 0) {
            const long long tt = t + hl;
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
        const double u1 = HSHFL(mu1, (int)(t & 15));
        const double u2 = HSHFL(mu2, (int)(t & 15));

        // ---- 1./2. weights and class choice (sim.py:120-122, kmc.py:21-54) ----------------------
        const double w = __dmul_rn((double)tot, rate);
        const ClassPick pick = pick_class_hw<LOCK>(w, pos, pstate, u1, hw);
        if (LOCK) {
            // a half that cannot continue ends the lockstep phase before anything of this event is written;
            // the class choice is redone (identically) by the lone loops
            if (__any_sync(FULL, pick.zero)) { pstate.need_sort = true; return; }
        }
        if (pick.zero) {                                        // kmc.py:31-32: nothing can happen
            flag |= FLAG_ZERO_RATE;
            if (p.log_mode != KMC_LOG_NONE && hl == 0) {
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
            if (hl < NC) rec->counts[hl] = (uint32_t)tot;
        }

        // ---- 3. in-class pick: rank in (row, slot, column) order -----------------------------------
        const int cnt = HSHFL(tot, csel);
        uint32_t r;
        {
            long long rr = (long long)__dmul_rn(u2, (double)cnt);
            if (rr > (long long)cnt - 1) rr = (long long)cnt - 1;
            r = (uint32_t)rr;
        }
        // 3a. row: four rows per half-lane per pass (one 64-bit load of four uint16 counts)
        int row = 0;
        {
            bool found = false;
            for (int base = 0; base < d0; base += 64) {
                const int r0 = base + 4 * hl;
                const u64 v4 = (r0 < d0) ? rc64[(csel * d0p + r0) >> 2] : 0ull;
                const uint32_t lo2 = (uint32_t)v4, hi2 = (uint32_t)(v4 >> 32);
                const uint32_t s = (lo2 & 0xFFFFu) + (lo2 >> 16) + (hi2 & 0xFFFFu) + (hi2 >> 16);
                uint32_t inc = s;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t tt = HSHFL_UP(inc, o);
                    if (hl >= o) inc += tt;
                }
                const uint32_t total = HSHFL(inc, 15);
                const int L = (__ffs(HS::ballot(hw, inc > r)) - 1) & 15;
                const uint32_t excl = HSHFL(inc - s, L);
                const uint32_t wl = HSHFL(lo2, L), wh = HSHFL(hi2, L);
                if (!found) {
                    if (r >= total) r -= total;
                    else {
                        found = true;
                        r -= excl;
                        const uint32_t f0 = wl & 0xFFFFu, f1 = wl >> 16, f2 = wh & 0xFFFFu;
                        int j = 0;
                        if (r >= f0) { r -= f0; j = 1; if (r >= f1) { r -= f1; j = 2; if (r >= f2) { r -= f2; j = 3; } } }
                        row = base + 4 * L + j;
                    }
                }
                if (TD0 && TD0 <= 64) break;
            }
        }
        // 3b/3c. slot, then site, within the row
        int col = 0, slot = 0, s0 = 0;
        const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 : 0);
        if (HS::any(hw, group == 1)) {
            u64 pr, ne, fa;
            hw_dir_masks<TD0>(PL, d0, row, my_dir, rb, pr, ne, fa);
            const u64 mine = hl < 6 ? kind_select(pr, ne, fa, (csel - C_MOVE_DIV) & 3) : 0ull;
            const uint32_t cntk = __popcll(mine);
            uint32_t pre = cntk, t1;
            t1 = HSHFL_UP(pre, 1); if (hl >= 1) pre += t1;
            t1 = HSHFL_UP(pre, 2); if (hl >= 2) pre += t1;
            t1 = HSHFL_UP(pre, 4); if (hl >= 4) pre += t1;
            const int ksel = (__ffs(HS::ballot(hw, pre > r) & 0x3Fu) - 1) & 7;
            const uint32_t rem = r - HSHFL(pre - cntk, ksel);
            const int c1 = nth_set_bit_hw<HS>(HSHFL(mine, ksel), rem, hw, my_low_lo, my_low_hi);
            if (group == 1) { col = c1; slot = 7 + ksel; s0 = S_DIVAC; }
        }
        if (HS::any(hw, group == 2)) {
            u64 up, dn;
            hw_trefoil_masks<TD0>(PL, d0, row, rb, up, dn);
            const uint32_t nu = __popcll(up);
            const bool second = r >= nu;
            const int c1 = nth_set_bit_hw<HS>(second ? dn : up, second ? r - nu : r, hw, my_low_lo, my_low_hi);
            if (group == 2) { col = c1; slot = 13 + (int)second; s0 = S_DIVAC; }
        }
        if (HS::any(hw, group == 0)) {
            int pa, pb = -1;
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
            u64 ma = PL[pa * d0 + row];
            const u64 mb = pb >= 0 ? PL[pb * d0 + row] : 0ull;
            if (csel == C_FLIP) ma |= PL[T_M2 * d0 + row];
            const uint32_t n1 = __popcll(ma);
            const bool second = r >= n1;
            const int c1 = nth_set_bit_hw<HS>(second ? mb : ma, second ? r - n1 : r, hw, my_low_lo, my_low_hi);
            if (group == 0) {
                col = c1;
                slot = class_slot_base(csel) + (int)second;
                switch (csel) {
                    case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0 = 0; break;
                    case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0 = 3; break;
                    case C_DESTROY_TREF: s0 = second ? 7 : 4; break;
                    case C_CMONO_MAKE_EMPTY: s0 = second ? 1 : 2; break;
                    case C_FMONO_MAKE_DOUBLE: s0 = second ? 2 : 1; break;
                    default: s0 = ((PL[T_M1 * d0 + row] >> col) & 1ull) ? 1 : 2; break;
                }
            }
        }
        const int site = row * d1 + col;

        // ---- event record -----------------------------------------------------------------------------
        if (p.log_mode != KMC_LOG_NONE && hl == 0) {
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
        if (hl == csel) hist++;
        if (pick.tie && hl == 0) n_ties++;

        // ---- 4. the move (uniform inside the half): touched nodes, old and new status ------------------
        MoveNodes mv;
        int os1, os2;
        {
            const MoveEntry e = MOVE_TABLE[slot];
            mv.ns0 = (int)((((uint32_t)s0 | (e.w0 & 0xF)) & ((e.w0 >> 4) & 0xF)) ^ ((e.w0 >> 8) & 0xF));
            mv.na = (int)((e.w0 >> 12) & 3);
            mv.da1 = (int)(e.w1 & 0xF) - 2;
            mv.a1 = wrapT<TD0>(row + mv.da1, d0);
            mv.b1 = wrapT<TD1>(col + (int)((e.w1 >> 4) & 0xF) - 2, d1);
            mv.a2 = wrapT<TD0>(row + (int)((e.w1 >> 8) & 0xF) - 2, d0);
            mv.b2 = wrapT<TD1>(col + (int)((e.w1 >> 12) & 0xF) - 2, d1);
            mv.ns1 = (int)((e.w1 >> 16) & 0xF); mv.ns2 = (int)((e.w1 >> 20) & 0xF);
            os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
        }
        const int na = mv.na;
        if (hl < HW_PLANES) {
            uint32_t *P32 = reinterpret_cast<uint32_t *>(PL + hl * d0);
            {
                uint32_t *wp = P32 + 2 * row + (col >> 5);
                const uint32_t bit = 1u << (col & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns0) == hl ? bit : 0u);
            }
            if (na > 1) {
                uint32_t *wp = P32 + 2 * mv.a1 + (mv.b1 >> 5);
                const uint32_t bit = 1u << (mv.b1 & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns1) == hl ? bit : 0u);
            }
            if (na > 2) {
                uint32_t *wp = P32 + 2 * mv.a2 + (mv.b2 >> 5);
                const uint32_t bit = 1u << (mv.b2 & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns2) == hl ? bit : 0u);
            }
        }
        // ---- 5a. own-status classes: owner lanes fold (old -> new) of each touched node -----------------
        if (my_g1) {
            int d = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
            if (d) rc16[hl * d0p + row] = (uint16_t)(rc16[hl * d0p + row] + d);
            int dsum = d;
            if (na > 1) {
                d = (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
                if (d) rc16[hl * d0p + mv.a1] = (uint16_t)(rc16[hl * d0p + mv.a1] + d);
                dsum += d;
            }
            if (na > 2) {
                d = (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
                if (d) rc16[hl * d0p + mv.a2] = (uint16_t)(rc16[hl * d0p + mv.a2] + d);
                dsum += d;
            }
            tot += dsum;
        }
        HS::sync(hw);

        // ---- 5b. MoveDivacancy: band rows row+lo .. row+hi re-derived and replaced, two rows per pass ----
        int lo = -1, hi = 1;
        if (slot >= 13) hi = 3;
        else if (slot >= 7) { lo += mv.da1 < 0 ? mv.da1 : 0; hi += mv.da1 > 0 ? mv.da1 : 0; }
        {
            const int nrows = hi - lo + 1;
            const int nloop = HS::umax(hw, nrows);              // lockstep: the longer band of the two halves
            uint32_t d01 = 0, d23 = 0;
            for (int j0 = 0; j0 < nloop; j0 += 2) {
                const int j = j0 + my_g;
                const bool act = hl < 12 && j < nrows;
                const int ra = wrapT<TD0>(row + lo + (j < nrows ? j : 0), d0);
                uint32_t pk01 = 0, pk23 = 0;
                if (act) {
                    u64 pr, ne, fa;
                    hw_dir_masks<TD0>(PL, d0, ra, my_dir, rb, pr, ne, fa);
                    kind_counts(pr, ne, fa, pk01, pk23);
                }
                pk01 += HSHFL_DOWN(pk01, 3);
                pk23 += HSHFL_DOWN(pk23, 3);
                pk01 += HSHFL_DOWN(pk01, 1) + HSHFL_DOWN(pk01, 2);
                pk23 += HSHFL_DOWN(pk23, 1) + HSHFL_DOWN(pk23, 2);
                if (act && my_k == 0) {
                    uint16_t *q = rc16 + C_MOVE_DIV * d0p + ra;
                    const uint32_t o0 = q[0], o1 = q[d0p], o2 = q[2 * d0p], o3 = q[3 * d0p];
                    q[0] = (uint16_t)pk01; q[d0p] = (uint16_t)(pk01 >> 16);
                    q[2 * d0p] = (uint16_t)pk23; q[3 * d0p] = (uint16_t)(pk23 >> 16);
                    d01 += pk01 - (o0 | (o1 << 16));
                    d23 += pk23 - (o2 | (o3 << 16));
                }
            }
            // the two group leaders (half-lanes 0 and 6) hold the packed signed deltas
            const int s01 = (int)(HSHFL(d01, 0) + HSHFL(d01, 6));
            const int s23 = (int)(HSHFL(d23, 0) + HSHFL(d23, 6));
            const int l01 = (int)(short)(s01 & 0xFFFF), l23 = (int)(short)(s23 & 0xFFFF);
            if (hl == C_MOVE_DIV) tot += l01;
            if (hl == C_MOVE_DIV + 1) tot += (s01 - l01) >> 16;
            if (hl == C_MOVE_DIV + 2) tot += l23;
            if (hl == C_MOVE_DIV + 3) tot += (s23 - l23) >> 16;
        }
        // ---- 5c. CreateTrefoil: anchor rows row-3 .. row+2 -------------------------------------------------
        {
            const int ng = d0 < 6 ? d0 : 6;
            int d = 0;
            if (hl < ng) {
                const int ra = wrapT<TD0>(row + hl - 3, d0);
                u64 up, dn;
                hw_trefoil_masks<TD0>(PL, d0, ra, rb, up, dn);
                const int nw = __popcll(up) + __popcll(dn);
                uint16_t *q = rc16 + C_CREATE_TREF * d0p + ra;
                d = nw - (int)q[0];
                q[0] = (uint16_t)nw;
            }
            int s = d + HSHFL_DOWN(d, 3);                      // half-lanes 0..5 hold the row deltas
            s += HSHFL_DOWN(s, 1) + HSHFL_DOWN(s, 2);
            s = HSHFL(s, 0);
            if (hl == C_CREATE_TREF) tot += s;
        }
        HS::sync(hw);
    }
    };   // run
    if (pair_ok) run(std::true_type());
    run(std::false_type());                 // the rest: nothing, or a half whose partner stopped / never existed


    // ---- epilogue ------------------------------------------------------------------------------------
    const long long done = t;
    if (p.validate) {
        bool bad = false;
        for (int a = hl; a < d0; a += 16) {
            uint32_t cnt[NC];
            hw_row_counts<TD0>(PL, d0, a, rb, cnt);
#pragma unroll
            for (int c = 0; c < NC; c++) bad |= (rc16[c * d0p + a] != (uint16_t)cnt[c]);
            const u64 z = PL[T_Z * d0 + a], dv = PL[T_D * d0 + a];
            const u64 types[N_TYPES] = {z, dv, PL[T_M1 * d0 + a], PL[T_M2 * d0 + a], PL[T_A4 * d0 + a], PL[T_A7 * d0 + a]};
            u64 seen = 0, dup = 0;
#pragma unroll
            for (int q = 0; q < N_TYPES; q++) { dup |= seen & types[q]; seen |= types[q]; }
            bad |= dup != 0;
        }
        if (hl < NC) {
            int fresh = 0;
            for (int a = 0; a < d0; a++) fresh += rc16[hl * d0p + a];
            bad |= (fresh != tot);
        }
        if (__any_sync(hw.mask, bad)) flag |= FLAG_VALIDATE;
    }
    for (int a = hl; a < d0; a += 16) {
        const u64 z = PL[T_Z * d0 + a], dv = PL[T_D * d0 + a], m1 = PL[T_M1 * d0 + a], m2 = PL[T_M2 * d0 + a];
        const u64 a4 = PL[T_A4 * d0 + a], a7 = PL[T_A7 * d0 + a];
        auto code_at = [&](int b) -> uint32_t {
            return ((z >> b) & 1) ? 0u : ((dv >> b) & 1) ? 3u : ((m1 >> b) & 1) ? 1u : ((m2 >> b) & 1) ? 2u
                 : ((a4 >> b) & 1) ? 4u : ((a7 >> b) & 1) ? 7u : 5u;
        };
        if ((d1 & 3) == 0) {
            uint32_t *g32 = reinterpret_cast<uint32_t *>(gst + a * d1);
            for (int b4 = 0; b4 < (d1 >> 2); b4++) {
                const int b = 4 * b4;
                g32[b4] = code_at(b) | (code_at(b + 1) << 8) | (code_at(b + 2) << 16) | (code_at(b + 3) << 24);
            }
        } else {
            for (int b = 0; b < d1; b++) gst[a * d1 + b] = (uint8_t)code_at(b);
        }
    }
    __syncwarp(hw.mask);
    for (int a = hl; a < d0; a += 16) {
        u64 a4 = PL[T_A4 * d0 + a], a7 = PL[T_A7 * d0 + a];
        const int a2 = wrap1(a + 2, d0);
        while (a4) {
            const int b = __ffsll((long long)a4) - 1; a4 &= a4 - 1;
            gst[a2 * d1 + b] = (uint8_t)(S_TREF + 1);
            gst[a * d1 + wrap1(b + 2, d1)] = (uint8_t)(S_TREF + 2);
        }
        while (a7) {
            const int b = __ffsll((long long)a7) - 1; a7 &= a7 - 1;
            gst[a2 * d1 + b] = (uint8_t)(S_TREF + 4);
            gst[a2 * d1 + wrap1(b - 2, d1)] = (uint8_t)(S_TREF + 5);
        }
    }
    if (hl < NC) p.hist[(size_t)rep * NC + hl] += hist;
    if (hl == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (p.clock) p.clock[rep] = __dadd_rn(p.clock[rep], tclock);
        p.ties[rep] += n_ties;
        if (flag) p.flags[rep] |= flag;
    }
}

#undef HSHFL
#undef HSHFL_UP
#undef HSHFL_DOWN

}  // namespace kmc
