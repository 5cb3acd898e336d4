// replica_hw_kernel.cuh -- K3 + K4 on the bit-sliced lattice with TWO replicas per warp (the batch kernel).
//
// Same per-event algorithm as replica_bp_kernel.cuh (bit-sliced lattice, class-major row-count table,
// recompute-and-replace refresh); what changes is the mapping and the footprint:
//   * a replica is owned by a HALF-warp (16 lanes).  Almost all of an event is work that needs at most 16 lanes
//     (class choice over 13+3 weight lanes, rank bookkeeping, move decoding, six plane lanes, 2 x 6 direction
//     lanes), so a full warp per replica spends most issue slots on instructions whose upper half is idle;
//   * the two halves of a warp run in LOCKSTEP (see HalfSync below): data-dependent branches that contain a
//     shuffle or vote are taken if either half needs them, so every intrinsic names the full warp with width 16;
//     a half whose partner stopped (total rate zero) or does not exist finishes with half-warp masks
//     (replica_hw_body.inc is included once for each mode);
//   * shared memory and registers bound residency together, so the kernel keeps only the six type planes
//     (neighbour rows are rotated on the fly, two funnel shifts per rotate) and row counts only for the five
//     neighbourhood classes (own-status row counts are plane popcounts): 3.6 KiB per replica at 64 rows,
//     one block of 28 warps = 56 replicas per SM at 72 registers.
//
// Lane roles inside a half (hl = lane & 15):
//   hl < 13   owns class hl (total count, rate)                             hl < 6   maintains bit plane hl
//   hl < 6    direction hl of the MoveDivacancy masks during the site search, CreateTrefoil anchor rows
//   hl < 12   refresh: direction hl % 6 of band rows 2*pass + hl / 6
#pragma once
#include "replica_bp_kernel.cuh"

namespace kmc {

// plane index == plane type (no pre-rotated copies here)
enum { HW_PLANES = N_TYPES };      // plane index == plane type: T_Z, T_D, T_M1, T_M2, T_A4, T_A7

// Row counts are kept only for the five neighbourhood classes (MoveDivacancy x 4, CreateTrefoil); the row
// counts of the eight own-status classes are popcounts of the planes (derived when such a class is chosen).
// With MoveMonovacancy (MM build) three more tables follow: natural (5), merge-divacancy (6), swap-divacancy (7);
// split-divacancy needs none: a divacancy with a pristine neighbour in direction k offers that hop in both layers,
// so its row count is twice the sum of the row's four MoveDivacancy counts.
enum { HW_RC = 5, HW_RC_MM = 8 };
__host__ __device__ inline size_t replica_hw_bytes(int d0, bool mm = false) {      // per replica
    const size_t d0p = ((size_t)d0 + 3) & ~(size_t)3;
    const size_t planes = (size_t)HW_PLANES * d0 * 8;
    const size_t rc = ((mm ? HW_RC_MM : HW_RC) * d0p * 2 + 15) & ~(size_t)15;
    return planes + rc;
}
// row table of a class, or -1 when its row counts are derived (own-status classes: plane popcounts; split: see above)
template <bool MM> __device__ __forceinline__ int hw_table_of(int c) {
    if (c >= C_MOVE_DIV && c <= C_CREATE_TREF) return c - C_MOVE_DIV;
    if (MM && (c == C_MOVE_MONO || c == C_MOVE_MONO + 1)) return c - 8;          // 13 -> 5, 14 -> 6
    if (MM && c == C_MOVE_MONO + 3) return 7;
    return -1;
}
// total number of moves of an own-status class over the lattice: sum over plane types of
// g1(status of the type) * popcount (g1: 2 bits per status, see g1_const)
__device__ __forceinline__ int hw_own_total(const u64 *PL, int d0, uint32_t g1) {
    int tot = 0;
    for (int a = 0; a < d0; a++)
        tot += (int)(g1 & 3u) * __popcll(PL[T_Z * d0 + a]) + (int)((g1 >> 6) & 3u) * __popcll(PL[T_D * d0 + a]) +
               (int)((g1 >> 2) & 3u) * __popcll(PL[T_M1 * d0 + a]) + (int)((g1 >> 4) & 3u) * __popcll(PL[T_M2 * d0 + a]) +
               (int)((g1 >> 8) & 3u) * __popcll(PL[T_A4 * d0 + a]) + (int)((g1 >> 14) & 3u) * __popcll(PL[T_A7 * d0 + a]);
    return tot;
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
// MoveMonovacancy of row `a`, direction constants `c` (only the neighbour part is used): per-kind masks of the
// sites whose vacancy can hop that way.  natural / merge are per layer (the layer-1 and layer-2 masks are
// disjoint, so their union has the summed popcount); split is hw_dir_masks' `present`; swap as below.
template <int TD0, int TD1>
__device__ __forceinline__ void hw_mm_masks(const u64 *PL, int d0, int a, const HwDir &c, const RowBits<TD1> &rb,
                                            u64 &nat1, u64 &nat2, u64 &mer1, u64 &mer2, u64 &swp1, u64 &swp2) {
    const int an = wrapT<TD0>(a + c.da_n, d0);
    const u64 zn = rot_site(PL[T_Z * d0 + an], c.db_n, c.sh_n, c.db_n > 0, rb);
    const u64 m1n = rot_site(PL[T_M1 * d0 + an], c.db_n, c.sh_n, c.db_n > 0, rb);
    const u64 m2n = rot_site(PL[T_M2 * d0 + an], c.db_n, c.sh_n, c.db_n > 0, rb);
    const u64 m1 = PL[T_M1 * d0 + a], m2 = PL[T_M2 * d0 + a], dv = PL[T_D * d0 + a];
    nat1 = m1 & zn; nat2 = m2 & zn;          // layer L vacancy -> pristine neighbour
    mer1 = m1 & m2n; mer2 = m2 & m1n;        // -> neighbour holding the opposite-layer monovacancy
    swp1 = dv & m2n; swp2 = dv & m1n;        // layer L of a divacancy -> opposite-layer monovacancy
}
// counts of (natural | merge << 16, swap) of one (row, direction)
template <int TD0, int TD1>
__device__ __forceinline__ void hw_mm_counts(const u64 *PL, int d0, int a, const HwDir &c, const RowBits<TD1> &rb,
                                             uint32_t &nm, uint32_t &sw) {
    u64 n1, n2, e1, e2, s1, s2;
    hw_mm_masks<TD0>(PL, d0, a, c, rb, n1, n2, e1, e2, s1, s2);
    nm = (uint32_t)__popcll(n1 | n2) | ((uint32_t)__popcll(e1 | e2) << 16);
    sw = (uint32_t)__popcll(s1) + (uint32_t)__popcll(s2);
}
template <int TD0, int TD1, bool MM = false>
__device__ __forceinline__ void hw_row_counts(const u64 *PL, int d0, int a, const RowBits<TD1> &rb, uint32_t out[NCA]) {
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
    if (MM) {
        uint32_t nm = 0, sw = 0;
        for (int k = 0; k < 6; k++) {
            const HwDir c = hw_dir_const(k);
            uint32_t a_nm, a_sw; hw_mm_counts<TD0>(PL, d0, a, c, rb, a_nm, a_sw);
            nm += a_nm; sw += a_sw;
        }
        out[C_MOVE_MONO] = nm & 0xFFFF; out[C_MOVE_MONO + 1] = nm >> 16;
        out[C_MOVE_MONO + 2] = 2 * ((k01 & 0xFFFF) + (k01 >> 16) + (k23 & 0xFFFF) + (k23 >> 16));
        out[C_MOVE_MONO + 3] = sw;
    }
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
    // first sorted position whose weight differs from the previous event's (st.wprev holds the weight this
    // lane's sorted position had one event ago; meaningless after a re-sort, which starts from 0 anyway)
    const uint32_t changed = HS::ballot(hw, ws != st.wprev);
    int first = changed ? __ffs((int)changed) - 1 : 16;
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
    st.wprev = ws;
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

// One block of up to 28 warps (56 replicas) per SM: 72 registers per thread let 28 warps be resident, 7 per
// scheduler (a block shape that is not a multiple of 4 warps loads the schedulers unevenly: 3 blocks of 9
// warps measured 1.89e9 events/s, 2 blocks of 14 warps 2.22e9, 1 block of 28 warps 2.37e9).
constexpr int HW_MAX_THREADS = 896;
// PLAIN = true: the statistics-only build (no event log, no injected draws, no clock) -- what a production
// batch runs; the general build keeps every option.
// MM = true: MoveMonovacancy's four classes are maintained as well.  Half-lanes 13, 14, 15 own classes 13, 14, 15; class
// 16 (swap-divacancy) lives on half-lane p.lane16, the lane of a disabled own-status class (17 classes, 16 lanes: a
// rule set with all 17 enabled runs on the byte-lattice kernel instead).
template <int TD0, int TD1, bool PLAIN = false, bool MM = false>
__global__ void __launch_bounds__(HW_MAX_THREADS, 1) kmc_replica_hw_kernel(const ReplicaParams p)
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
    u64 *PL = reinterpret_cast<u64 *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)slot_in_block * replica_hw_bytes(d0, MM));
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(PL + HW_PLANES * d0);      // [NRC][d0p]: classes 2..6 (+ 13, 14, 16)
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
        uint32_t cnt[NCA];
        if (a < d0) hw_row_counts<TD0, TD1, MM>(PL, d0, a, rb, cnt);
#pragma unroll
        for (int c = 0; c < HW_RC; c++) rc16[c * d0p + a] = a < d0 ? (uint16_t)cnt[C_MOVE_DIV + c] : (uint16_t)0;
        if (MM) {
            rc16[5 * d0p + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO] : (uint16_t)0;
            rc16[6 * d0p + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO + 1] : (uint16_t)0;
            rc16[7 * d0p + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO + 3] : (uint16_t)0;
        }
    }
    __syncwarp(hw.mask);

    // per-lane class state: half-lane c < 13 owns class c; 13..15 are permanent zero weights -- or, in the MM build,
    // own classes 13..15, with class 16 on half-lane p.lane16
    const int lane16 = MM ? p.lane16 : -1;
    const int nown = MM ? 16 : NC;
    const int myc = (MM && hl == lane16) ? C_MOVE_MONO + 3 : (hl < nown ? hl : 0);      // the class this lane owns
    const bool has_class = hl < nown;
    const double rate = (has_class && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (has_class && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + hl;
    const uint32_t my_g1 = (myc < NC && has_class && !(myc >= C_MOVE_DIV && myc <= C_CREATE_TREF)) ? g1_const(myc) : 0u;
    const int my_tab = has_class ? hw_table_of<MM>(myc) : -1;
    int tot = 0;
    if (my_tab >= 0)
        for (int a = 0; a < d0; a++) tot += rc16[my_tab * d0p + a];
    else if (MM && myc == C_MOVE_MONO + 2) {
        for (int a = 0; a < d0; a++) tot += 2 * (rc16[a] + rc16[d0p + a] + rc16[2 * d0p + a] + rc16[3 * d0p + a]);
    } else if (my_g1) tot = hw_own_total(PL, d0, my_g1);
    uint32_t hist = 0, hops = 0, n_ties = 0;                 // per launch (a launch holds at most 2^30 events)
    PickState pstate = pick_state_init(hl);
    double *const clockp = PLAIN ? nullptr : p.clock;
    double tclock = clockp ? clockp[rep] : 0.0;            // continues across launches in event order
    uint32_t *const tracer = (PLAIN || !p.tracer) ? nullptr : p.tracer + (size_t)rep * n;
    // refresh roles: half-lanes 0..11 = direction hl % 6 of band rows 2*pass + hl / 6
    const int my_g = hl / 6, my_k = hl % 6;
    HwDir my_dir = hw_dir_const(my_k);
    asm volatile("" : "+r"(my_dir.da_n), "+r"(my_dir.db_n), "+r"(my_dir.sh_n), "+r"(my_dir.da_e), "+r"(my_dir.db_e),
                      "+r"(my_dir.sh_e), "+r"(my_dir.da_f), "+r"(my_dir.db_f), "+r"(my_dir.sh_f));
    // MM site search: half-lane l < 12 holds slot 17 + l = (direction l >> 1, layer (l & 1) + 1)
    HwDir mm_dir = hw_dir_const(hl < 12 ? hl >> 1 : 0);
    if (MM) asm volatile("" : "+r"(mm_dir.da_n), "+r"(mm_dir.db_n), "+r"(mm_dir.sh_n));
    // this lane's mask of the sites below column 4*hl + 4 (for popcount-select)
    uint32_t my_low_lo = hl >= 7 ? 0xFFFFFFFFu : ((16u << (4 * hl)) - 1u);
    uint32_t my_low_hi = hl < 8 ? 0u : (hl == 15 ? 0xFFFFFFFFu : ((16u << (4 * (hl - 8))) - 1u));
    asm volatile("" : "+r"(my_low_lo), "+r"(my_low_hi));

    const unsigned long long step_base = p.steps_done[rep];
    const int log_mode = PLAIN ? KMC_LOG_NONE : p.log_mode;
    const double *const draws = PLAIN ? nullptr : p.draws;
    double mu1 = 0.0, mu2 = 0.0;
    const int nsteps = (int)p.nsteps;
    int t = 0;
    uint32_t flag = 0;

    // the event loop: in lockstep while both halves can continue, then whatever is left on its own
    // (normally nothing; a half whose partner hit a zero total rate, or never existed)
    if (pair_ok) {
#define KMC_HW_LOCK true
#include "replica_hw_body.inc"
#undef KMC_HW_LOCK
    }
#define KMC_HW_LOCK false
#include "replica_hw_body.inc"
#undef KMC_HW_LOCK


    // ---- epilogue ------------------------------------------------------------------------------------
    const long long done = t;
    if (p.validate) {
        bool bad = false;
        for (int a = hl; a < d0; a += 16) {
            uint32_t cnt[NCA];
            hw_row_counts<TD0, TD1, MM>(PL, d0, a, rb, cnt);
#pragma unroll
            for (int c = 0; c < HW_RC; c++) bad |= (rc16[c * d0p + a] != (uint16_t)cnt[C_MOVE_DIV + c]);
            if (MM) {
                bad |= rc16[5 * d0p + a] != (uint16_t)cnt[C_MOVE_MONO];
                bad |= rc16[6 * d0p + a] != (uint16_t)cnt[C_MOVE_MONO + 1];
                bad |= rc16[7 * d0p + a] != (uint16_t)cnt[C_MOVE_MONO + 3];
            }
            const u64 z = PL[T_Z * d0 + a], dv = PL[T_D * d0 + a];
            const u64 types[N_TYPES] = {z, dv, PL[T_M1 * d0 + a], PL[T_M2 * d0 + a], PL[T_A4 * d0 + a], PL[T_A7 * d0 + a]};
            u64 seen = 0, dup = 0;
#pragma unroll
            for (int q = 0; q < N_TYPES; q++) { dup |= seen & types[q]; seen |= types[q]; }
            bad |= dup != 0;
        }
        if (my_tab >= 0) {
            int fresh = 0;
            for (int a = 0; a < d0; a++) fresh += rc16[my_tab * d0p + a];
            bad |= (fresh != tot);
        } else if (MM && myc == C_MOVE_MONO + 2) {
            int fresh = 0;
            for (int a = 0; a < d0; a++) fresh += 2 * (rc16[a] + rc16[d0p + a] + rc16[2 * d0p + a] + rc16[3 * d0p + a]);
            bad |= (fresh != tot);
        } else if (my_g1) {
            bad |= (hw_own_total(PL, d0, my_g1) != tot);
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
    if (has_class) p.hist[(size_t)rep * NCA + myc] += hist;
    if (hl < 6) p.hops[(size_t)rep * 6 + hl] += hops;
    if (hl == 0) {
        p.steps_done[rep] = step_base + (unsigned long long)done;
        if (clockp) clockp[rep] = tclock;
        p.ties[rep] += n_ties;
        if (flag) { p.flags[rep] |= flag; atomicOr(p.any_flag, flag); }
    }
}

#undef HSHFL
#undef HSHFL_UP
#undef HSHFL_DOWN

}  // namespace kmc
