// replica_bp_kernel.cuh -- K3 + K4 for lattices with rows of <= 64 sites: the fused
// select -> apply -> refresh kernel on a BIT-SLICED lattice.
//
// One warp owns one replica.  The lattice is held in shared memory as six predicate bit-planes, one
// 64-bit word per lattice row per plane:
//     Z  pristine          D  divacancy         M1 / M2  monovacancy in layer 1 / 2
//     A4 / A7  anchor (role 0) of an Up / Down trefoil          (no bit set = other trefoil member)
// so that every neighbourhood predicate of the reference's rules (demo1/rules.py) becomes a handful
// of 64-site-wide Boolean operations on rotated row words (periodic wrap = bit rotation):
//     MoveDivacancy direction k, kind q :  D[a] & rot(Z[a+da_n]) & (+/-)rot(D[near]) & (+/-)rot(D[far])
//     CreateTrefoil Up / Down           :  D[a] & D[a+2] & rot(D[a], +2)   /   D[a] & D[a+2] & rot(D[a+2], -2)
// The Fenwick-style hierarchy for the in-class rank search is the per-row x per-class move count table
// (uint16, class-major) next to the planes; class totals live in registers (lane c owns class c).
//
// Neighbourhood refresh after an event (reference sim.py:149-157, rules.py pre/post_status_change) is
// "recompute and replace": the <= 5 rows whose MoveDivacancy counts can change are re-derived by
// 30 lanes (row x direction) with popcounts, the <= 6 anchor rows of CreateTrefoil by 6 lanes, and the
// own-status classes by their owner lanes from (old status, new status) of the <= 3 touched nodes.
// No per-site re-enumeration, no atomics.
//
// Results are bit-identical to replica_kernel.cuh (the byte-lattice kernel used for wider rows) and to
// the CPU oracle; the tests run both.
#pragma once
#include "replica_kernel.cuh"

namespace kmc {

typedef unsigned long long u64;
// Plane types (what a status byte maps to) and the planes held in shared memory.  Z and D are kept
// three times: as is, and pre-rotated by one site either way (ZP: bit b = Z bit b+1, ZM: bit b = Z bit
// b-1), so that "the neighbour at column offset +-1 is pristine / a divacancy" is a plain load of the
// right plane instead of a per-lane variable 64-bit rotate.
enum { T_Z = 0, T_D = 1, T_M1 = 2, T_M2 = 3, T_A4 = 4, T_A7 = 5, N_TYPES = 6 };
enum { PL_Z = 0, PL_ZP = 1, PL_ZM = 2, PL_D = 3, PL_DP = 4, PL_DM = 5, PL_M1 = 6, PL_M2 = 7, PL_A4 = 8, PL_A7 = 9,
       N_PLANES = 10 };

// plane stride in 64-bit words: rows of different planes at the same row index would otherwise share a bank
// (a 64-row plane is 512 B), which makes the plane update (one lane per plane) and the direction masks
// (different planes, neighbouring rows) serialise; KMC_BP_PAD extra words spread the planes over the banks
#ifndef KMC_BP_PAD
#define KMC_BP_PAD 3
#endif
#define BPS(d0) ((d0) + KMC_BP_PAD)
// Row tables: one per class of the reference's own rules (13); the MoveMonovacancy build keeps three more -- natural
// (13), merge-divacancy (14), swap-divacancy (table 15 = class 16); split-divacancy (class 15) is derived: twice the sum
// of the row's four MoveDivacancy counts.
constexpr int BP_NCT_MM = 16;
__host__ __device__ inline size_t replica_bp_warp_bytes(int d0, bool mm = false) {
    const size_t d0e = ((size_t)d0 + 1) & ~(size_t)1;
    const size_t planes = (size_t)N_PLANES * BPS(d0) * 8;
    const size_t rc = ((mm ? BP_NCT_MM : NC) * d0e * 2 + 15) & ~(size_t)15;
    return planes + rc;
}
__device__ __forceinline__ int bp_table_of(int c) { return c < C_MOVE_MONO + 2 ? c : (c == C_MOVE_MONO + 3 ? 15 : -1); }

// neighbors_and_mutuals (reference state.py:443-459), direction k: (da, db) of nbr / near / far
constexpr uint32_t NBR_DA = (uint32_t)pack_nibbles({-1, 0, 1, 0, -1, 1}),  NBR_DB = (uint32_t)pack_nibbles({1, -1, 0, 1, 0, -1});
constexpr uint32_t NEAR_DA = (uint32_t)pack_nibbles({0, -1, 1, -1, 0, 1}), NEAR_DB = (uint32_t)pack_nibbles({1, 0, -1, 1, -1, 0});
constexpr uint32_t FAR_DA = (uint32_t)pack_nibbles({-1, 1, 0, 1, -1, 0}),  FAR_DB = (uint32_t)pack_nibbles({0, -1, 1, 0, 1, -1});
__host__ __device__ __forceinline__ int nib(uint32_t t, int k) { return (int)((t >> (4 * k)) & 0xF) - 2; }

// status -> plane type + 1 (0 = none: non-anchor trefoil member)
//   s:     0     1      2      3     4      5  6  7      8  9
//   type:  Z(1)  M1(3)  M2(4)  D(2)  A4(5)  0  0  A7(6)  0  0
// (3 bits per status so the table fits one 32-bit register and the lookup is a shift + mask)
constexpr uint32_t STATUS_TYPE = 1u | (3u << 3) | (4u << 6) | (2u << 9) | (5u << 12) | (6u << 21);
__host__ __device__ __forceinline__ int status_plane(int s) { return (int)((STATUS_TYPE >> (3 * s)) & 7u) - 1; }
// plane q (0..9) -> its type, and the column offset at which a site's bit lives in it
//   q:      Z  ZP  ZM  D  DP  DM  M1 M2 A4 A7
//   type:   0  0   0   1  1   1   2  3  4  5
//   cshift: 0  -1  +1  0  -1  +1  0  0  0  0     (stored + 1)
constexpr u64 PLANE_TYPE = 0x5432111000ull, PLANE_CSHIFT = 0x1111201201ull;   // read once per lane

template <int TD1>
struct RowBits {
    int d1; u64 mask;
    __device__ __forceinline__ RowBits(int d1_) : d1(TD1 ? TD1 : d1_), mask(d1 >= 64 ? ~0ull : ((1ull << d1) - 1)) {}
    // word whose bit b is x's bit (b + db), periodic in d1;  s = (-db) mod d1 is the left-rotate amount
    __device__ __forceinline__ u64 rotl(u64 x, int s) const {
        if (TD1 == 64) return (x << s) | (x >> ((64 - s) & 63));
        return s ? (((x << s) | (x >> (d1 - s))) & mask) : x;
    }
    __device__ __forceinline__ int amount(int db) const { return db > 0 ? d1 - db : -db; }
    // bits [0, n) set, n in 0..64
    __device__ __forceinline__ static u64 low(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1); }
};

__device__ __forceinline__ int wrap1(int x, int d) { return x < 0 ? x + d : (x >= d ? x - d : x); }
// periodic index for |offset| <= 3; a compile-time power-of-two extent wraps with one AND
template <int TD> __device__ __forceinline__ int wrapT(int x, int d) {
    if (TD > 0 && (TD & (TD - 1)) == 0) return x & (TD - 1);
    return wrap1(x, d);
}

// per-direction constants: element offsets (plane * d0) of the pre-rotated plane to read for the
// neighbour / near / far site, and their row offsets
struct DirConst { int off_n, da_n, off_e, da_e, off_f, da_f; };
__device__ __forceinline__ DirConst dir_const(int k, int d0) {
    DirConst c;
    // column offset db -> plane: db = 0 -> base, +1 -> P (bit b = bit b+1), -1 -> M
    const int dbn = nib(NBR_DB, k), dbe = nib(NEAR_DB, k), dbf = nib(FAR_DB, k);
    c.off_n = (PL_Z + (dbn > 0 ? 1 : dbn < 0 ? 2 : 0)) * BPS(d0);  c.da_n = nib(NBR_DA, k);
    c.off_e = (PL_D + (dbe > 0 ? 1 : dbe < 0 ? 2 : 0)) * BPS(d0);  c.da_e = nib(NEAR_DA, k);
    c.off_f = (PL_D + (dbf > 0 ? 1 : dbf < 0 ? 2 : 0)) * BPS(d0);  c.da_f = nib(FAR_DA, k);
    return c;
}

// MoveDivacancy of row `a`, direction constants `c`: present / near / far masks over the row's sites
template <int TD0>
__device__ __forceinline__ void dir_masks(const u64 *PL, int d0, int a, const DirConst &c,
                                          u64 &present, u64 &near, u64 &far) {
    present = PL[PL_D * BPS(d0) + a] & PL[c.off_n + wrapT<TD0>(a + c.da_n, d0)];
    near = PL[c.off_e + wrapT<TD0>(a + c.da_e, d0)];
    far = PL[c.off_f + wrapT<TD0>(a + c.da_f, d0)];
}
// counts of the four kinds (natural, missing-metal, missing-comb, missing-both), 16 bits each
__device__ __forceinline__ void kind_counts(u64 present, u64 near, u64 far, uint32_t &pk01, uint32_t &pk23) {
    const u64 pn = present & near, pf = present & far;
    const uint32_t n3 = __popcll(pn & far), n1 = __popcll(pn) - n3, n2 = __popcll(pf) - n3;
    const uint32_t n0 = __popcll(present) - n1 - n2 - n3;
    pk01 = n0 | (n1 << 16);
    pk23 = n2 | (n3 << 16);
}
__device__ __forceinline__ u64 kind_select(u64 present, u64 near, u64 far, int q) {
    return present & ((q & 1) ? near : ~near) & ((q & 2) ? far : ~far);
}
// CreateTrefoil anchored in row `a`: Up / Down masks
template <int TD0, int TD1>
__device__ __forceinline__ void trefoil_masks(const u64 *PL, int d0, int a, const RowBits<TD1> &rb, u64 &up, u64 &dn) {
    const u64 Da = PL[PL_D * BPS(d0) + a];
    const u64 Db = PL[PL_D * BPS(d0) + wrapT<TD0>(a + 2, d0)];
    const u64 both = Da & Db;
    up = both & rb.rotl(Da, rb.amount(2));
    dn = both & rb.rotl(Db, rb.amount(-2));
}

// MoveMonovacancy, row `a`, direction k: the neighbour row's Z comes from the pre-rotated plane, its M1 / M2 are rotated
// on the fly.  own / nbr masks for the site search and the counts (natural | merge << 16, swap) for the tables.
struct MMRow { u64 m1, m2, dv, zn, m1n, m2n; };
template <int TD0, int TD1>
__device__ __forceinline__ MMRow bp_mm_row(const u64 *PL, int d0, int a, int k, const RowBits<TD1> &rb) {
    const int an = wrapT<TD0>(a + nib(NBR_DA, k), d0), db = nib(NBR_DB, k);
    MMRow r;
    r.m1 = PL[PL_M1 * BPS(d0) + a]; r.m2 = PL[PL_M2 * BPS(d0) + a]; r.dv = PL[PL_D * BPS(d0) + a];
    r.zn = PL[(PL_Z + (db > 0 ? 1 : db < 0 ? 2 : 0)) * BPS(d0) + an];
    r.m1n = rb.rotl(PL[PL_M1 * BPS(d0) + an], rb.amount(db));
    r.m2n = rb.rotl(PL[PL_M2 * BPS(d0) + an], rb.amount(db));
    return r;
}
__device__ __forceinline__ void bp_mm_counts(const MMRow &r, uint32_t &nm, uint32_t &sw) {
    nm = (uint32_t)__popcll((r.m1 | r.m2) & r.zn) | ((uint32_t)__popcll((r.m1 & r.m2n) | (r.m2 & r.m1n)) << 16);
    sw = (uint32_t)__popcll(r.dv & (r.m1n | r.m2n));
}
// all class counts of row `a` (full enumeration of one row; used at launch start and by validate)
template <int TD0, int TD1, bool MM = false>
__device__ __forceinline__ void bp_row_counts(const u64 *PL, int d0, int a, const RowBits<TD1> &rb, uint32_t out[NCA]) {
    const uint32_t pz = __popcll(PL[PL_Z * BPS(d0) + a]), pd = __popcll(PL[PL_D * BPS(d0) + a]);
    const uint32_t pm = __popcll(PL[PL_M1 * BPS(d0) + a] | PL[PL_M2 * BPS(d0) + a]);
    const uint32_t pa = __popcll(PL[PL_A4 * BPS(d0) + a] | PL[PL_A7 * BPS(d0) + a]);
    out[C_CREATE_DIV] = pz; out[C_FILL_DIV] = pd; out[C_DESTROY_TREF] = pa;
    out[C_CMONO_FROM_DOUBLE] = 2 * pz; out[C_CMONO_MAKE_EMPTY] = pm; out[C_FMONO_MAKE_DOUBLE] = pm;
    out[C_FMONO_FROM_EMPTY] = 2 * pd; out[C_FLIP] = pm;
    uint32_t k01 = 0, k23 = 0;
    for (int k = 0; k < 6; k++) {
        const DirConst c = dir_const(k, d0);
        u64 p, n, f; dir_masks<TD0>(PL, d0, a, c, p, n, f);
        uint32_t a01, a23; kind_counts(p, n, f, a01, a23);
        k01 += a01; k23 += a23;
    }
    out[C_MOVE_DIV] = k01 & 0xFFFF; out[C_MOVE_DIV + 1] = k01 >> 16;
    out[C_MOVE_DIV + 2] = k23 & 0xFFFF; out[C_MOVE_DIV + 3] = k23 >> 16;
    u64 up, dn; trefoil_masks<TD0>(PL, d0, a, rb, up, dn);
    out[C_CREATE_TREF] = __popcll(up) + __popcll(dn);
    if (MM) {
        uint32_t nm = 0, sw = 0;
        for (int k = 0; k < 6; k++) {
            uint32_t x, y; bp_mm_counts(bp_mm_row<TD0>(PL, d0, a, k, rb), x, y);
            nm += x; sw += y;
        }
        out[C_MOVE_MONO] = nm & 0xFFFF; out[C_MOVE_MONO + 1] = nm >> 16;
        out[C_MOVE_MONO + 2] = 2 * ((k01 & 0xFFFF) + (k01 >> 16) + (k23 & 0xFFFF) + (k23 >> 16));
        out[C_MOVE_MONO + 3] = sw;
    }
}

// position of the (idx)-th set bit (0-based) of a row mask, found by 32 lanes in parallel:
// lane l counts the set bits below site 2l+2.  Returns the column.
// `low_lo/low_hi` = this lane's constant mask of the sites below 2l+2.
__device__ __forceinline__ int nth_set_bit(u64 m, uint32_t idx, int lane, uint32_t low_lo, uint32_t low_hi) {
    const uint32_t mlo = (uint32_t)m, mhi = (uint32_t)(m >> 32);
    const uint32_t c = __popc(mlo & low_lo) + __popc(mhi & low_hi);
    const uint32_t ball = __ballot_sync(FULL, c > idx);
    const int L = __ffs(ball) - 1;
    const uint32_t below = __shfl_sync(FULL, c, (L + 31) & 31);          // count below site 2L (lane L-1)
    const uint32_t half = L & 16 ? mhi : mlo;                            // sites 2L, 2L+1 live in one half
    const uint32_t first = (half >> ((2 * L) & 31)) & 1u;
    return 2 * L + (int)!(first && idx == (L ? below : 0u));
}

// Rule.perform + nodes_affected_by as a table (reference demo1/rules.py perform()/nodes_affected_by of the
// 8 rules): per slot, how the anchor's status changes and which other nodes change to what.
//   word0: [0:4) or-mask, [4:8) and-mask, [8:12) xor-mask of the anchor status  (new = ((s | or) & and) ^ xor)
//          [12:14) number of touched nodes
//   word1: [0:4) da1+2 [4:8) db1+2 [8:12) da2+2 [12:16) db2+2 [16:20) new1 [20:24) new2 [24:28) old1 [28:32) old2
struct MoveEntry { uint32_t w0, w1; };
constexpr MoveEntry make_move_entry(int slot) {
    uint32_t orm = 0, andm = 0, xorm = 0, na = 1;
    int da1 = 0, db1 = 0, da2 = 0, db2 = 0;
    uint32_t n1 = 0, n2 = 0, o1 = 0, o2 = 0;
    if (slot == 0) { xorm = 3; }                                    // -> divacancy
    else if (slot == 1) { }                                         // -> pristine
    else if (slot <= 3) { orm = slot - 1; andm = 0xF; }             // s | layer
    else if (slot <= 5) { andm = 0xF & ~(uint32_t)(slot - 3); }     // s & ~layer
    else if (slot == 6) { andm = 0xF; xorm = 3; }                   // flip layer
    else if (slot <= 12) {                                          // MoveDivacancy: target = nbr_k
        constexpr int NA[6] = {-1, 0, 1, 0, -1, 1}, NB[6] = {1, -1, 0, 1, 0, -1};
        da1 = NA[slot - 7]; db1 = NB[slot - 7]; na = 2; n1 = 3; o1 = 0;
    } else if (slot >= MM_SLOT0) {                                  // MoveMonovacancy (direction k, layer L)
        constexpr int NA[6] = {-1, 0, 1, 0, -1, 1}, NB[6] = {1, -1, 0, 1, 0, -1};
        const int k = (slot - MM_SLOT0) >> 1, layer = ((slot - MM_SLOT0) & 1) + 1;
        da1 = NA[k]; db1 = NB[k]; na = 2; andm = 0xF & ~(uint32_t)layer;   // source: s & ~L
        n1 = 0; o1 = 0;                                             // target: old | L, filled in from the kind
    } else {
        const int o = (slot - 13) & 1;
        const bool create = slot < 15;
        na = 3; da1 = 2; db1 = 0; da2 = o ? 2 : 0; db2 = o ? -2 : 2;
        xorm = create ? 4 + 3 * o : 3;
        n1 = create ? 4 + 3 * o + 1 : 3; n2 = create ? 4 + 3 * o + 2 : 3;
        o1 = create ? 3 : 4 + 3 * o + 1; o2 = create ? 3 : 4 + 3 * o + 2;
    }
    MoveEntry e = {0, 0};
    e.w0 = orm | (andm << 4) | (xorm << 8) | (na << 12);
    e.w1 = (uint32_t)(da1 + 2) | ((uint32_t)(db1 + 2) << 4) | ((uint32_t)(da2 + 2) << 8) | ((uint32_t)(db2 + 2) << 12) |
           (n1 << 16) | (n2 << 20) | (o1 << 24) | (o2 << 28);
    return e;
}
__constant__ MoveEntry MOVE_TABLE[NSA] = {
    make_move_entry(0), make_move_entry(1), make_move_entry(2), make_move_entry(3), make_move_entry(4),
    make_move_entry(5), make_move_entry(6), make_move_entry(7), make_move_entry(8), make_move_entry(9),
    make_move_entry(10), make_move_entry(11), make_move_entry(12), make_move_entry(13), make_move_entry(14),
    make_move_entry(15), make_move_entry(16), make_move_entry(17), make_move_entry(18), make_move_entry(19),
    make_move_entry(20), make_move_entry(21), make_move_entry(22), make_move_entry(23), make_move_entry(24),
    make_move_entry(25), make_move_entry(26), make_move_entry(27), make_move_entry(28)};

// position of the (n)-th set bit (0-based) of a small mask; n is warp-uniform and tiny
__device__ __forceinline__ int nth_bit_small(uint32_t m, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) m &= m - 1;
    return __ffs(m) - 1;
}

// LAT = true: the latency build for small batches and single trajectories -- same code, but the register
// allocator is not held to 64 registers for occupancy that such launches cannot use
// MM = true: MoveMonovacancy's four classes as well (17 classes on 32 weight lanes, three more row tables).
template <int TD0, int TD1, bool LAT = false, bool MM = false>
__global__ void __launch_bounds__(256, LAT ? 2 : 4) kmc_replica_bp_kernel(const ReplicaParams p)
{
    constexpr int NCX = MM ? NCA : NC;                 // classes
    constexpr int NL = MM ? 32 : 16;                   // weight lanes of the class choice
    const int d0 = TD0 ? TD0 : p.d0;
    const int d1 = TD1 ? TD1 : p.d1;
    const int n = d0 * d1;
    const int d0e = (d0 + 1) & ~1;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * (blockDim.x >> 5) + warp;
    if (rep >= p.n_replicas) return;       // whole warp exits together; no block barriers are used
    const RowBits<TD1> rb(d1);

    extern __shared__ uint4 smem_raw[];
    u64 *PL = reinterpret_cast<u64 *>(reinterpret_cast<uint8_t *>(smem_raw) + (size_t)warp * replica_bp_warp_bytes(d0, MM));
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(PL + N_PLANES * BPS(d0));      // [NC][d0e]
    const uint32_t *rc32 = reinterpret_cast<const uint32_t *>(rc16);

    // ---- load: status bytes -> bit planes (one row per lane at a time) --------------------------
    uint8_t *gst = p.status + (size_t)rep * n;
    for (int a = lane; a < d0; a += 32) {
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
        PL[PL_Z * BPS(d0) + a] = w[T_Z];   PL[PL_ZP * BPS(d0) + a] = rb.rotl(w[T_Z], rb.amount(1));
        PL[PL_ZM * BPS(d0) + a] = rb.rotl(w[T_Z], rb.amount(-1));
        PL[PL_D * BPS(d0) + a] = w[T_D];   PL[PL_DP * BPS(d0) + a] = rb.rotl(w[T_D], rb.amount(1));
        PL[PL_DM * BPS(d0) + a] = rb.rotl(w[T_D], rb.amount(-1));
        PL[PL_M1 * BPS(d0) + a] = w[T_M1]; PL[PL_M2 * BPS(d0) + a] = w[T_M2];
        PL[PL_A4 * BPS(d0) + a] = w[T_A4]; PL[PL_A7 * BPS(d0) + a] = w[T_A7];
    }
    __syncwarp();
    // ---- build the row hierarchy (full enumeration, as Rule.initialize_moves does) ---------------
    for (int a = lane; a < d0e; a += 32) {
        uint32_t cnt[NCA];
        if (a < d0) bp_row_counts<TD0, TD1, MM>(PL, d0, a, rb, cnt);
#pragma unroll
        for (int c = 0; c < NC; c++) rc16[c * d0e + a] = a < d0 ? (uint16_t)cnt[c] : (uint16_t)0;
        if (MM) {
            rc16[13 * d0e + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO] : (uint16_t)0;
            rc16[14 * d0e + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO + 1] : (uint16_t)0;
            rc16[15 * d0e + a] = a < d0 ? (uint16_t)cnt[C_MOVE_MONO + 3] : (uint16_t)0;
        }
    }
    __syncwarp();

    // per-lane class state: lane c owns class c; the other weight lanes are permanent zero weights
    const int myc = lane < NCX ? lane : 0;
    const double rate = (lane < NCX && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NCX && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    const int my_tab = lane < NC ? lane : ((MM && lane < NCX) ? bp_table_of(lane) : -1);
    int tot = 0;
    if (my_tab >= 0)
        for (int a = 0; a < d0; a++) tot += rc16[my_tab * d0e + a];
    else if (MM && lane == C_MOVE_MONO + 2)
        for (int a = 0; a < d0; a++)
            tot += 2 * (rc16[C_MOVE_DIV * d0e + a] + rc16[(C_MOVE_DIV + 1) * d0e + a] + rc16[(C_MOVE_DIV + 2) * d0e + a] +
                        rc16[(C_MOVE_DIV + 3) * d0e + a]);
    // own-status classes: 2 bits per status = moves of my class a site of that status carries
    const uint32_t my_g1 = (lane < NC && !(lane >= C_MOVE_DIV && lane <= C_CREATE_TREF)) ? g1_const(lane) : 0u;
    unsigned long long hist = 0, hops = 0, n_ties = 0;
    PickState pstate = pick_state_init_n<NL>(lane);
    double tclock = p.clock ? p.clock[rep] : 0.0;          // continues across launches in event order
    uint32_t *const tracer = p.tracer ? p.tracer + (size_t)rep * n : nullptr;
    // refresh roles: lane = 6*j + k -> (row slot j, direction k)
    const int my_j = lane / 6, my_k = lane % 6;
    DirConst my_dir = dir_const(my_k, d0);
    // keep the six per-lane constants in registers: ptxas otherwise re-derives them from my_k inside the
    // event loop (~30 instructions per event, profiles/round1_r6_replica_bp_by_line.txt)
    asm volatile("" : "+r"(my_dir.off_n), "+r"(my_dir.da_n), "+r"(my_dir.off_e), "+r"(my_dir.da_e),
                      "+r"(my_dir.off_f), "+r"(my_dir.da_f));
    // this lane's mask of the sites below column 2*lane + 2 (for popcount-select)
    uint32_t my_low_lo = lane >= 15 ? 0xFFFFFFFFu : ((4u << (2 * lane)) - 1u);
    uint32_t my_low_hi = lane < 16 ? 0u : (lane == 31 ? 0xFFFFFFFFu : ((4u << (2 * (lane - 16))) - 1u));
    asm volatile("" : "+r"(my_low_lo), "+r"(my_low_hi));
    // plane maintenance role: lane q < 10 keeps plane q; a site's bit sits at column col + my_cshift
    const int my_ptype = lane < N_PLANES ? (int)((PLANE_TYPE >> (4 * lane)) & 0xF) : -1;
    const int my_cshift = lane < N_PLANES ? (int)((PLANE_CSHIFT >> (4 * lane)) & 0xF) - 1 : 0;

    const unsigned long long step_base = p.steps_done[rep];
    double mu1 = 0.0, mu2 = 0.0;
    long long t = 0;
    uint32_t flag = 0;

    for (; t < p.nsteps; t++) {
        // ---- draws: a block of 32 events at a time, one event per lane ----------------------
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

        // ---- 1./2. weights and class choice (sim.py:120-122, kmc.py:21-54) ----------------------
        const double w = __dmul_rn((double)tot, rate);
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
        if (p.clock) tclock = __dadd_rn(tclock, kmc_time_step(p.clock_mode, pick.total, p.seed, p.replica0 + (uint32_t)rep,
                                                            step_base + (unsigned long long)t));
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) +
                                  ((size_t)rep * (size_t)p.nsteps + (size_t)t);
            if (lane < NCA) rec->counts[lane] = lane < NCX ? (uint32_t)tot : 0u;
        }

        // ---- 3. in-class pick: rank in (row, slot, column) order -----------------------------------
        const int cnt = __shfl_sync(FULL, tot, csel);
        uint32_t r;
        {
            long long rr = (long long)__dmul_rn(u2, (double)cnt);
            if (rr > (long long)cnt - 1) rr = (long long)cnt - 1;
            r = (uint32_t)rr;
        }
        // 3a. row: two rows per lane per pass (one 32-bit load of two uint16 counts)
        int row = 0;
        const bool split = MM && csel == C_MOVE_MONO + 2;
        const int tab = MM ? (split ? C_MOVE_DIV : bp_table_of(csel)) : csel;
        for (int base = 0; base < d0; base += 64) {
            const int r0 = base + 2 * lane;
            uint32_t v01 = (r0 < d0) ? rc32[(tab * d0e + r0) >> 1] : 0u;
            if (split && r0 < d0)          // twice the four MoveDivacancy tables (packed pairs, no carries: <= 3072)
                v01 = 2u * (v01 + rc32[((C_MOVE_DIV + 1) * d0e + r0) >> 1] + rc32[((C_MOVE_DIV + 2) * d0e + r0) >> 1] +
                            rc32[((C_MOVE_DIV + 3) * d0e + r0) >> 1]);
            const uint32_t v0 = v01 & 0xFFFFu, v1 = v01 >> 16;
            int L;
            if (warp_rank_search(v0 + v1, r, L, lane)) {
                const uint32_t v0L = __shfl_sync(FULL, v0, L);
                row = base + 2 * L;
                if (r >= v0L) { r -= v0L; row++; }
                break;
            }
        }
        // 3b/3c. slot, then site, within the row (canonical order: row, slot, column)
        int col, slot, s0;
        if (MM && csel >= C_MOVE_MONO) {
            // lanes 0..11 hold the mask of slot 17 + lane = (direction lane >> 1, layer (lane & 1) + 1)
            const MMRow mr = bp_mm_row<TD0>(PL, d0, row, lane < 12 ? lane >> 1 : 0, rb);
            const bool l2 = lane & 1;
            const int q = csel - C_MOVE_MONO;
            const u64 src = q < 2 ? (l2 ? mr.m2 : mr.m1) : mr.dv;
            const u64 dst = (q & 1) ? (l2 ? mr.m1n : mr.m2n) : mr.zn;
            const u64 mine = lane < 12 ? (src & dst) : 0ull;
            const uint32_t cntk = __popcll(mine);
            uint32_t pre = cntk, t1;
            t1 = __shfl_up_sync(FULL, pre, 1); if (lane >= 1) pre += t1;
            t1 = __shfl_up_sync(FULL, pre, 2); if (lane >= 2) pre += t1;
            t1 = __shfl_up_sync(FULL, pre, 4); if (lane >= 4) pre += t1;
            t1 = __shfl_up_sync(FULL, pre, 8); if (lane >= 8) pre += t1;
            const int ksel = __ffs(__ballot_sync(FULL, pre > r) & 0xFFFu) - 1;
            const uint32_t rem = r - __shfl_sync(FULL, pre - cntk, ksel);
            col = nth_set_bit(__shfl_sync(FULL, mine, ksel), rem, lane, my_low_lo, my_low_hi);
            slot = MM_SLOT0 + ksel;
            s0 = q < 2 ? (ksel & 1) + 1 : S_DIVAC;
        } else if (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) {
            // lanes 0..5 hold their direction's kind mask for this row; a 6-lane prefix picks the direction
            u64 pr, ne, fa;
            dir_masks<TD0>(PL, d0, row, my_dir, pr, ne, fa);
            const u64 mine = lane < 6 ? kind_select(pr, ne, fa, csel - C_MOVE_DIV) : 0ull;
            const uint32_t cntk = __popcll(mine);
            uint32_t pre = cntk, t1;
            t1 = __shfl_up_sync(FULL, pre, 1); if (lane >= 1) pre += t1;
            t1 = __shfl_up_sync(FULL, pre, 2); if (lane >= 2) pre += t1;
            t1 = __shfl_up_sync(FULL, pre, 4); if (lane >= 4) pre += t1;
            const int ksel = __ffs(__ballot_sync(FULL, pre > r) & 0x3Fu) - 1;
            const uint32_t rem = r - __shfl_sync(FULL, pre - cntk, ksel);
            col = nth_set_bit(__shfl_sync(FULL, mine, ksel), rem, lane, my_low_lo, my_low_hi);
            slot = 7 + ksel;
            s0 = S_DIVAC;
        } else if (csel == C_CREATE_TREF) {
            u64 up, dn;
            trefoil_masks<TD0>(PL, d0, row, rb, up, dn);
            const uint32_t nu = __popcll(up);
            const bool second = r >= nu;
            col = nth_set_bit(second ? dn : up, second ? r - nu : r, lane, my_low_lo, my_low_hi);
            slot = 13 + (int)second;
            s0 = S_DIVAC;
        } else {
            // own-status classes: one or two slots, each slot's sites are one plane (or a union)
            //   class:        0   1   7    8   9    10   11  12
            //   first slot:   Z   D   A4   Z   M2   M1   D   M1|M2
            //   second slot:  -   -   A7   Z   M1   M2   D   -
            int pa, pb = -1;
            switch (csel) {
                case C_CREATE_DIV: pa = PL_Z; break;
                case C_FILL_DIV: pa = PL_D; break;
                case C_DESTROY_TREF: pa = PL_A4; pb = PL_A7; break;
                case C_CMONO_FROM_DOUBLE: pa = PL_Z; pb = PL_Z; break;
                case C_CMONO_MAKE_EMPTY: pa = PL_M2; pb = PL_M1; break;
                case C_FMONO_MAKE_DOUBLE: pa = PL_M1; pb = PL_M2; break;
                case C_FMONO_FROM_EMPTY: pa = PL_D; pb = PL_D; break;
                default: pa = PL_M1; break;
            }
            u64 ma = PL[pa * BPS(d0) + row];
            const u64 mb = pb >= 0 ? PL[pb * BPS(d0) + row] : 0ull;
            if (csel == C_FLIP) ma |= PL[PL_M2 * BPS(d0) + row];
            const uint32_t n1 = __popcll(ma);
            const bool second = r >= n1;
            const u64 m = second ? mb : ma;
            col = nth_set_bit(m, second ? r - n1 : r, lane, my_low_lo, my_low_hi);
            slot = class_slot_base(csel) + (int)second;
            // status of the chosen site before the event
            switch (csel) {
                case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0 = 0; break;
                case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0 = 3; break;
                case C_DESTROY_TREF: s0 = second ? 7 : 4; break;
                case C_CMONO_MAKE_EMPTY: s0 = second ? 1 : 2; break;
                case C_FMONO_MAKE_DOUBLE: s0 = second ? 2 : 1; break;
                default: s0 = ((PL[PL_M1 * BPS(d0) + row] >> col) & 1ull) ? 1 : 2; break;
            }
        }
        const int site = row * d1 + col;

        // ---- event record -----------------------------------------------------------------------------
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
        if (lane == slot - 7 && slot < 13) hops++;             // lanes 0..5: divacancy hops per direction
        if (pick.tie && lane == 0) n_ties++;

        // ---- 4. the move (uniform): touched nodes, old and new status --------------------------------
        MoveNodes mv;
        int os1, os2;                              // old status of nodes 1, 2 (node 0: s0)
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
            if (MM && slot >= MM_SLOT0) {
                // the target gains the layer: pristine -> L (natural, split), opposite monovacancy -> divacancy
                const int layer = ((slot - MM_SLOT0) & 1) + 1;
                const bool onto_mono = (csel - C_MOVE_MONO) & 1;
                os1 = onto_mono ? 3 - layer : 0;
                mv.ns1 = onto_mono ? 3 : layer;
            }
        }
        const int na = mv.na;
        if (tracer && lane == 31)
            tracer_update(tracer, slot, na, row * d1 + col, mv.ns0, mv.a1 * d1 + mv.b1, mv.ns1, mv.a2 * d1 + mv.b2, mv.ns2, s0);
        // bit planes: lane q < 10 maintains plane q, as 32-bit halves (program order handles nodes
        // sharing a word): clear the site's bit, set it again if the new status belongs to this plane
        if (lane < N_PLANES) {
            uint32_t *P32 = reinterpret_cast<uint32_t *>(PL + lane * BPS(d0));
            {
                const int c = wrapT<TD1>(col + my_cshift, d1);
                uint32_t *wp = P32 + 2 * row + (c >> 5);
                const uint32_t bit = 1u << (c & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns0) == my_ptype ? bit : 0u);
            }
            if (na > 1) {
                const int c = wrapT<TD1>(mv.b1 + my_cshift, d1);
                uint32_t *wp = P32 + 2 * mv.a1 + (c >> 5);
                const uint32_t bit = 1u << (c & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns1) == my_ptype ? bit : 0u);
            }
            if (na > 2) {
                const int c = wrapT<TD1>(mv.b2 + my_cshift, d1);
                uint32_t *wp = P32 + 2 * mv.a2 + (c >> 5);
                const uint32_t bit = 1u << (c & 31);
                *wp = (*wp & ~bit) | (status_plane(mv.ns2) == my_ptype ? bit : 0u);
            }
        }
        // ---- 5a. own-status classes: owner lanes fold (old -> new) of each touched node -----------------
        if (my_g1) {
            int d = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
            if (d) rc16[lane * d0e + row] = (uint16_t)(rc16[lane * d0e + row] + d);
            int dsum = d;
            if (na > 1) {
                d = (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
                if (d) rc16[lane * d0e + mv.a1] = (uint16_t)(rc16[lane * d0e + mv.a1] + d);
                dsum += d;
            }
            if (na > 2) {
                d = (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
                if (d) rc16[lane * d0e + mv.a2] = (uint16_t)(rc16[lane * d0e + mv.a2] + d);
                dsum += d;
            }
            tot += dsum;
        }
        __syncwarp();

        // ---- 5b. MoveDivacancy: rows within +-1 of the touched rows, re-derived and replaced -------------
        // touched rows relative to `row`: 0 (always), da of the target (moves), +2 (trefoils)
        int lo = -1, hi = 1;
        if (slot >= 13 && slot < MM_SLOT0) hi = 3;
        else if (slot >= 7) { lo += mv.da1 < 0 ? mv.da1 : 0; hi += mv.da1 > 0 ? mv.da1 : 0; }
        {
            const bool act = my_j <= hi - lo;
            const int ra = wrapT<TD0>(row + lo + my_j, d0);
            uint32_t pk01 = 0, pk23 = 0, pknm = 0, pksw = 0;
            if (act) {
                u64 pr, ne, fa;
                dir_masks<TD0>(PL, d0, ra, my_dir, pr, ne, fa);
                kind_counts(pr, ne, fa, pk01, pk23);
                if (MM) bp_mm_counts(bp_mm_row<TD0>(PL, d0, ra, my_k, rb), pknm, pksw);
            }
            pk01 += __shfl_down_sync(FULL, pk01, 3);
            pk23 += __shfl_down_sync(FULL, pk23, 3);
            pk01 += __shfl_down_sync(FULL, pk01, 1) + __shfl_down_sync(FULL, pk01, 2);
            pk23 += __shfl_down_sync(FULL, pk23, 1) + __shfl_down_sync(FULL, pk23, 2);
            if (MM) {
                pknm += __shfl_down_sync(FULL, pknm, 3);
                pksw += __shfl_down_sync(FULL, pksw, 3);
                pknm += __shfl_down_sync(FULL, pknm, 1) + __shfl_down_sync(FULL, pknm, 2);
                pksw += __shfl_down_sync(FULL, pksw, 1) + __shfl_down_sync(FULL, pksw, 2);
            }
            uint32_t d01 = 0, d23 = 0, dnm = 0, dsw = 0;
            if (act && my_k == 0) {
                uint16_t *q = rc16 + C_MOVE_DIV * d0e + ra;
                const uint32_t o0 = q[0], o1 = q[d0e], o2 = q[2 * d0e], o3 = q[3 * d0e];
                q[0] = (uint16_t)pk01; q[d0e] = (uint16_t)(pk01 >> 16);
                q[2 * d0e] = (uint16_t)pk23; q[3 * d0e] = (uint16_t)(pk23 >> 16);
                // signed 16-bit pair deltas as one integer each (hi*65536 + lo)
                d01 = pk01 - (o0 | (o1 << 16));
                d23 = pk23 - (o2 | (o3 << 16));
                if (MM) {                                           // tables 13, 14, 15 = classes 13, 14, 16
                    uint16_t *m = rc16 + 13 * d0e + ra;
                    const uint32_t o5 = m[0], o6 = m[d0e], o7 = m[2 * d0e];
                    m[0] = (uint16_t)pknm; m[d0e] = (uint16_t)(pknm >> 16); m[2 * d0e] = (uint16_t)pksw;
                    dnm = pknm - (o5 | (o6 << 16));
                    dsw = pksw - o7;
                }
            }
            const int s01 = (int)__reduce_add_sync(FULL, d01);
            const int s23 = (int)__reduce_add_sync(FULL, d23);
            const int l01 = (int)(short)(s01 & 0xFFFF), l23 = (int)(short)(s23 & 0xFFFF);
            const int h01 = (s01 - l01) >> 16, h23 = (s23 - l23) >> 16;
            if (lane == C_MOVE_DIV) tot += l01;
            if (lane == C_MOVE_DIV + 1) tot += h01;
            if (lane == C_MOVE_DIV + 2) tot += l23;
            if (lane == C_MOVE_DIV + 3) tot += h23;
            if (MM) {
                const int snm = (int)__reduce_add_sync(FULL, dnm);
                const int ssw = (int)__reduce_add_sync(FULL, dsw);
                const int lnm = (int)(short)(snm & 0xFFFF);
                if (lane == C_MOVE_MONO) tot += lnm;
                if (lane == C_MOVE_MONO + 1) tot += (snm - lnm) >> 16;
                if (lane == C_MOVE_MONO + 2) tot += 2 * (l01 + h01 + l23 + h23);      // split = 2 x MoveDivacancy hops
                if (lane == C_MOVE_MONO + 3) tot += ssw;
            }
        }
        // ---- 5c. CreateTrefoil: anchor rows row-3 .. row+2 (covers t and t-2 of every touched row t) -----
        {
            const int ng = d0 < 6 ? d0 : 6;
            int d = 0;
            if (lane < ng) {
                const int ra = wrapT<TD0>(row + lane - 3, d0);
                u64 up, dn;
                trefoil_masks<TD0>(PL, d0, ra, rb, up, dn);
                const int nw = __popcll(up) + __popcll(dn);
                uint16_t *q = rc16 + C_CREATE_TREF * d0e + ra;
                d = nw - (int)q[0];
                q[0] = (uint16_t)nw;
            }
            const int s = (int)__reduce_add_sync(FULL, (uint32_t)d);
            if (lane == C_CREATE_TREF) tot += s;
        }
        __syncwarp();
    }

    // ---- epilogue ------------------------------------------------------------------------------------
    const long long done = t;
    if (p.validate) {
        // KmcSim.validate (sim.py:159-168): incremental tables vs a fresh enumeration
        bool bad = false;
        for (int a = lane; a < d0; a += 32) {
            uint32_t cnt[NCA];
            bp_row_counts<TD0, TD1, MM>(PL, d0, a, rb, cnt);
#pragma unroll
            for (int c = 0; c < NC; c++) bad |= (rc16[c * d0e + a] != (uint16_t)cnt[c]);
            if (MM) {
                bad |= rc16[13 * d0e + a] != (uint16_t)cnt[C_MOVE_MONO];
                bad |= rc16[14 * d0e + a] != (uint16_t)cnt[C_MOVE_MONO + 1];
                bad |= rc16[15 * d0e + a] != (uint16_t)cnt[C_MOVE_MONO + 3];
            }
            // plane consistency: at most one type bit per site; rotated copies agree with their base
            const u64 z = PL[PL_Z * BPS(d0) + a], dv = PL[PL_D * BPS(d0) + a];
            const u64 types[N_TYPES] = {z, dv, PL[PL_M1 * BPS(d0) + a], PL[PL_M2 * BPS(d0) + a], PL[PL_A4 * BPS(d0) + a], PL[PL_A7 * BPS(d0) + a]};
            u64 seen = 0, dup = 0;
#pragma unroll
            for (int q = 0; q < N_TYPES; q++) { dup |= seen & types[q]; seen |= types[q]; }
            bad |= dup != 0;
            bad |= PL[PL_ZP * BPS(d0) + a] != rb.rotl(z, rb.amount(1)) || PL[PL_ZM * BPS(d0) + a] != rb.rotl(z, rb.amount(-1));
            bad |= PL[PL_DP * BPS(d0) + a] != rb.rotl(dv, rb.amount(1)) || PL[PL_DM * BPS(d0) + a] != rb.rotl(dv, rb.amount(-1));
        }
        if (my_tab >= 0) {
            int fresh = 0;
            for (int a = 0; a < d0; a++) fresh += rc16[my_tab * d0e + a];
            bad |= (fresh != tot);
        } else if (MM && lane == C_MOVE_MONO + 2) {
            int fresh = 0;
            for (int a = 0; a < d0; a++)
                fresh += 2 * (rc16[C_MOVE_DIV * d0e + a] + rc16[(C_MOVE_DIV + 1) * d0e + a] + rc16[(C_MOVE_DIV + 2) * d0e + a] +
                              rc16[(C_MOVE_DIV + 3) * d0e + a]);
            bad |= (fresh != tot);
        }
        if (__any_sync(FULL, bad)) flag |= FLAG_VALIDATE;
    }
    // bit planes -> status bytes.  Pass 1: every site from its own plane bit (non-anchor trefoil
    // members get a placeholder); pass 2: each anchor writes the codes of its two partners.
    for (int a = lane; a < d0; a += 32) {
        const u64 z = PL[PL_Z * BPS(d0) + a], dv = PL[PL_D * BPS(d0) + a], m1 = PL[PL_M1 * BPS(d0) + a], m2 = PL[PL_M2 * BPS(d0) + a];
        const u64 a4 = PL[PL_A4 * BPS(d0) + a], a7 = PL[PL_A7 * BPS(d0) + a];
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
    __syncwarp();
    for (int a = lane; a < d0; a += 32) {
        u64 a4 = PL[PL_A4 * BPS(d0) + a], a7 = PL[PL_A7 * BPS(d0) + a];
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
