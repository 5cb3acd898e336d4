// kmc_common.cuh -- lattice geometry, class/slot tables and the per-site move-count function
// shared by the sweep kernel (K1/K2) and the fused replica kernel (K3/K4).
//
// Everything here is a from-scratch device formulation of what the reference derives at run time:
//   stencils           reference demo1/state.py:392-401,443-459 (+ hexagonal.py:56-62)
//   per-site moves     reference demo1/rules.py all_moves()/moves_dependent_on() of the 8 rules
// The functions marked KMC_HD also compile for the host so tests can exercise them without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KMC_HD __host__ __device__ __forceinline__
#else
#define KMC_HD inline
#endif

namespace kmc {

// Two widths.  NC / NS: the classes and slots of the rules the reference implements itself -- what the
// bit-sliced replica kernels and the aligned sweep kernels maintain.  NCA / NSA: those plus MoveMonovacancy
// (4 kinds, 12 (direction, layer) slots; a stub in the reference, rules.py:295-296, semantics in DESIGN.md
// section 2) -- the width of every table that crosses the C ABI and of the general (byte-lattice) kernels.
constexpr int NC = 13;        // event classes of the reference's own rules
constexpr int NS = 17;        // move slots per site of the reference's own rules
constexpr int NCA = 17;       // all event classes (ABI width)
constexpr int NSA = 29;       // all move slots per site (ABI width)
constexpr int MM_SLOT0 = 17;  // MoveMonovacancy slot of (direction k, layer L): 17 + 2k + (L - 1)
// row-count words per row of the byte-lattice kernel (two uint16 fields each): 13 classes -> 7, 17 -> 9
template <bool MM> struct RcWords { static constexpr int value = MM ? 9 : 7; };
constexpr int RC_WORDS = 7;

// status codes
constexpr int S_PRISTINE = 0, S_DIVAC = 3, S_TREF = 4;

// class ids
enum : int { C_CREATE_DIV = 0, C_FILL_DIV = 1, C_MOVE_DIV = 2, C_CREATE_TREF = 6, C_DESTROY_TREF = 7,
             C_CMONO_FROM_DOUBLE = 8, C_CMONO_MAKE_EMPTY = 9, C_FMONO_MAKE_DOUBLE = 10,
             C_FMONO_FROM_EMPTY = 11, C_FLIP = 12,
             C_MOVE_MONO = 13 /* natural, 14 merge-divacancy, 15 split-divacancy, 16 swap-divacancy */ };

// The six neighbours in ring order (60-degree rotations, state.py:392-394):
//   R0=(-1,0) R1=(0,-1) R2=(1,-1) R3=(1,0) R4=(0,1) R5=(-1,1)
// neighbors_and_mutuals (state.py:443-459) yields direction k with nbr = R[KRING[k]]; for odd ring
// positions near = R[i-1], far = R[i+1]; for even ones near = R[i+1], far = R[i-1].
// KRING = {5, 1, 3, 4, 0, 2}
KMC_HD int kring(int k) { return (0x204315 >> (4 * k)) & 7; }

// ---- per-site class counts, byte-packed -------------------------------------------------------
// F4.w[j] byte q = number of moves of class 4*j+q anchored at the site.
// (w4 byte 0 = class 16; only the MoveMonovacancy-aware build looks at it)
struct F4 { uint32_t w0, w1, w2, w3, w4; };

// contributions that depend on the site's own status only (classes 0,1,7,8..12)
KMC_HD F4 g1_of_status(int s) {
    F4 f = {0u, 0u, 0u, 0u, 0u};
    if (s == 0) { f.w0 = 0x00000001u; f.w2 = 0x00000002u; }                 // c0 x1, c8 x2
    else if (s == 1 || s == 2) { f.w2 = 0x00010100u; f.w3 = 0x00000001u; }  // c9, c10, c12
    else if (s == 3) { f.w0 = 0x00000100u; f.w2 = 0x02000000u; }            // c1 x1, c11 x2
    else if (s == 4 || s == 7) { f.w1 = 0x01000000u; }                      // c7 (trefoil anchor)
    return f;
}

// 6-bit ring masks -> per-kind direction masks of MoveDivacancy (rules.py:119-136)
struct RingMasks { uint32_t P, N, F; };   // pristine nbr, near is divacancy, far is divacancy
KMC_HD RingMasks ring_masks(int r0, int r1, int r2, int r3, int r4, int r5) {
    uint32_t P = (r0 == 0) | ((r1 == 0) << 1) | ((r2 == 0) << 2) | ((r3 == 0) << 3) | ((r4 == 0) << 4) |
                 ((r5 == 0) << 5);
    uint32_t D = (r0 == 3) | ((r1 == 3) << 1) | ((r2 == 3) << 2) | ((r3 == 3) << 3) | ((r4 == 3) << 4) |
                 ((r5 == 3) << 5);
    uint32_t Dm = ((D << 1) | (D >> 5)) & 0x3Fu;    // bit i = D[i-1]
    uint32_t Dp = ((D >> 1) | (D << 5)) & 0x3Fu;    // bit i = D[i+1]
    RingMasks m;
    m.P = P;
    m.N = (Dm & 0x2Au) | (Dp & 0x15u);
    m.F = (Dp & 0x2Au) | (Dm & 0x15u);
    return m;
}
// ring-position mask of the directions whose kind is `kind` (0 natural,1 metal,2 comb,3 both)
KMC_HD uint32_t kind_mask(const RingMasks &m, int kind) {
    uint32_t n = (kind & 1) ? m.N : ~m.N;
    uint32_t f = (kind & 2) ? m.F : ~m.F;
    return m.P & n & f & 0x3Fu;
}

KMC_HD int popc6(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// wrap helpers for offsets of magnitude <= 3 (dims >= 5)
KMC_HD int wrap_inc(int x, int d) { return x >= d ? x - d : x; }
KMC_HD int wrap_dec(int x, int d) { return x < 0 ? x + d : x; }

// Full per-site evaluation: byte-packed counts of all 13 classes for moves anchored at (a, b).
// `S` is any accessor with int operator()(int site_index).
// MoveMonovacancy counts of a site of status s0 in 1..3 from its six ring statuses: the vacancy in layer L hops
// to a neighbour whose layer set is defined and lacks L; kind = 2 * (s0 is a divacancy) + (target not pristine).
// Returns natural | merge << 8 | split << 16 | swap << 24.
KMC_HD uint32_t mm_counts(int s0, int r0, int r1, int r2, int r3, int r4, int r5) {
    const uint32_t p = (r0 == 0) + (r1 == 0) + (r2 == 0) + (r3 == 0) + (r4 == 0) + (r5 == 0);
    const uint32_t m1 = (r0 == 1) + (r1 == 1) + (r2 == 1) + (r3 == 1) + (r4 == 1) + (r5 == 1);
    const uint32_t m2 = (r0 == 2) + (r1 == 2) + (r2 == 2) + (r3 == 2) + (r4 == 2) + (r5 == 2);
    if (s0 == 1) return p | (m2 << 8);
    if (s0 == 2) return p | (m1 << 8);
    return ((2u * p) << 16) | ((m1 + m2) << 24);
}

template <bool MM = false, class S>
KMC_HD F4 eval_site(const S &s, int a, int b, int d0, int d1) {
    const int am = wrap_dec(a - 1, d0), ap = wrap_inc(a + 1, d0);
    const int bm = wrap_dec(b - 1, d1), bp = wrap_inc(b + 1, d1);
    const int s0 = s(a * d1 + b);
    F4 f = g1_of_status(s0);
    if (MM && s0 >= 1 && s0 <= 3) {
        const uint32_t c = mm_counts(s0, s(am * d1 + b), s(a * d1 + bm), s(ap * d1 + bm), s(ap * d1 + b),
                                     s(a * d1 + bp), s(am * d1 + bp));
        f.w3 |= (c & 0x00FFFFFFu) << 8;               // classes 13, 14, 15 = bytes 1..3 of word 3
        f.w4 = c >> 24;                               // class 16
    }
    if (s0 == S_DIVAC) {
        RingMasks m = ring_masks(s(am * d1 + b), s(a * d1 + bm), s(ap * d1 + bm), s(ap * d1 + b),
                                 s(a * d1 + bp), s(am * d1 + bp));
        uint32_t nat = popc6(kind_mask(m, 0)), met = popc6(kind_mask(m, 1));
        uint32_t cmb = popc6(kind_mask(m, 2)), bth = popc6(kind_mask(m, 3));
        f.w0 |= (nat << 16) | (met << 24);
        f.w1 |= cmb | (bth << 8);
        const int ap2 = wrap_inc(a + 2, d0), bp2 = wrap_inc(b + 2, d1), bm2 = wrap_dec(b - 2, d1);
        uint32_t t20 = s(ap2 * d1 + b) == S_DIVAC;
        uint32_t up = t20 & (uint32_t)(s(a * d1 + bp2) == S_DIVAC);
        uint32_t dn = t20 & (uint32_t)(s(ap2 * d1 + bm2) == S_DIVAC);
        f.w1 |= (up + dn) << 16;
    }
    return f;
}

// number of class-c moves in a byte-packed F4
KMC_HD uint32_t f4_class(const F4 &f, int c) {
    uint32_t w = (c < 4) ? f.w0 : (c < 8) ? f.w1 : (c < 12) ? f.w2 : (c < 16) ? f.w3 : f.w4;
    return (w >> (8 * (c & 3))) & 0xFFu;
}

// row-count field of class c inside a row's 14 uint16 fields (see replica kernel): classes of one
// packed byte-word j land in pair-words 2j (bytes 0,2) and 2j+1 (bytes 1,3)
KMC_HD int rc_field(int c) { return 4 * (c >> 2) + 2 * (c & 1) + ((c >> 1) & 1); }

// slot -> class of the move in that slot, given the status of the site and (for MoveDivacancy) the
// ring masks.  Used by the slot-plane writer and by tests.
// ---- Philox4x32-10 ------------------------------------------------------------------------------
KMC_HD void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// draw recipe shared with the oracle (oracle/philox.py)
KMC_HD void kmc_draw(uint64_t seed, uint32_t replica, uint64_t step, double &u1, double &u2) {
    uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = (uint32_t)(seed >> 32), c3 = 0u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, replica);
    u1 = (double)((((uint64_t)c0 << 32) | c1) >> 11) * 0x1.0p-53;
    u2 = (double)((((uint64_t)c2 << 32) | c3) >> 11) * 0x1.0p-53;
}

// ---- stochastic KMC clock (extension) -----------------------------------------
// The reference draws no time step (it logs total_rate "so that time can be reconstructed", README.md:36-37,
// sim.py:138).  The BKL residence time of an event is dt = -ln(u3) / total_rate with a third uniform draw:
//     (x0, x1, ..) = Philox4x32-10(counter = (step lo, step hi, seed hi, 3), key = (seed lo, replica))
//     u3 = ((((x0 << 32) | x1) >> 11) + 1) * 2^-53        in (0, 1]  (never 0: ln stays finite)
// kmc_log is a fixed sequence of IEEE fp64 +, -, *, / (no FMA contraction: --fmad=false here,
// -ffp-contract=off in the oracle), so the device, the host and oracle/kmc_oracle.c produce the same bits;
// it is within 2 ulp of the correctly rounded logarithm (tests/test_hostcheck.py).
KMC_HD double kmc_draw3(uint64_t seed, uint32_t replica, uint64_t step) {
    uint32_t c0 = (uint32_t)step, c1 = (uint32_t)(step >> 32), c2 = (uint32_t)(seed >> 32), c3 = 3u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, replica);
    return (double)(((((uint64_t)c0 << 32) | c1) >> 11) + 1ull) * 0x1.0p-53;
}
KMC_HD double kmc_log(double x) {
    unsigned long long bits;
#ifdef __CUDA_ARCH__
    bits = (unsigned long long)__double_as_longlong(x);
#else
    __builtin_memcpy(&bits, &x, 8);
#endif
    int e = (int)(bits >> 52) - 1023;                       // x is a positive normal number
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
#ifdef __CUDA_ARCH__
    m = __longlong_as_double((long long)bits);
#else
    __builtin_memcpy(&m, &bits, 8);
#endif
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }    // m in (0.7071, 1.4143]
    const double s = (m - 1.0) / (m + 1.0);                 // |s| <= 0.1716
    const double z = s * s;
    // ln m = 2 atanh s = 2 s (1 + z/3 + z^2/5 + ... + z^11/23), Horner from the top
    double q = 1.0 / 23.0;
    q = q * z + 1.0 / 21.0; q = q * z + 1.0 / 19.0; q = q * z + 1.0 / 17.0; q = q * z + 1.0 / 15.0;
    q = q * z + 1.0 / 13.0; q = q * z + 1.0 / 11.0; q = q * z + 1.0 / 9.0;  q = q * z + 1.0 / 7.0;
    q = q * z + 1.0 / 5.0;  q = q * z + 1.0 / 3.0;  q = q * z + 1.0;
    const double lnm = (s + s) * q;
    const double ed = (double)e;
    return ed * 0x1.62e42fee00000p-1 + (ed * 0x1.a39ef35793c76p-33 + lnm);   // ln 2 = hi + lo, hi has 21 zero bits
}
// one event's time increment: mode 1 = mean residence time 1 / R, mode 2 = -ln(u3) / R
KMC_HD double kmc_time_step(int mode, double total, uint64_t seed, uint32_t replica, uint64_t step) {
    if (mode == 2) return (0.0 - kmc_log(kmc_draw3(seed, replica, step))) / total;
    return 1.0 / total;
}

}  // namespace kmc
