// sweep_kernel.cuh -- K1 + K2: full-lattice enumerate + classify sweep (the `--no-incremental` path).
//
// Replaces Rule.initialize_moves x 8 rules (reference demo1/sim.py:204-206, rules.py:19-21,43-45 and
// each rule's all_moves) and ReverseDict.value_counts (incremental.py:287-294).
//
// in : status planes   uint8  [R][d0][d1]            (1 B/site read)
// out: slot planes     uint8  [R][17][d0*d1]         (17 B/site written; class id or 0xFF)
//      row counts      uint32 [R][d0][16]            (13 used)
//
// HBM-bound by construction: 18 algorithmic bytes per site.  The arithmetic is byte-parallel: one
// thread owns 4 consecutive sites of a row as one 32-bit word and evaluates all predicates with
// per-byte 0/1 flag words, so the instruction count per site stays far below the memory time.
// Neighbour rows come from L1/L2 (the status plane is 1/18 of the traffic and every word is re-read by
// a handful of neighbouring threads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kmc_common.cuh"

namespace kmc {

constexpr int RCG_STRIDE = 16;            // uint32 per row in the global row-count table
constexpr int ROLL_ACC_STRIDE = 32;       // uint64 per replica in the running class totals of the aligned kernels
constexpr uint32_t ONES = 0x01010101u;

// per-byte flags (0x01 / 0x00) of a word of four status bytes (status < 16)
struct ByteFlags { uint32_t zero, div, mono, up_anchor, dn_anchor; };
__host__ __device__ __forceinline__ ByteFlags byte_flags(uint32_t w) {
    const uint32_t e0 = w & ONES, e1 = (w >> 1) & ONES, e2 = (w >> 2) & ONES, e3 = (w >> 3) & ONES;
    ByteFlags f;
    const uint32_t hi = e2 | e3;
    f.zero = (e0 | e1 | hi) ^ ONES;
    f.div = e0 & e1 & ~hi;
    f.mono = (e0 ^ e1) & ~hi;
    f.up_anchor = e2 & ~(e0 | e1 | e3);          // status 4
    f.dn_anchor = e0 & e1 & e2 & ~e3;            // status 7
    return f;
}
// flags of the sites one / two to the right (+) or left (-) of each byte's site
__host__ __device__ __forceinline__ uint32_t shr_sites(uint32_t cur, uint32_t next, int k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(cur, next, 8 * k);
#else
    return k == 0 ? cur : (cur >> (8 * k)) | (next << (32 - 8 * k));
#endif
}
__host__ __device__ __forceinline__ uint32_t shl_sites(uint32_t prev, uint32_t cur, int k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(prev, cur, 8 * k);
#else
    return k == 0 ? cur : (cur << (8 * k)) | (prev >> (32 - 8 * k));
#endif
}
// code plane word: class id where flag set, 0xFF elsewhere
__host__ __device__ __forceinline__ uint32_t code(uint32_t flag, uint32_t cls_bytes) {
    const uint32_t m = flag * 0xFFu;
    return (cls_bytes & m) | ~m;
}
__host__ __device__ __forceinline__ uint32_t hsum(uint32_t bytes) { return (bytes * ONES) >> 24; }

// All 17 slot words + 13 class counts of one word of 4 sites.
// Inputs: status words of rows a-1 (cur,next), a (prev,cur,next), a+1 (prev,cur), a+2 (prev,cur).
struct WordIn { uint32_t m_cur, m_next, c_prev, c_cur, c_next, p_prev, p_cur, q_prev, q_cur; };

__host__ __device__ __forceinline__ void sweep_word(const WordIn &in, uint32_t out[NS], uint32_t cnt[NC]) {
    const ByteFlags c = byte_flags(in.c_cur);
    const ByteFlags cp = byte_flags(in.c_prev), cn = byte_flags(in.c_next);
    const ByteFlags m0 = byte_flags(in.m_cur), m1 = byte_flags(in.m_next);
    const ByteFlags p0 = byte_flags(in.p_cur), pm = byte_flags(in.p_prev);
    const ByteFlags q0 = byte_flags(in.q_cur), qm = byte_flags(in.q_prev);

    // own-status slots (rules.py:71,88,220-227,289; trefoil anchors for DestroyTrefoil :189)
    out[0] = code(c.zero, C_CREATE_DIV * ONES);
    out[1] = code(c.div, C_FILL_DIV * ONES);
    // CreateMonovacancy layer 1 (slot 2): status 0 -> from-double, status 2 -> make-empty
    const uint32_t e0 = in.c_cur & ONES, e1 = (in.c_cur >> 1) & ONES;
    const uint32_t s1 = c.mono & e0, s2 = c.mono & e1;          // status == 1 / == 2
    out[2] = code(c.zero | s2, (C_CMONO_FROM_DOUBLE * ONES) + s2);
    out[3] = code(c.zero | s1, (C_CMONO_FROM_DOUBLE * ONES) + s1);
    // FillMonovacancy layer 1 (slot 4): status 1 -> make-double, status 3 -> from-empty
    out[4] = code(c.div | s1, (C_FMONO_MAKE_DOUBLE * ONES) + c.div);
    out[5] = code(c.div | s2, (C_FMONO_MAKE_DOUBLE * ONES) + c.div);
    out[6] = code(c.mono, C_FLIP * ONES);
    cnt[C_CREATE_DIV] = hsum(c.zero);
    cnt[C_FILL_DIV] = hsum(c.div);
    cnt[C_CMONO_FROM_DOUBLE] = 2 * hsum(c.zero);
    cnt[C_CMONO_MAKE_EMPTY] = hsum(c.mono);
    cnt[C_FMONO_MAKE_DOUBLE] = hsum(c.mono);
    cnt[C_FMONO_FROM_EMPTY] = 2 * hsum(c.div);
    cnt[C_FLIP] = hsum(c.mono);

    // ring neighbours R0..R5 = (-1,0) (0,-1) (1,-1) (1,0) (0,1) (-1,1): pristine / divacancy flags
    uint32_t P[6], D[6];
    P[0] = m0.zero;                              D[0] = m0.div;
    P[1] = shl_sites(cp.zero, c.zero, 1);        D[1] = shl_sites(cp.div, c.div, 1);
    P[2] = shl_sites(pm.zero, p0.zero, 1);       D[2] = shl_sites(pm.div, p0.div, 1);
    P[3] = p0.zero;                              D[3] = p0.div;
    P[4] = shr_sites(c.zero, cn.zero, 1);        D[4] = shr_sites(c.div, cn.div, 1);
    P[5] = shr_sites(m0.zero, m1.zero, 1);       D[5] = shr_sites(m0.div, m1.div, 1);
    uint32_t k_nat = 0, k_met = 0, k_cmb = 0, k_bth = 0;      // per-byte counts (<= 6)
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int i = kring(k);
        const uint32_t near = (i & 1) ? D[(i + 5) % 6] : D[(i + 1) % 6];
        const uint32_t far = (i & 1) ? D[(i + 1) % 6] : D[(i + 5) % 6];
        const uint32_t present = c.div & P[i];
        out[7 + k] = code(present, (C_MOVE_DIV * ONES) + near + (far << 1));
        k_nat += present & ~near & ~far;
        k_met += present & near & ~far;
        k_cmb += present & ~near & far;
        k_bth += present & near & far;
    }
    cnt[C_MOVE_DIV + 0] = hsum(k_nat); cnt[C_MOVE_DIV + 1] = hsum(k_met);
    cnt[C_MOVE_DIV + 2] = hsum(k_cmb); cnt[C_MOVE_DIV + 3] = hsum(k_bth);

    // trefoils: Up(p) = {p, p+(2,0), p+(0,2)}, Down(p) = {p, p+(2,0), p+(2,-2)} (rules.py:159-174)
    const uint32_t t20 = q0.div;
    const uint32_t t02 = shr_sites(c.div, cn.div, 2);
    const uint32_t t2m2 = shl_sites(qm.div, q0.div, 2);
    const uint32_t up = c.div & t20 & t02, dn = c.div & t20 & t2m2;
    out[13] = code(up, C_CREATE_TREF * ONES);
    out[14] = code(dn, C_CREATE_TREF * ONES);
    out[15] = code(c.up_anchor, C_DESTROY_TREF * ONES);
    out[16] = code(c.dn_anchor, C_DESTROY_TREF * ONES);
    cnt[C_CREATE_TREF] = hsum(up) + hsum(dn);
    cnt[C_DESTROY_TREF] = hsum(c.up_anchor) + hsum(c.dn_anchor);
}

struct SweepParams {
    int d0, d1, n, n_replicas;
    const uint8_t *status;       // [R][n]
    uint8_t *slots;              // [R][17][n] or nullptr (count-only sweep)
    uint32_t *rowcnt;            // [R][d0][16], zeroed by the caller
    long long total_words;       // R * n / 4
};

// Fast path: d1 % 4 == 0.  One thread = one 32-bit word = 4 consecutive sites of a row.
__global__ void __launch_bounds__(256) kmc_sweep_kernel(const SweepParams p)
{
    const long long gw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gw < p.total_words;
    const int wpr = p.d1 >> 2;                              // words per row
    const int wpl = p.n >> 2;                               // words per lattice
    long long seg = -1 - (long long)(threadIdx.x & 31);     // row id for the count reduction
    uint32_t cnt[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) cnt[c] = 0;

    if (active) {
        const int rep = (int)(gw / wpl);
        const int wl = (int)(gw - (long long)rep * wpl);
        const int a = wl / wpr;
        const int wi = wl - a * wpr;
        const uint32_t *S = reinterpret_cast<const uint32_t *>(p.status + (size_t)rep * p.n);
        const int am = a == 0 ? p.d0 - 1 : a - 1;
        const int ap = a + 1 == p.d0 ? 0 : a + 1;
        const int ap2 = ap + 1 == p.d0 ? 0 : ap + 1;
        const int wm = wi == 0 ? wpr - 1 : wi - 1;
        const int wn = wi + 1 == wpr ? 0 : wi + 1;
        WordIn in;
        in.m_cur = __ldg(S + am * wpr + wi);  in.m_next = __ldg(S + am * wpr + wn);
        in.c_prev = __ldg(S + a * wpr + wm);  in.c_cur = __ldg(S + a * wpr + wi);  in.c_next = __ldg(S + a * wpr + wn);
        in.p_prev = __ldg(S + ap * wpr + wm); in.p_cur = __ldg(S + ap * wpr + wi);
        in.q_prev = __ldg(S + ap2 * wpr + wm); in.q_cur = __ldg(S + ap2 * wpr + wi);
        uint32_t out[NS];
        sweep_word(in, out, cnt);
        if (p.slots) {
            uint32_t *O = reinterpret_cast<uint32_t *>(p.slots + (size_t)rep * NS * p.n) + wl;
#pragma unroll
            for (int j = 0; j < NS; j++) __stcs(O + (size_t)j * wpl, out[j]);   // streaming stores
        }
        seg = (long long)rep * p.d0 + a;
    }
    // K2: per-row class counts.  Lanes of the same row pool their counts (16-bit pairs), the
    // lowest lane of each row adds them to the table.
    const uint32_t mask = __match_any_sync(0xFFFFFFFFu, seg);
    uint32_t pk[7];
#pragma unroll
    for (int j = 0; j < 6; j++) pk[j] = cnt[2 * j] | (cnt[2 * j + 1] << 16);
    pk[6] = cnt[12];
#pragma unroll
    for (int j = 0; j < 7; j++) pk[j] = __reduce_add_sync(mask, pk[j]);
    if (active && (__ffs(mask) - 1) == (int)(threadIdx.x & 31)) {
        uint32_t *R = p.rowcnt + (size_t)seg * RCG_STRIDE;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const uint32_t lo = pk[j] & 0xFFFFu, hi = pk[j] >> 16;
            if (lo) atomicAdd(R + 2 * j, lo);
            if (hi) atomicAdd(R + 2 * j + 1, hi);
        }
    }
}

// ---- 16 sites per thread ---------------------------------------------------------------------------
// Fast path for d1 % 16 == 0 with CTAs aligned to rows and replicas.  One thread owns 16 consecutive
// sites of a row (one 128-bit load per neighbour row, 17 128-bit streaming stores).  The per-byte flags
// of every loaded word are derived once; slot planes are emitted plane by plane (4 words -> one store)
// so few registers stay live; class counts are per-byte sums over the thread's words, pooled per row
// through shared memory, so row counts are written with plain stores (no memset, no global atomics)
// and class totals are finished by the last CTA of each replica (ticket) -- all in this one launch.
struct ZD { uint32_t zero, div; };
__host__ __device__ __forceinline__ ZD zd_flags(uint32_t w) {
    const uint32_t e0 = w & ONES, e1 = (w >> 1) & ONES, hi = ((w >> 2) | (w >> 3)) & ONES;
    ZD f; f.zero = (e0 | e1 | hi) ^ ONES; f.div = e0 & e1 & ~hi; return f;
}
struct ByteAcc { uint32_t zero, div, mono, anch, nat, met, cmb, bth, tref; };

// Flags of one thread's 16 sites and their halo: own row c[0..5] = prev word, 4 own words, next word;
// row a-1 m[0..4] = 4 words + next; row a+1 p[0..4] = prev + 4 words; row a+2 q[0..4] (divacancy only).
struct Chunk16 { uint32_t cw[4]; ZD c[6], m[5], p[5]; uint32_t q[5]; };

// pristine / divacancy flags of ring neighbour I (R0..R5) of the sites of word j
template <int I> __host__ __device__ __forceinline__ ZD ring_flags(const Chunk16 &k, int j) {
    ZD r;
    if (I == 0) r = k.m[j];
    else if (I == 1) { r.zero = shl_sites(k.c[j].zero, k.c[j + 1].zero, 1); r.div = shl_sites(k.c[j].div, k.c[j + 1].div, 1); }
    else if (I == 2) { r.zero = shl_sites(k.p[j].zero, k.p[j + 1].zero, 1); r.div = shl_sites(k.p[j].div, k.p[j + 1].div, 1); }
    else if (I == 3) r = k.p[j + 1];
    else if (I == 4) { r.zero = shr_sites(k.c[j + 1].zero, k.c[j + 2].zero, 1); r.div = shr_sites(k.c[j + 1].div, k.c[j + 2].div, 1); }
    else { r.zero = shr_sites(k.m[j].zero, k.m[j + 1].zero, 1); r.div = shr_sites(k.m[j].div, k.m[j + 1].div, 1); }
    return r;
}
// slot plane 7+K of word j (MoveDivacancy direction K), accumulating the kind counts
template <int K> __host__ __device__ __forceinline__ uint32_t move_plane(const Chunk16 &k, int j, ByteAcc &acc) {
    constexpr int I = (0x204315 >> (4 * K)) & 7;           // kring(K)
    constexpr int IN = (I & 1) ? (I + 5) % 6 : (I + 1) % 6, IF = (I & 1) ? (I + 1) % 6 : (I + 5) % 6;
    const uint32_t present = k.c[j + 1].div & ring_flags<I>(k, j).zero;
    const uint32_t near = ring_flags<IN>(k, j).div, far = ring_flags<IF>(k, j).div;
    const uint32_t pn = present & near, pf = present & far, pnf = pn & far;
    acc.bth += pnf; acc.met += pn ^ pnf; acc.cmb += pf ^ pnf; acc.nat += present & ~(near | far);
    return code(present, (C_MOVE_DIV * ONES) + near + (far << 1));
}
// own-status planes 0..6, 15, 16 of word j
__host__ __device__ __forceinline__ void own_planes(const Chunk16 &k, int j, uint32_t o[9], ByteAcc &acc) {
    const uint32_t cw = k.cw[j];
    const ZD c = k.c[j + 1];
    const uint32_t e0 = cw & ONES, e1 = (cw >> 1) & ONES, e2 = (cw >> 2) & ONES, e3 = (cw >> 3) & ONES;
    const uint32_t mono = (e0 ^ e1) & ~(e2 | e3);
    const uint32_t up_a = e2 & ~(e0 | e1 | e3), dn_a = e0 & e1 & e2 & ~e3;
    const uint32_t s1 = mono & e0, s2 = mono & e1;
    o[0] = code(c.zero, C_CREATE_DIV * ONES);
    o[1] = code(c.div, C_FILL_DIV * ONES);
    o[2] = code(c.zero | s2, (C_CMONO_FROM_DOUBLE * ONES) + s2);
    o[3] = code(c.zero | s1, (C_CMONO_FROM_DOUBLE * ONES) + s1);
    o[4] = code(c.div | s1, (C_FMONO_MAKE_DOUBLE * ONES) + c.div);
    o[5] = code(c.div | s2, (C_FMONO_MAKE_DOUBLE * ONES) + c.div);
    o[6] = code(mono, C_FLIP * ONES);
    o[7] = code(up_a, C_DESTROY_TREF * ONES);
    o[8] = code(dn_a, C_DESTROY_TREF * ONES);
    acc.zero += c.zero; acc.div += c.div; acc.mono += mono; acc.anch += up_a + dn_a;
}
// CreateTrefoil planes 13 (Up) and 14 (Down) of word j
__host__ __device__ __forceinline__ void trefoil_planes(const Chunk16 &k, int j, uint32_t &up, uint32_t &dn, ByteAcc &acc) {
    const uint32_t both = k.c[j + 1].div & k.q[j + 1];
    const uint32_t u = both & shr_sites(k.c[j + 1].div, k.c[j + 2].div, 2);
    const uint32_t d = both & shl_sites(k.q[j], k.q[j + 1], 2);
    acc.tref += u + d;
    up = code(u, C_CREATE_TREF * ONES);
    dn = code(d, C_CREATE_TREF * ONES);
}

struct Sweep16Params {
    int d0, d1, n, n_replicas;
    const uint8_t *status;             // [R][n]
    uint8_t *slots;                    // [R][17][n] or nullptr
    uint32_t *rowcnt;                  // [R][d0][16]
    unsigned long long *acc;           // [R][ROLL_ACC_STRIDE] running class totals, zero between launches
    unsigned int *ticket;              // [R] CTAs finished, zero between launches
    unsigned long long *counts;        // [R][13] class totals (output)
    int ctas_per_replica;
    int nplanes;                       // planes per replica in `slots`: 17, or 29 when the MoveMonovacancy planes follow
    uint32_t *rowmm;                   // MM build: [R][d0][4] row counts of the four MoveMonovacancy classes
};

#ifdef __CUDACC__
template <int K> __device__ __forceinline__ void emit_move_plane(const Chunk16 &k, uint4 *O, size_t cpl, ByteAcc &acc) {
    const uint32_t a = move_plane<K>(k, 0, acc), b = move_plane<K>(k, 1, acc);
    const uint32_t c = move_plane<K>(k, 2, acc), d = move_plane<K>(k, 3, acc);
    if (O) __stcs(O + (size_t)(7 + K) * cpl, make_uint4(a, b, c, d));
}

__global__ void __launch_bounds__(256, 3) kmc_sweep16_kernel(const Sweep16Params p)
{
    __shared__ uint32_t rows_sm[256 * RCG_STRIDE];          // [rows_in_cta][16], rows_in_cta <= 256
    const int tid = threadIdx.x, lane = tid & 31;
    const int cpr = p.d1 >> 4;                  // 16-site chunks per row (divides 256)
    const int cpl = p.n >> 4;                   // chunks per lattice (multiple of 256)
    const int rows_in_cta = 256 / cpr;
    const long long g = (long long)blockIdx.x * 256 + tid;
    const int rep = (int)(g / cpl);
    const int cl = (int)(g - (long long)rep * cpl);
    const int a = cl / cpr, ci = cl - a * cpr;
    const int row_local = tid / cpr;
    const int d0 = p.d0;

    // issue all loads first; the shared-memory zeroing and its barrier overlap their latency
    Chunk16 k;
    {
        const int am = a == 0 ? d0 - 1 : a - 1, ap = a + 1 == d0 ? 0 : a + 1, ap2 = ap + 1 == d0 ? 0 : ap + 1;
        const int cim = ci == 0 ? cpr - 1 : ci - 1, cin = ci + 1 == cpr ? 0 : ci + 1;
        const uint8_t *S = p.status + (size_t)rep * p.n;
        const uint4 *rm = reinterpret_cast<const uint4 *>(S + (size_t)am * p.d1);
        const uint4 *r0 = reinterpret_cast<const uint4 *>(S + (size_t)a * p.d1);
        const uint4 *rp = reinterpret_cast<const uint4 *>(S + (size_t)ap * p.d1);
        const uint4 *rq = reinterpret_cast<const uint4 *>(S + (size_t)ap2 * p.d1);
        const uint4 c4 = __ldg(r0 + ci), m4 = __ldg(rm + ci), p4 = __ldg(rp + ci), q4 = __ldg(rq + ci);
        const uint32_t c_prev = __ldg(reinterpret_cast<const uint32_t *>(r0 + cim) + 3);
        const uint32_t c_next = __ldg(reinterpret_cast<const uint32_t *>(r0 + cin));
        const uint32_t m_next = __ldg(reinterpret_cast<const uint32_t *>(rm + cin));
        const uint32_t p_prev = __ldg(reinterpret_cast<const uint32_t *>(rp + cim) + 3);
        const uint32_t q_prev = __ldg(reinterpret_cast<const uint32_t *>(rq + cim) + 3);
        for (int i = tid; i < rows_in_cta * RCG_STRIDE; i += 256) rows_sm[i] = 0;
        k.cw[0] = c4.x; k.cw[1] = c4.y; k.cw[2] = c4.z; k.cw[3] = c4.w;
        k.c[0] = zd_flags(c_prev); k.c[5] = zd_flags(c_next);
        k.c[1] = zd_flags(c4.x); k.c[2] = zd_flags(c4.y); k.c[3] = zd_flags(c4.z); k.c[4] = zd_flags(c4.w);
        k.m[0] = zd_flags(m4.x); k.m[1] = zd_flags(m4.y); k.m[2] = zd_flags(m4.z); k.m[3] = zd_flags(m4.w);
        k.m[4] = zd_flags(m_next);
        k.p[0] = zd_flags(p_prev); k.p[1] = zd_flags(p4.x); k.p[2] = zd_flags(p4.y); k.p[3] = zd_flags(p4.z);
        k.p[4] = zd_flags(p4.w);
        k.q[0] = zd_flags(q_prev).div; k.q[1] = zd_flags(q4.x).div; k.q[2] = zd_flags(q4.y).div;
        k.q[3] = zd_flags(q4.z).div; k.q[4] = zd_flags(q4.w).div;
    }
    ByteAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint4 *O = p.slots ? reinterpret_cast<uint4 *>(p.slots + (size_t)rep * p.nplanes * p.n) + cl : nullptr;
    {
        uint32_t o0[9], o1[9], o2[9], o3[9];
        own_planes(k, 0, o0, acc); own_planes(k, 1, o1, acc); own_planes(k, 2, o2, acc); own_planes(k, 3, o3, acc);
        if (O) {
#pragma unroll
            for (int j = 0; j < 7; j++) __stcs(O + (size_t)j * cpl, make_uint4(o0[j], o1[j], o2[j], o3[j]));
            __stcs(O + (size_t)15 * cpl, make_uint4(o0[7], o1[7], o2[7], o3[7]));
            __stcs(O + (size_t)16 * cpl, make_uint4(o0[8], o1[8], o2[8], o3[8]));
        }
    }
    emit_move_plane<0>(k, O, cpl, acc); emit_move_plane<1>(k, O, cpl, acc); emit_move_plane<2>(k, O, cpl, acc);
    emit_move_plane<3>(k, O, cpl, acc); emit_move_plane<4>(k, O, cpl, acc); emit_move_plane<5>(k, O, cpl, acc);
    {
        uint32_t u[4], d[4];
#pragma unroll
        for (int j = 0; j < 4; j++) trefoil_planes(k, j, u[j], d[j], acc);
        if (O) {
            __stcs(O + (size_t)13 * cpl, make_uint4(u[0], u[1], u[2], u[3]));
            __stcs(O + (size_t)14 * cpl, make_uint4(d[0], d[1], d[2], d[3]));
        }
    }
    // K2: this thread's 13 class counts as 16-bit pairs, pooled per row
    uint32_t pk[7];
    {
        const uint32_t z = hsum(acc.zero), dv = hsum(acc.div), mo = hsum(acc.mono);
        pk[0] = z | (dv << 16);                                   // classes 0, 1
        pk[1] = hsum(acc.nat) | (hsum(acc.met) << 16);            // 2, 3
        pk[2] = hsum(acc.cmb) | (hsum(acc.bth) << 16);            // 4, 5
        pk[3] = hsum(acc.tref) | (hsum(acc.anch) << 16);          // 6, 7
        pk[4] = (2 * z) | (mo << 16);                             // 8, 9
        pk[5] = mo | ((2 * dv) << 16);                            // 10, 11
        pk[6] = mo;                                               // 12
    }
    const uint32_t mask = cpr >= 32 ? 0xFFFFFFFFu : __match_any_sync(0xFFFFFFFFu, row_local);
#pragma unroll
    for (int j = 0; j < 7; j++) pk[j] = __reduce_add_sync(mask, pk[j]);
    __syncthreads();                                              // rows_sm zeroed by everyone
    if ((__ffs(mask) - 1) == lane) {
        uint32_t *R = rows_sm + row_local * RCG_STRIDE;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const uint32_t lo = pk[j] & 0xFFFFu, hi = pk[j] >> 16;
            if (lo) atomicAdd(R + 2 * j, lo);
            if (hi) atomicAdd(R + 2 * j + 1, hi);
        }
    }
    __syncthreads();
    // rows are complete inside this CTA: plain coalesced stores
    {
        const int first_row = a - row_local;
        uint32_t *G = p.rowcnt + ((size_t)rep * d0 + first_row) * RCG_STRIDE;
        for (int i = tid; i < rows_in_cta * RCG_STRIDE; i += 256) G[i] = rows_sm[i];
    }
    // class totals: CTA partial sums -> running totals; the last CTA of the replica publishes them.
    // Only warp 0 takes part; its lane 0 orders the 13 atomics before the ticket with one fence
    // (fence cumulativity through __syncwarp), instead of every warp paying for a gpu-scope fence.
    if (tid < 32) {
        unsigned long long s = 0;
        if (tid < NC)
            for (int r = 0; r < rows_in_cta; r++) s += rows_sm[r * RCG_STRIDE + tid];
        if (p.ctas_per_replica == 1) {
            if (tid < NC) p.counts[(size_t)rep * NCA + tid] = s;
        } else {
            if (tid < NC && s) atomicAdd(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], s);
            __syncwarp();
            unsigned int last = 0;
            if (tid == 0) {
                __threadfence();
                last = (atomicAdd(&p.ticket[rep], 1u) == (unsigned)p.ctas_per_replica - 1u);
                if (last) __threadfence();
            }
            last = __shfl_sync(0xFFFFFFFFu, last, 0);
            if (last) {
                if (tid < NC) p.counts[(size_t)rep * NCA + tid] = atomicExch(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], 0ull);
                if (tid == 0) p.ticket[rep] = 0u;
            }
        }
    }
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// K2 epilogue shared by the aligned kernels: per-row class counts from each thread's 16-bit count
// pairs, plain row stores, class totals via running totals + ticket (last CTA of a replica publishes).
template <int NP>
__device__ __forceinline__ void row_counts_epilogue(const Sweep16Params &p, uint32_t *rows_sm, uint32_t *rows_mm, uint32_t pk[NP],
                                                    int tid, int lane, int cpr, int rows_in_cta, int row_local,
                                                    int rep, int a)
{
    constexpr bool MM = NP > 7;
    constexpr int SCR = MM ? 12 : 8;                                        // scratch words per warp
    if (cpr >= 32) {
        // every warp lies inside one row: warp sums go to per-warp slots, no zeroing, no atomics
#pragma unroll
        for (int j = 0; j < NP; j++) pk[j] = __reduce_add_sync(0xFFFFFFFFu, pk[j]);
        if (lane < NP) {
            uint32_t v = pk[0];
#pragma unroll
            for (int j = 1; j < NP; j++) v = lane == j ? pk[j] : v;
            rows_sm[256 * RCG_STRIDE / 2 + (tid >> 5) * SCR + lane] = v;   // scratch in the upper half
        }
        __syncthreads();
        const int wpr = cpr >> 5;                                          // warps per row
        uint32_t s = 0, s_mm = 0;
        if (tid < rows_in_cta * RCG_STRIDE) {
            const int r = tid >> 4, f = tid & 15;
            if (f < 2 * 7)
                for (int w = 0; w < wpr; w++) {
                    const uint32_t v = rows_sm[256 * RCG_STRIDE / 2 + (r * wpr + w) * SCR + (f >> 1)];
                    s += (f & 1) ? (v >> 16) : (v & 0xFFFFu);
                }
        }
        if (MM && tid < rows_in_cta * 4) {
            const int r = tid >> 2, f = tid & 3;
            for (int w = 0; w < wpr; w++) {
                const uint32_t v = rows_sm[256 * RCG_STRIDE / 2 + (r * wpr + w) * SCR + 7 + (f >> 1)];
                s_mm += (f & 1) ? (v >> 16) : (v & 0xFFFFu);
            }
        }
        if (tid < rows_in_cta * RCG_STRIDE) rows_sm[tid] = (tid & 15) < NC ? s : 0u;   // (rows_in_cta <= 8 here: below the scratch)
        if (MM && tid < rows_in_cta * 4) rows_mm[tid] = s_mm;
        __syncthreads();
    } else {
        // rows shorter than a warp: a row is cpr consecutive lanes (cpr is a power of two), summed by xor shuffles inside
        // the group; its first lane stores the row's record (every row of the CTA has exactly one owner: no zeroing, no
        // atomics).  16-bit fields: a row of < 512 sites holds < 6144 moves of a class.
#pragma unroll
        for (int j = 0; j < NP; j++)
            for (int o = cpr >> 1; o >= 1; o >>= 1) pk[j] += __shfl_xor_sync(0xFFFFFFFFu, pk[j], o);
        if ((lane & (cpr - 1)) == 0) {
            uint32_t *R = rows_sm + row_local * RCG_STRIDE;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                R[2 * j] = pk[j] & 0xFFFFu;
                R[2 * j + 1] = 2 * j + 1 < NC ? pk[j] >> 16 : 0u;
            }
            R[14] = 0; R[15] = 0;
            if (MM) {
                uint32_t *R4 = rows_mm + row_local * 4;
#pragma unroll
                for (int j = 7; j < NP; j++) { R4[2 * (j - 7)] = pk[j] & 0xFFFFu; R4[2 * (j - 7) + 1] = pk[j] >> 16; }
            }
        }
        __syncthreads();
    }
    {
        // (the CTA's rows are consecutive in the [replica][row] tables, also when it spans several small replicas)
        const size_t first_row = (size_t)blockIdx.x * rows_in_cta;
        (void)a; (void)row_local;
        uint32_t *G = p.rowcnt + first_row * RCG_STRIDE;
        for (int i = tid; i < rows_in_cta * RCG_STRIDE; i += 256) G[i] = rows_sm[i];
        if (MM) {
            uint32_t *G4 = p.rowmm + first_row * 4;
            for (int i = tid; i < rows_in_cta * 4; i += 256) G4[i] = rows_mm[i];
        }
    }
    if (p.ctas_per_replica == 0) {
        // lattices of fewer than 4096 sites: the CTA holds 256 / (chunks per lattice) whole replicas; thread (local
        // replica, class) sums the replica's rows and publishes its total directly
        constexpr int NCX = MM ? NCA : NC;
        const int cpl = p.n >> 4, reps_in_cta = 256 / cpl, d0 = p.d0;
        for (int idx = tid; idx < reps_in_cta * NCX; idx += 256) {
            const int lr = idx / NCX, c = idx - lr * NCX;
            unsigned long long s = 0;
            for (int r = 0; r < d0; r++)
                s += c < NC ? rows_sm[(lr * d0 + r) * RCG_STRIDE + c] : rows_mm[(lr * d0 + r) * 4 + c - NC];
            p.counts[((size_t)blockIdx.x * reps_in_cta + lr) * NCA + c] = s;
        }
        return;
    }
    if (tid < 32) {
        constexpr int NCX = MM ? NCA : NC;
        unsigned long long s = 0;
        if (tid < NC)
            for (int r = 0; r < rows_in_cta; r++) s += rows_sm[r * RCG_STRIDE + tid];
        else if (MM && tid < NCA)
            for (int r = 0; r < rows_in_cta; r++) s += rows_mm[r * 4 + tid - NC];
        if (p.ctas_per_replica == 1) {
            if (tid < NCX) p.counts[(size_t)rep * NCA + tid] = s;
        } else {
            if (tid < NCX && s) atomicAdd(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], s);
            __syncwarp();
            unsigned int last = 0;
            if (tid == 0) {
                __threadfence();
                last = (atomicAdd(&p.ticket[rep], 1u) == (unsigned)p.ctas_per_replica - 1u);
                if (last) __threadfence();
            }
            last = __shfl_sync(0xFFFFFFFFu, last, 0);
            if (last) {
                if (tid < NCX) p.counts[(size_t)rep * NCA + tid] = atomicExch(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], 0ull);
                if (tid == 0) p.ticket[rep] = 0u;
            }
        }
    }
}
#endif

// ---- 16 sites per thread, nibble-parallel arithmetic, PRMT table output --------------------------------
// Second-generation aligned kernel.  The byte-parallel version above is bound by the ALU pipe (LOP3/SHF
// issue at half rate), not by HBM.  Here the 16 status bytes of a thread are packed to 4-bit fields
// (8 sites per 32-bit word), so every predicate op covers 8 sites, and each slot plane is produced by
// the byte-permute unit used as an 8-entry lookup table: the four low selector nibbles of a PRMT pick,
// per site, the plane's code byte for that site's status (own-status planes) or for its
// (present, near, far) triple (MoveDivacancy planes).  Counting uses per-nibble accumulators.
// slot-plane store flavour of the nibble kernels: 0 = st.global.cs (streaming / evict-first),
// 1 = default write-back, 2 = st.global.wt
#ifndef KMC_SLOT_STORE_KIND
#define KMC_SLOT_STORE_KIND 0
#endif
#if KMC_SLOT_STORE_KIND == 0
#define KMC_SLOT_STORE(ptr, val) __stcs((ptr), (val))
#elif KMC_SLOT_STORE_KIND == 1
#define KMC_SLOT_STORE(ptr, val) (*(ptr) = (val))
#else
#define KMC_SLOT_STORE(ptr, val) __stwt((ptr), (val))
#endif
constexpr uint32_t ONESN = 0x11111111u;
__host__ __device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (n & 7))) & 0xFF;
        if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
// 8 status bytes (two words) -> 8 nibbles
__host__ __device__ __forceinline__ uint32_t pack_nib(uint32_t x, uint32_t y) {
    return prmt(x | (x >> 4), y | (y >> 4), 0x6420);
}
__host__ __device__ __forceinline__ ZD nzd(uint32_t n) {
    const uint32_t e0 = n & ONESN, e1 = (n >> 1) & ONESN, hi = ((n >> 2) | (n >> 3)) & ONESN;
    ZD f; f.zero = (e0 | e1 | hi) ^ ONESN; f.div = e0 & e1 & ~hi; return f;
}
__host__ __device__ __forceinline__ uint32_t nshr(uint32_t cur, uint32_t next, int k) {   // site b -> b + k
#ifdef __CUDA_ARCH__
    return __funnelshift_r(cur, next, 4 * k);
#else
    return (cur >> (4 * k)) | (next << (32 - 4 * k));
#endif
}
__host__ __device__ __forceinline__ uint32_t nshl(uint32_t prev, uint32_t cur, int k) {   // site b -> b - k
#ifdef __CUDA_ARCH__
    return __funnelshift_l(prev, cur, 4 * k);
#else
    return (cur << (4 * k)) | (prev >> (32 - 4 * k));
#endif
}
// own row c[0..3] = prev (its top nibbles), two own words, next (its low nibbles); row a-1 m[0..2] = own, own,
// next; row a+1 p[0..2] = prev, own, own; row a+2 q[0..2] (divacancy flags only)
struct Chunk16N { uint32_t cn[2]; ZD c[4], m[3], p[3]; uint32_t q[3]; };
struct NibAcc { uint32_t pres, pn, pf, pnf; };      // per-nibble sums over directions and both words (<= 12)

template <int I> __host__ __device__ __forceinline__ ZD ring_flags_n(const Chunk16N &k, int j) {
    ZD r;
    if (I == 0) r = k.m[j];
    else if (I == 1) { r.zero = nshl(k.c[j].zero, k.c[j + 1].zero, 1); r.div = nshl(k.c[j].div, k.c[j + 1].div, 1); }
    else if (I == 2) { r.zero = nshl(k.p[j].zero, k.p[j + 1].zero, 1); r.div = nshl(k.p[j].div, k.p[j + 1].div, 1); }
    else if (I == 3) r = k.p[j + 1];
    else if (I == 4) { r.zero = nshr(k.c[j + 1].zero, k.c[j + 2].zero, 1); r.div = nshr(k.c[j + 1].div, k.c[j + 2].div, 1); }
    else { r.zero = nshr(k.m[j].zero, k.m[j + 1].zero, 1); r.div = nshr(k.m[j].div, k.m[j + 1].div, 1); }
    return r;
}
// MoveDivacancy direction K of nibble word j: two output words (sites 0-3, 4-7 of the word)
template <int K, bool STORE = true>
__host__ __device__ __forceinline__ void move_plane_n(const Chunk16N &k, int j, uint32_t &o0, uint32_t &o1, NibAcc &acc) {
    constexpr int I = (0x204315 >> (4 * K)) & 7;           // kring(K)
    constexpr int IN = (I & 1) ? (I + 5) % 6 : (I + 1) % 6, IF = (I & 1) ? (I + 1) % 6 : (I + 5) % 6;
    const uint32_t present = k.c[j + 1].div & ring_flags_n<I>(k, j).zero;
    const uint32_t near = ring_flags_n<IN>(k, j).div, far = ring_flags_n<IF>(k, j).div;
    const uint32_t pn = present & near, pf = present & far;
    acc.pres += present; acc.pn += pn; acc.pf += pf; acc.pnf += pn & far;
    // selector nibble = present + 2*near + 4*far -> code 2 + near + 2*far, or 0xFF when absent
    if (STORE) {
        const uint32_t sel = far * 4u + (near * 2u + present);
        o0 = prmt(0x03FF02FFu, 0x05FF04FFu, sel);
        o1 = prmt(0x03FF02FFu, 0x05FF04FFu, sel >> 16);
    }
}
__host__ __device__ __forceinline__ uint32_t nib_sum(uint32_t x) {      // sum of the 8 nibbles
    return (((x & 0x0F0F0F0Fu) + ((x >> 4) & 0x0F0F0F0Fu)) * ONES) >> 24;
}

#ifdef __CUDACC__
// STORE = false: the count-only sweep of the no-incremental path (no slot planes are built at all)
template <int K, bool STORE> __device__ __forceinline__ void emit_move_plane_n(const Chunk16N &k, uint4 *O, size_t cpl, NibAcc &acc) {
    uint32_t a = 0, b = 0, c = 0, d = 0;
    move_plane_n<K, STORE>(k, 0, a, b, acc);
    move_plane_n<K, STORE>(k, 1, c, d, acc);
    if (STORE && O) KMC_SLOT_STORE(O + (size_t)(7 + K) * cpl, make_uint4(a, b, c, d));
}
template <bool STORE>
__device__ __forceinline__ void emit_lut_plane(uint4 *O, size_t plane, size_t cpl, uint32_t lo, uint32_t hi,
                                               uint32_t s0, uint32_t s0h, uint32_t s1, uint32_t s1h) {
    if (STORE && O) KMC_SLOT_STORE(O + plane * cpl, make_uint4(prmt(lo, hi, s0), prmt(lo, hi, s0h), prmt(lo, hi, s1), prmt(lo, hi, s1h)));
}

// all 17 slot planes of one 16-site chunk (stored through O, stride cpl between planes) and the chunk's
// 13 class counts as seven 16-bit pairs
template <bool STORE = true>
__device__ __forceinline__ void sweep_chunk_n(const Chunk16N &k, uint4 *O, size_t cpl, uint32_t pk[7])
{
    // own-status planes: table lookup by status (8, 9 -> 5, 6: all of 5, 6, 8, 9 carry no own-status move)
    uint32_t n_zero, n_div, n_mono, n_anch;
    {
        const uint32_t e3a = (k.cn[0] >> 3) & ONESN, e3b = (k.cn[1] >> 3) & ONESN;
        const uint32_t s0 = k.cn[0] - 3u * e3a, s1 = k.cn[1] - 3u * e3b;
        const uint32_t s0h = s0 >> 16, s1h = s1 >> 16;
        const uint32_t F = 0xFFFFFFFFu;
        emit_lut_plane<STORE>(O, 0, cpl, 0xFFFFFF00u, F, s0, s0h, s1, s1h);      // CreateDivacancy      s0 -> 0
        emit_lut_plane<STORE>(O, 1, cpl, 0x01FFFFFFu, F, s0, s0h, s1, s1h);      // FillDivacancy        s3 -> 1
        emit_lut_plane<STORE>(O, 2, cpl, 0xFF09FF08u, F, s0, s0h, s1, s1h);      // CreateMono layer 1   s0 -> 8, s2 -> 9
        emit_lut_plane<STORE>(O, 3, cpl, 0xFFFF0908u, F, s0, s0h, s1, s1h);      // CreateMono layer 2   s0 -> 8, s1 -> 9
        emit_lut_plane<STORE>(O, 4, cpl, 0x0BFF0AFFu, F, s0, s0h, s1, s1h);      // FillMono layer 1     s1 -> 10, s3 -> 11
        emit_lut_plane<STORE>(O, 5, cpl, 0x0B0AFFFFu, F, s0, s0h, s1, s1h);      // FillMono layer 2     s2 -> 10, s3 -> 11
        emit_lut_plane<STORE>(O, 6, cpl, 0xFF0C0CFFu, F, s0, s0h, s1, s1h);      // FlipMonovacancy      s1, s2 -> 12
        emit_lut_plane<STORE>(O, 15, cpl, F, 0xFFFFFF07u, s0, s0h, s1, s1h);     // DestroyTrefoil Up    s4 -> 7
        emit_lut_plane<STORE>(O, 16, cpl, F, 0x07FFFFFFu, s0, s0h, s1, s1h);     // DestroyTrefoil Down  s7 -> 7
        // counts of the own-status classes from single-bit flags
        uint32_t mono = 0, anch = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t n = k.cn[j];
            const uint32_t e0 = n & ONESN, e1 = (n >> 1) & ONESN, e2 = (n >> 2) & ONESN, e3 = (n >> 3) & ONESN;
            mono += __popc((e0 ^ e1) & ~(e2 | e3));
            anch += __popc((e2 & ~(e0 | e1 | e3)) | (e0 & e1 & e2 & ~e3));
        }
        n_mono = mono; n_anch = anch;
        n_zero = __popc(k.c[1].zero) + __popc(k.c[2].zero);
        n_div = __popc(k.c[1].div) + __popc(k.c[2].div);
    }
    NibAcc acc = {0, 0, 0, 0};
    emit_move_plane_n<0, STORE>(k, O, cpl, acc); emit_move_plane_n<1, STORE>(k, O, cpl, acc); emit_move_plane_n<2, STORE>(k, O, cpl, acc);
    emit_move_plane_n<3, STORE>(k, O, cpl, acc); emit_move_plane_n<4, STORE>(k, O, cpl, acc); emit_move_plane_n<5, STORE>(k, O, cpl, acc);
    uint32_t n_tref;
    {
        uint32_t u[2], d[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t both = k.c[j + 1].div & k.q[j + 1];
            u[j] = both & nshr(k.c[j + 1].div, k.c[j + 2].div, 2);
            d[j] = both & nshl(k.q[j], k.q[j + 1], 2);
        }
        n_tref = __popc(u[0]) + __popc(u[1]) + __popc(d[0]) + __popc(d[1]);
        emit_lut_plane<STORE>(O, 13, cpl, 0xFFFF06FFu, 0xFFFFFFFFu, u[0], u[0] >> 16, u[1], u[1] >> 16);
        emit_lut_plane<STORE>(O, 14, cpl, 0xFFFF06FFu, 0xFFFFFFFFu, d[0], d[0] >> 16, d[1], d[1] >> 16);
    }
    // K2: this thread's 13 class counts as 16-bit pairs, pooled per row
    {
        const uint32_t P = nib_sum(acc.pres), PN = nib_sum(acc.pn), PF = nib_sum(acc.pf), PNF = nib_sum(acc.pnf);
        const uint32_t met = PN - PNF, cmb = PF - PNF, nat = P - PN - PF + PNF;
        pk[0] = n_zero | (n_div << 16);                           // classes 0, 1
        pk[1] = nat | (met << 16);                                // 2, 3
        pk[2] = cmb | (PNF << 16);                                // 4, 5
        pk[3] = n_tref | (n_anch << 16);                          // 6, 7
        pk[4] = (2 * n_zero) | (n_mono << 16);                    // 8, 9
        pk[5] = n_mono | ((2 * n_div) << 16);                     // 10, 11
        pk[6] = n_mono;                                           // 12
    }
}

// ---- MoveMonovacancy planes on the nibble layout (fused into the rolling-window kernel) ----------------------------
// Per site (nibble) three target-side flags travel together in one word T: bit 0 = its layer set is defined and lacks
// layer 1, bit 1 = defined and lacks layer 2, bit 2 = defined and not pristine.  One funnel shift moves all three to
// the neighbour's position.  With the source flags l1 (layer 1 vacant), l2 << 1 and the divacancy bit placed in the
// free bit, one LOP3 per layer gives the PRMT selector: layer 1 index = exists + 2 divacancy + 4 non-pristine, layer 2
// index = divacancy + 2 exists + 4 non-pristine; the table bytes are the class codes 13 + 2 divacancy + non-pristine.
constexpr uint32_t TWOSN = 0x22222222u, FOURSN = 0x44444444u;
__host__ __device__ __forceinline__ uint32_t mm_target_word(uint32_t n) {
    const uint32_t e0 = n & ONESN, e1 = (n >> 1) & ONESN, hi = ((n >> 2) | (n >> 3)) & ONESN;
    const uint32_t def = hi ^ ONESN;
    return (def & ~e0) | ((def & ~e1) << 1) | ((def & (e0 | e1)) << 2);
}
struct MMTargets { uint32_t t[4]; };          // target words of a row around the chunk: prev, own0, own1, next
struct MMNibAcc { uint32_t e1, e2, n1, n2; };  // per-nibble sums over the directions: exists (layer 1 | layer 2 at bit 1), same & non-pristine

#ifdef __CUDACC__
template <bool STORE>
__device__ __forceinline__ void mm_emit(uint4 *O, size_t cpl, int k, const uint32_t T[2], const uint32_t m1[2], const uint32_t m2[2],
                                        const uint32_t dv[2], MMNibAcc acc[2])
{
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t s1 = (T[j] & m1[j]) | (dv[j] << 1), s2 = (T[j] & m2[j]) | dv[j];
        acc[j].e1 += s1 & ONESN; acc[j].e2 += s2 & TWOSN;
        acc[j].n1 += s1 & (T[j] >> 2) & ONESN; acc[j].n2 += s2 & (T[j] >> 1) & TWOSN;
        if (STORE) {
            o1[2 * j] = prmt(0x0FFF0DFFu, 0x10FF0EFFu, s1); o1[2 * j + 1] = prmt(0x0FFF0DFFu, 0x10FF0EFFu, s1 >> 16);
            o2[2 * j] = prmt(0x0F0DFFFFu, 0x100EFFFFu, s2); o2[2 * j + 1] = prmt(0x0F0DFFFFu, 0x100EFFFFu, s2 >> 16);
        }
    }
    if (STORE && O) {
        KMC_SLOT_STORE(O + (size_t)(MM_SLOT0 + 2 * k) * cpl, make_uint4(o1[0], o1[1], o1[2], o1[3]));
        KMC_SLOT_STORE(O + (size_t)(MM_SLOT0 + 2 * k + 1) * cpl, make_uint4(o2[0], o2[1], o2[2], o2[3]));
    }
}
// planes 17..28 of one 16-site chunk and its four MoveMonovacancy class counts (natural | merge << 16, split | swap << 16)
// n0, n1: the chunk's own nibble words; rows a-1 (m), a (c), a+1 (p)
template <bool STORE>
__device__ __forceinline__ void sweep_chunk_mm(uint32_t n0, uint32_t n1, const MMTargets &m, const MMTargets &c, const MMTargets &p,
                                               uint4 *O, size_t cpl, uint32_t &pk_nm, uint32_t &pk_ss)
{
    uint32_t m1[2], m2[2], dv[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t n = j ? n1 : n0;
        const uint32_t e0 = n & ONESN, e1 = (n >> 1) & ONESN, def = (((n >> 2) | (n >> 3)) & ONESN) ^ ONESN;
        const uint32_t l1 = e0 & def, l2 = e1 & def;
        m1[j] = l1 | FOURSN; m2[j] = (l2 << 1) | FOURSN; dv[j] = l1 & l2;
    }
    MMNibAcc acc[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    uint32_t T[2];
    // neighbors_and_mutuals order: (-1,+1) (0,-1) (+1,0) (0,+1) (-1,0) (+1,-1); t[] = prev, own0, own1, next
    T[0] = nshr(m.t[1], m.t[2], 1); T[1] = nshr(m.t[2], m.t[3], 1); mm_emit<STORE>(O, cpl, 0, T, m1, m2, dv, acc);
    T[0] = nshl(c.t[0], c.t[1], 1); T[1] = nshl(c.t[1], c.t[2], 1); mm_emit<STORE>(O, cpl, 1, T, m1, m2, dv, acc);
    T[0] = p.t[1];                  T[1] = p.t[2];                  mm_emit<STORE>(O, cpl, 2, T, m1, m2, dv, acc);
    T[0] = nshr(c.t[1], c.t[2], 1); T[1] = nshr(c.t[2], c.t[3], 1); mm_emit<STORE>(O, cpl, 3, T, m1, m2, dv, acc);
    T[0] = m.t[1];                  T[1] = m.t[2];                  mm_emit<STORE>(O, cpl, 4, T, m1, m2, dv, acc);
    T[0] = nshl(p.t[0], p.t[1], 1); T[1] = nshl(p.t[1], p.t[2], 1); mm_emit<STORE>(O, cpl, 5, T, m1, m2, dv, acc);
    // per nibble: X = moves of the site (<= 12), XN = those onto a non-pristine neighbour; split by the source kind
    uint32_t sx = 0, sxn = 0, sxd = 0, sxnd = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t X = acc[j].e1 + (acc[j].e2 >> 1), XN = acc[j].n1 + (acc[j].n2 >> 1), dm = dv[j] * 15u;
        const uint32_t Xd = X & dm, XNd = XN & dm;
        sx += (X & 0x0F0F0F0Fu) + ((X >> 4) & 0x0F0F0F0Fu);      sxn += (XN & 0x0F0F0F0Fu) + ((XN >> 4) & 0x0F0F0F0Fu);
        sxd += (Xd & 0x0F0F0F0Fu) + ((Xd >> 4) & 0x0F0F0F0Fu);   sxnd += (XNd & 0x0F0F0F0Fu) + ((XNd >> 4) & 0x0F0F0F0Fu);
    }
    const uint32_t tx = (sx * ONES) >> 24, txn = (sxn * ONES) >> 24, txd = (sxd * ONES) >> 24, txnd = (sxnd * ONES) >> 24;
    pk_nm = (tx - txn - txd + txnd) | ((txn - txnd) << 16);       // natural | merge-divacancy
    pk_ss = (txd - txnd) | (txnd << 16);                          // split-divacancy | swap-divacancy
}
#endif


#ifndef KMC_SWEEP_OCC
#define KMC_SWEEP_OCC 4
#endif
template <bool MM = false>
__global__ void __launch_bounds__(256, MM ? 4 : KMC_SWEEP_OCC) kmc_sweep16n_kernel(const Sweep16Params p)
{
    __shared__ uint32_t rows_sm[256 * RCG_STRIDE];          // [rows_in_cta][16], rows_in_cta <= 256
    __shared__ uint32_t rows_mm[MM ? 256 * 4 : 1];          // MM: [rows_in_cta][4]
    const int tid = threadIdx.x, lane = tid & 31;
    const int cpr = p.d1 >> 4;                  // 16-site chunks per row (divides 256)
    const int cpl = p.n >> 4;                   // chunks per lattice (multiple of 256)
    const int rows_in_cta = 256 / cpr;
    const long long g = (long long)blockIdx.x * 256 + tid;
    const int rep = (int)(g / cpl);
    const int cl = (int)(g - (long long)rep * cpl);
    const int a = cl / cpr, ci = cl - a * cpr;
    const int row_local = tid / cpr;
    const int d0 = p.d0;

    Chunk16N k;
    MMTargets tm, tc, tp;
    {
        const int am = a == 0 ? d0 - 1 : a - 1, ap = a + 1 == d0 ? 0 : a + 1, ap2 = ap + 1 == d0 ? 0 : ap + 1;
        const int cim = ci == 0 ? cpr - 1 : ci - 1, cin = ci + 1 == cpr ? 0 : ci + 1;
        const uint8_t *S = p.status + (size_t)rep * p.n;
        const uint4 *rm = reinterpret_cast<const uint4 *>(S + (size_t)am * p.d1);
        const uint4 *r0 = reinterpret_cast<const uint4 *>(S + (size_t)a * p.d1);
        const uint4 *rp = reinterpret_cast<const uint4 *>(S + (size_t)ap * p.d1);
        const uint4 *rq = reinterpret_cast<const uint4 *>(S + (size_t)ap2 * p.d1);
        const uint4 c4 = __ldg(r0 + ci), m4 = __ldg(rm + ci), p4 = __ldg(rp + ci), q4 = __ldg(rq + ci);
        const uint32_t c_prev = __ldg(reinterpret_cast<const uint32_t *>(r0 + cim) + 3);
        const uint32_t c_next = __ldg(reinterpret_cast<const uint32_t *>(r0 + cin));
        const uint32_t m_next = __ldg(reinterpret_cast<const uint32_t *>(rm + cin));
        const uint32_t p_prev = __ldg(reinterpret_cast<const uint32_t *>(rp + cim) + 3);
        const uint32_t q_prev = __ldg(reinterpret_cast<const uint32_t *>(rq + cim) + 3);
        const uint32_t ncp = pack_nib(0u, c_prev), ncn = pack_nib(c_next, 0u);
        const uint32_t nm0 = pack_nib(m4.x, m4.y), nm1 = pack_nib(m4.z, m4.w), nmn = pack_nib(m_next, 0u);
        const uint32_t npp = pack_nib(0u, p_prev), np0 = pack_nib(p4.x, p4.y), np1 = pack_nib(p4.z, p4.w);
        k.cn[0] = pack_nib(c4.x, c4.y); k.cn[1] = pack_nib(c4.z, c4.w);
        k.c[0] = nzd(ncp); k.c[1] = nzd(k.cn[0]); k.c[2] = nzd(k.cn[1]); k.c[3] = nzd(ncn);
        k.m[0] = nzd(nm0); k.m[1] = nzd(nm1); k.m[2] = nzd(nmn);
        k.p[0] = nzd(npp); k.p[1] = nzd(np0); k.p[2] = nzd(np1);
        k.q[0] = nzd(pack_nib(0u, q_prev)).div; k.q[1] = nzd(pack_nib(q4.x, q4.y)).div;
        k.q[2] = nzd(pack_nib(q4.z, q4.w)).div;
        if (MM) {
            // MoveMonovacancy target words of rows a-1, a, a+1 (prev, own0, own1, next; the unused ends stay 0)
            tm.t[0] = 0; tm.t[1] = mm_target_word(nm0); tm.t[2] = mm_target_word(nm1); tm.t[3] = mm_target_word(nmn);
            tc.t[0] = mm_target_word(ncp); tc.t[1] = mm_target_word(k.cn[0]); tc.t[2] = mm_target_word(k.cn[1]); tc.t[3] = mm_target_word(ncn);
            tp.t[0] = mm_target_word(npp); tp.t[1] = mm_target_word(np0); tp.t[2] = mm_target_word(np1); tp.t[3] = 0;
        }
    }
    uint4 *O = p.slots ? reinterpret_cast<uint4 *>(p.slots + (size_t)rep * p.nplanes * p.n) + cl : nullptr;
    uint32_t pk[MM ? 9 : 7];
    sweep_chunk_n(k, O, (size_t)cpl, pk);
    if (MM) sweep_chunk_mm<true>(k.cn[0], k.cn[1], tm, tc, tp, O, (size_t)cpl, pk[MM ? 7 : 0], pk[MM ? 8 : 0]);
    row_counts_epilogue<MM ? 9 : 7>(p, rows_sm, rows_mm, pk, tid, lane, cpr, rows_in_cta, row_local, rep, a);
}
#endif  // __CUDACC__


#ifdef __CUDACC__
// ---- rolling window over rows -----------------------------------------------------------------------------
// For rows that fill a whole CTA (d1 = 4096: 256 threads x 16 sites).  A CTA walks `rows_per_cta`
// consecutive rows.  Each thread keeps the nibble-packed flags of its 16-site column for the four rows a-1,
// a, a+1, a+2 in registers and, per step, loads only row a+3 (one 128-bit load + two halo words, issued
// before the current row is computed so the latency is hidden): every status word is unpacked once per
// thread instead of four times, and the barrier / ticket epilogue is paid once per CTA.
struct RowRec { uint32_t n0, n1; ZD prev, own0, own1, next; };
struct RowRaw { uint4 own; uint32_t prev, next; };
__device__ __forceinline__ RowRaw load_row_raw(const uint8_t *S, int d1, int a, int ci, int cim, int cin) {
    const uint4 *r = reinterpret_cast<const uint4 *>(S + (size_t)a * d1);
    RowRaw w;
    w.own = __ldg(r + ci);
    w.prev = __ldg(reinterpret_cast<const uint32_t *>(r + cim) + 3);
    w.next = __ldg(reinterpret_cast<const uint32_t *>(r + cin));
    return w;
}
// MM: also the row's MoveMonovacancy target words (see mm_target_word)
template <bool MM>
__device__ __forceinline__ RowRec make_row(const RowRaw &w, MMTargets &t) {
    RowRec r;
    r.n0 = pack_nib(w.own.x, w.own.y); r.n1 = pack_nib(w.own.z, w.own.w);
    const uint32_t np = pack_nib(0u, w.prev), nn = pack_nib(w.next, 0u);
    r.prev = nzd(np); r.own0 = nzd(r.n0); r.own1 = nzd(r.n1); r.next = nzd(nn);
    if (MM) { t.t[0] = mm_target_word(np); t.t[1] = mm_target_word(r.n0); t.t[2] = mm_target_word(r.n1); t.t[3] = mm_target_word(nn); }
    return r;
}

struct SweepRollParams {
    int d0, d1, n, n_replicas;
    const uint8_t *status;
    uint8_t *slots;
    uint32_t *rowcnt;
    unsigned long long *acc;            // [R][ROLL_ACC_STRIDE] running class totals, zero between launches
    unsigned int *ticket;
    unsigned long long *counts;
    int rows_per_cta, ctas_per_replica;
    int nplanes;
    uint32_t *rowmm;                    // MM: [R][d0][4] row counts of the four MoveMonovacancy classes (plain stores)
};
constexpr int ROLL_MAX_ROWS = 16;

#ifndef KMC_ROLL_OCC
#define KMC_ROLL_OCC 3
#endif
// MM = true: every rule of config/WSe2.yaml in ONE pass -- the 17 planes and 13 classes above plus the twelve
// MoveMonovacancy planes (17..28) and its four classes from the same registers (30 algorithmic bytes per site).
template <bool STORE, bool MM = false>
__global__ void __launch_bounds__(256, MM ? 2 : KMC_ROLL_OCC) kmc_sweep_roll_kernel(const SweepRollParams p)
{
    __shared__ uint32_t part[ROLL_MAX_ROWS][8][MM ? 10 : 8];       // [row step][warp][7 (MM: 9) count pairs]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d0 = p.d0, d1 = p.d1;
    const int cpr = d1 >> 4;                              // == 256
    const int cpl = p.n >> 4;
    const int rep = blockIdx.x / p.ctas_per_replica;
    const int a0 = (blockIdx.x - rep * p.ctas_per_replica) * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, d0 - a0);
    const int ci = tid, cim = ci == 0 ? cpr - 1 : ci - 1, cin = ci + 1 == cpr ? 0 : ci + 1;
    const uint8_t *S = p.status + (size_t)rep * p.n;
    auto wrap = [d0](int a) { return a < 0 ? a + d0 : (a >= d0 ? a - d0 : a); };

    MMTargets tm, tc, tp, tq;
    RowRec rm = make_row<MM>(load_row_raw(S, d1, wrap(a0 - 1), ci, cim, cin), tm);
    RowRec rc = make_row<MM>(load_row_raw(S, d1, a0, ci, cim, cin), tc);
    RowRec rp = make_row<MM>(load_row_raw(S, d1, wrap(a0 + 1), ci, cim, cin), tp);
    RowRec rq = make_row<MM>(load_row_raw(S, d1, wrap(a0 + 2), ci, cim, cin), tq);
    for (int i = 0; i < nrows; i++) {
        const int a = a0 + i;
        // prefetch the row that enters the window next
        const RowRaw nxt = load_row_raw(S, d1, wrap(wrap(a + 3)), ci, cim, cin);
        Chunk16N k;
        k.cn[0] = rc.n0; k.cn[1] = rc.n1;
        k.c[0] = rc.prev; k.c[1] = rc.own0; k.c[2] = rc.own1; k.c[3] = rc.next;
        k.m[0] = rm.own0; k.m[1] = rm.own1; k.m[2] = rm.next;
        k.p[0] = rp.prev; k.p[1] = rp.own0; k.p[2] = rp.own1;
        k.q[0] = rq.prev.div; k.q[1] = rq.own0.div; k.q[2] = rq.own1.div;
        uint4 *O = (STORE && p.slots) ? reinterpret_cast<uint4 *>(p.slots + (size_t)rep * p.nplanes * p.n) + ((size_t)a * cpr + ci) : nullptr;
        uint32_t pk[MM ? 9 : 7];
        sweep_chunk_n<STORE>(k, O, (size_t)cpl, pk);
        if (MM) sweep_chunk_mm<STORE>(rc.n0, rc.n1, tm, tc, tp, O, (size_t)cpl, pk[MM ? 7 : 0], pk[MM ? 8 : 0]);
#pragma unroll
        for (int j = 0; j < (MM ? 9 : 7); j++) pk[j] = __reduce_add_sync(0xFFFFFFFFu, pk[j]);
        if (lane < (MM ? 9 : 7)) {
            uint32_t v = pk[0];
#pragma unroll
            for (int j = 1; j < (MM ? 9 : 7); j++) v = lane == j ? pk[j] : v;
            part[i][warp][lane] = v;
        }
        rm = rc; rc = rp; rp = rq;
        if (MM) { tm = tc; tc = tp; tp = tq; }
        rq = make_row<MM>(nxt, tq);
    }
    __syncthreads();
    // row counts: thread (row step i, field f) sums the 8 warps' partials; plain stores
    unsigned long long mine = 0;
    if (tid < nrows * RCG_STRIDE) {
        const int i = tid >> 4, f = tid & 15;
        uint32_t sum = 0;
        if (f < 14)
            for (int w = 0; w < 8; w++) { const uint32_t v = part[i][w][f >> 1]; sum += (f & 1) ? (v >> 16) : (v & 0xFFFFu); }
        sum = f < NC ? sum : 0u;
        p.rowcnt[((size_t)rep * d0 + a0 + i) * RCG_STRIDE + f] = sum;
        mine = sum;
    }
    // class totals of this CTA: reduce over its rows (threads with the same f), then running totals + ticket
    __shared__ unsigned long long tot_sm[20];
    if (tid < 20) tot_sm[tid] = 0;
    __syncthreads();
    if (tid < nrows * RCG_STRIDE && mine) atomicAdd(&tot_sm[tid & 15], mine);
    if (MM && tid < nrows * 4) {
        // the four MoveMonovacancy classes: their own row table (a 16-word row of `rowcnt` holds the other 13)
        const int i = tid >> 2, f = tid & 3;
        uint32_t sum = 0;
        for (int w = 0; w < 8; w++) { const uint32_t v = part[i][w][7 + (f >> 1)]; sum += (f & 1) ? (v >> 16) : (v & 0xFFFFu); }
        p.rowmm[((size_t)rep * d0 + a0 + i) * 4 + f] = sum;
        if (sum) atomicAdd(&tot_sm[16 + f], (unsigned long long)sum);
    }
    __syncthreads();
    if (tid < 32) {
        // lane c = class c (MM: classes 13..16 sit in tot_sm[16..19])
        constexpr int NCX = MM ? NCA : NC;
        const unsigned long long sum = tid < NC ? tot_sm[tid] : ((MM && tid < NCA) ? tot_sm[16 + tid - NC] : 0ull);
        if (p.ctas_per_replica == 1) {
            if (tid < NCX) p.counts[(size_t)rep * NCA + tid] = sum;
        } else {
            if (tid < NCX && sum) atomicAdd(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], sum);
            __syncwarp();
            unsigned int last = 0;
            if (tid == 0) {
                __threadfence();
                last = (atomicAdd(&p.ticket[rep], 1u) == (unsigned)p.ctas_per_replica - 1u);
                if (last) __threadfence();
            }
            last = __shfl_sync(0xFFFFFFFFu, last, 0);
            if (last) {
                if (tid < NCX) p.counts[(size_t)rep * NCA + tid] = atomicExch(&p.acc[(size_t)rep * ROLL_ACC_STRIDE + tid], 0ull);
                if (tid == 0) p.ticket[rep] = 0u;
            }
        }
    }
}

#endif  // __CUDACC__

// Generic path (any d1 >= 5): one thread per site, scalar evaluation.
__global__ void __launch_bounds__(256) kmc_sweep_kernel_generic(const SweepParams p)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)p.n_replicas * p.n) return;
    const int rep = (int)(g / p.n);
    const int i = (int)(g - (long long)rep * p.n);
    const int a = i / p.d1, b = i - a * p.d1;
    const uint8_t *st = p.status + (size_t)rep * p.n;
    struct G { const uint8_t *p; __device__ int operator()(int k) const { return __ldg(p + k); } } S{st};
    const int d0 = p.d0, d1 = p.d1;
    const int s0 = st[i];
    uint8_t out[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) out[j] = 0xFF;
    if (s0 == 0) { out[0] = C_CREATE_DIV; out[2] = out[3] = C_CMONO_FROM_DOUBLE; }
    if (s0 == 1) { out[3] = C_CMONO_MAKE_EMPTY; out[4] = C_FMONO_MAKE_DOUBLE; out[6] = C_FLIP; }
    if (s0 == 2) { out[2] = C_CMONO_MAKE_EMPTY; out[5] = C_FMONO_MAKE_DOUBLE; out[6] = C_FLIP; }
    if (s0 == 3) {
        out[1] = C_FILL_DIV; out[4] = out[5] = C_FMONO_FROM_EMPTY;
        const int am = wrap_dec(a - 1, d0), ap = wrap_inc(a + 1, d0);
        const int bm = wrap_dec(b - 1, d1), bp = wrap_inc(b + 1, d1);
        RingMasks m = ring_masks(S(am * d1 + b), S(a * d1 + bm), S(ap * d1 + bm), S(ap * d1 + b),
                                 S(a * d1 + bp), S(am * d1 + bp));
        for (int k = 0; k < 6; k++) {
            const int ri = kring(k);
            if ((m.P >> ri) & 1) out[7 + k] = (uint8_t)(C_MOVE_DIV + ((m.N >> ri) & 1) + 2 * ((m.F >> ri) & 1));
        }
        const int ap2 = wrap_inc(a + 2, d0), bp2 = wrap_inc(b + 2, d1), bm2 = wrap_dec(b - 2, d1);
        const bool t20 = S(ap2 * d1 + b) == 3;
        if (t20 && S(a * d1 + bp2) == 3) out[13] = C_CREATE_TREF;
        if (t20 && S(ap2 * d1 + bm2) == 3) out[14] = C_CREATE_TREF;
    }
    if (s0 == 4) out[15] = C_DESTROY_TREF;
    if (s0 == 7) out[16] = C_DESTROY_TREF;
    uint32_t *R = p.rowcnt + ((size_t)rep * d0 + a) * RCG_STRIDE;
    for (int j = 0; j < NS; j++) {
        if (p.slots) p.slots[((size_t)rep * NS + j) * p.n + i] = out[j];
        if (out[j] != 0xFF) atomicAdd(R + out[j], 1u);
    }
}

// class totals per replica from the row counts: counts[R][NCA]
__global__ void kmc_reduce_counts_kernel(const uint32_t *rowcnt, int d0, int n_replicas,
                                         unsigned long long *counts)
{
    const int rep = blockIdx.x;
    const int c = threadIdx.x & 15, part = threadIdx.x >> 4, nparts = blockDim.x >> 4;
    __shared__ unsigned long long acc[16];
    if (threadIdx.x < 16) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long s = 0;
    for (int a = part; a < d0; a += nparts) s += rowcnt[((size_t)rep * d0 + a) * RCG_STRIDE + c];
    atomicAdd(&acc[c], s);
    __syncthreads();
    if (threadIdx.x < NCA) counts[(size_t)rep * NCA + threadIdx.x] = threadIdx.x < NC ? acc[threadIdx.x] : 0ull;
}

// ---- MoveMonovacancy-aware full enumeration: all 9 rules, 29 slot planes, 17 classes -----------------------
// One thread per site (any shape).  Row counts use a stride of RCG_STRIDE_MM words and must be zero on entry.
constexpr int RCG_STRIDE_MM = 32;
__global__ void __launch_bounds__(256) kmc_sweep_kernel_mm(const SweepParams p)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)p.n_replicas * p.n) return;
    const int rep = (int)(g / p.n);
    const int i = (int)(g - (long long)rep * p.n);
    const int a = i / p.d1, b = i - a * p.d1;
    const uint8_t *st = p.status + (size_t)rep * p.n;
    struct G { const uint8_t *p; __device__ int operator()(int k) const { return __ldg(p + k); } } S{st};
    const int d0 = p.d0, d1 = p.d1;
    const int s0 = st[i];
    uint8_t out[NSA];
#pragma unroll
    for (int j = 0; j < NSA; j++) out[j] = 0xFF;
    if (s0 == 0) { out[0] = C_CREATE_DIV; out[2] = out[3] = C_CMONO_FROM_DOUBLE; }
    if (s0 == 1) { out[3] = C_CMONO_MAKE_EMPTY; out[4] = C_FMONO_MAKE_DOUBLE; out[6] = C_FLIP; }
    if (s0 == 2) { out[2] = C_CMONO_MAKE_EMPTY; out[5] = C_FMONO_MAKE_DOUBLE; out[6] = C_FLIP; }
    if (s0 >= 1 && s0 <= 3) {
        const int am = wrap_dec(a - 1, d0), ap = wrap_inc(a + 1, d0);
        const int bm = wrap_dec(b - 1, d1), bp = wrap_inc(b + 1, d1);
        const int ring[6] = {S(am * d1 + b), S(a * d1 + bm), S(ap * d1 + bm), S(ap * d1 + b), S(a * d1 + bp), S(am * d1 + bp)};
        // MoveMonovacancy: layer L of this site hops to neighbour k if that neighbour's layer set lacks L
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int dst = ring[kring(k)];
#pragma unroll
            for (int layer = 1; layer <= 2; layer++)
                if ((s0 & layer) && dst <= 3 && !(dst & layer))
                    out[MM_SLOT0 + 2 * k + layer - 1] = (uint8_t)(C_MOVE_MONO + 2 * (s0 == 3) + (dst != 0));
        }
        if (s0 == 3) {
            out[1] = C_FILL_DIV; out[4] = out[5] = C_FMONO_FROM_EMPTY;
            RingMasks m = ring_masks(ring[0], ring[1], ring[2], ring[3], ring[4], ring[5]);
            for (int k = 0; k < 6; k++) {
                const int ri = kring(k);
                if ((m.P >> ri) & 1) out[7 + k] = (uint8_t)(C_MOVE_DIV + ((m.N >> ri) & 1) + 2 * ((m.F >> ri) & 1));
            }
            const int ap2 = wrap_inc(a + 2, d0), bp2 = wrap_inc(b + 2, d1), bm2 = wrap_dec(b - 2, d1);
            const bool t20 = S(ap2 * d1 + b) == 3;
            if (t20 && S(a * d1 + bp2) == 3) out[13] = C_CREATE_TREF;
            if (t20 && S(ap2 * d1 + bm2) == 3) out[14] = C_CREATE_TREF;
        }
    }
    if (s0 == 4) out[15] = C_DESTROY_TREF;
    if (s0 == 7) out[16] = C_DESTROY_TREF;
    uint32_t *R = p.rowcnt + ((size_t)rep * d0 + a) * RCG_STRIDE_MM;
    for (int j = 0; j < NSA; j++) {
        if (p.slots) p.slots[((size_t)rep * NSA + j) * p.n + i] = out[j];
        if (out[j] != 0xFF) atomicAdd(R + out[j], 1u);
    }
}
__global__ void kmc_reduce_counts_mm_kernel(const uint32_t *rowcnt, int d0, int n_replicas, unsigned long long *counts)
{
    const int rep = blockIdx.x;
    const int c = threadIdx.x & 31, part = threadIdx.x >> 5, nparts = blockDim.x >> 5;
    __shared__ unsigned long long acc[32];
    if (threadIdx.x < 32) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long s = 0;
    for (int a = part; a < d0; a += nparts) s += rowcnt[((size_t)rep * d0 + a) * RCG_STRIDE_MM + c];
    atomicAdd(&acc[c], s);
    __syncthreads();
    if (threadIdx.x < NCA) counts[(size_t)rep * NCA + threadIdx.x] = acc[threadIdx.x];
}

}  // namespace kmc
