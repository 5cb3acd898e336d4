// hostcheck.cuh -- host-callable wrappers around the __host__ __device__ pieces of the kernels, so the
// CPU test-suite (no GPU) can check the byte-parallel sweep arithmetic and the per-site evaluator
// against the oracle.  Included by capi.cu (one translation unit); performs no CUDA calls.
#pragma once
#include <stdint.h>
#include <string.h>
#include "kmc_common.cuh"
#include "sweep_kernel.cuh"

using namespace kmc;

// Emulates kmc_sweep_kernel's word gathering on the host: slots [17][n], counts [13].
extern "C" int kmc_hostcheck_sweep_words(const uint8_t *status, int d0, int d1, uint8_t *slots, int64_t *counts)
{
    if (d1 & 3) return -1;
    const int n = d0 * d1, wpr = d1 >> 2;
    const uint32_t *S = reinterpret_cast<const uint32_t *>(status);
    for (int c = 0; c < NC; c++) counts[c] = 0;
    for (int a = 0; a < d0; a++)
        for (int wi = 0; wi < wpr; wi++) {
            const int am = a == 0 ? d0 - 1 : a - 1, ap = a + 1 == d0 ? 0 : a + 1, ap2 = ap + 1 == d0 ? 0 : ap + 1;
            const int wm = wi == 0 ? wpr - 1 : wi - 1, wn = wi + 1 == wpr ? 0 : wi + 1;
            WordIn in;
            in.m_cur = S[am * wpr + wi]; in.m_next = S[am * wpr + wn];
            in.c_prev = S[a * wpr + wm]; in.c_cur = S[a * wpr + wi]; in.c_next = S[a * wpr + wn];
            in.p_prev = S[ap * wpr + wm]; in.p_cur = S[ap * wpr + wi];
            in.q_prev = S[ap2 * wpr + wm]; in.q_cur = S[ap2 * wpr + wi];
            uint32_t out[NS], cnt[NC];
            sweep_word(in, out, cnt);
            for (int j = 0; j < NS; j++) memcpy(slots + (size_t)j * n + (a * wpr + wi) * 4, &out[j], 4);
            for (int c = 0; c < NC; c++) counts[c] += cnt[c];
        }
    return 0;
}

// Emulates kmc_sweep16_kernel's chunk gathering (16 sites per thread) on the host.
extern "C" int kmc_hostcheck_sweep_words16(const uint8_t *status, int d0, int d1, uint8_t *slots, int64_t *counts)
{
    if (d1 & 15) return -1;
    const int n = d0 * d1, cpr = d1 >> 4, wpr = d1 >> 2;
    const uint32_t *S = reinterpret_cast<const uint32_t *>(status);
    for (int c = 0; c < NC; c++) counts[c] = 0;
    for (int a = 0; a < d0; a++)
        for (int ci = 0; ci < cpr; ci++) {
            const int am = a == 0 ? d0 - 1 : a - 1, ap = a + 1 == d0 ? 0 : a + 1, ap2 = ap + 1 == d0 ? 0 : ap + 1;
            const int cim = ci == 0 ? cpr - 1 : ci - 1, cin = ci + 1 == cpr ? 0 : ci + 1;
            const uint32_t *r0 = S + a * wpr, *rm = S + am * wpr, *rp = S + ap * wpr, *rq = S + ap2 * wpr;
            Chunk16 k;
            k.c[0] = zd_flags(r0[4 * cim + 3]); k.c[5] = zd_flags(r0[4 * cin]);
            k.p[0] = zd_flags(rp[4 * cim + 3]); k.q[0] = zd_flags(rq[4 * cim + 3]).div;
            k.m[4] = zd_flags(rm[4 * cin]);
            for (int j = 0; j < 4; j++) {
                k.cw[j] = r0[4 * ci + j];
                k.c[1 + j] = zd_flags(k.cw[j]); k.m[j] = zd_flags(rm[4 * ci + j]);
                k.p[1 + j] = zd_flags(rp[4 * ci + j]); k.q[1 + j] = zd_flags(rq[4 * ci + j]).div;
            }
            ByteAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < 4; j++) {
                uint32_t out[NS], o[9];
                own_planes(k, j, o, acc);
                for (int q = 0; q < 7; q++) out[q] = o[q];
                out[15] = o[7]; out[16] = o[8];
                out[7] = move_plane<0>(k, j, acc); out[8] = move_plane<1>(k, j, acc); out[9] = move_plane<2>(k, j, acc);
                out[10] = move_plane<3>(k, j, acc); out[11] = move_plane<4>(k, j, acc); out[12] = move_plane<5>(k, j, acc);
                trefoil_planes(k, j, out[13], out[14], acc);
                for (int q = 0; q < NS; q++) memcpy(slots + (size_t)q * n + (size_t)a * d1 + 16 * ci + 4 * j, &out[q], 4);
            }
            const uint32_t z = hsum(acc.zero), dv = hsum(acc.div), mo = hsum(acc.mono);
            counts[0] += z; counts[1] += dv; counts[2] += hsum(acc.nat); counts[3] += hsum(acc.met);
            counts[4] += hsum(acc.cmb); counts[5] += hsum(acc.bth); counts[6] += hsum(acc.tref);
            counts[7] += hsum(acc.anch); counts[8] += 2 * z; counts[9] += mo; counts[10] += mo;
            counts[11] += 2 * dv; counts[12] += mo;
        }
    return 0;
}

// Emulates kmc_sweep16n_kernel (nibble-parallel arithmetic + PRMT tables) on the host.
static inline void hc_lut(uint8_t *slots, size_t n, size_t base, int plane, uint32_t lo, uint32_t hi,
                          uint32_t s0, uint32_t s1)
{
    const uint32_t w[4] = {prmt(lo, hi, s0), prmt(lo, hi, s0 >> 16), prmt(lo, hi, s1), prmt(lo, hi, s1 >> 16)};
    memcpy(slots + (size_t)plane * n + base, w, 16);
}
template <int K> static inline void hc_move(const Chunk16N &k, uint8_t *slots, size_t n, size_t base, NibAcc &acc)
{
    uint32_t w[4];
    move_plane_n<K>(k, 0, w[0], w[1], acc);
    move_plane_n<K>(k, 1, w[2], w[3], acc);
    memcpy(slots + (size_t)(7 + K) * n + base, w, 16);
}
extern "C" int kmc_hostcheck_sweep_nibbles(const uint8_t *status, int d0, int d1, uint8_t *slots, int64_t *counts)
{
    if (d1 & 15) return -1;
    const size_t n = (size_t)d0 * d1;
    const int cpr = d1 >> 4, wpr = d1 >> 2;
    const uint32_t *S = reinterpret_cast<const uint32_t *>(status);
    for (int c = 0; c < NC; c++) counts[c] = 0;
    for (int a = 0; a < d0; a++)
        for (int ci = 0; ci < cpr; ci++) {
            const int am = a == 0 ? d0 - 1 : a - 1, ap = a + 1 == d0 ? 0 : a + 1, ap2 = ap + 1 == d0 ? 0 : ap + 1;
            const int cim = ci == 0 ? cpr - 1 : ci - 1, cin = ci + 1 == cpr ? 0 : ci + 1;
            const uint32_t *r0 = S + a * wpr + 4 * ci, *rm = S + am * wpr + 4 * ci, *rp = S + ap * wpr + 4 * ci;
            const uint32_t *rq = S + ap2 * wpr + 4 * ci;
            Chunk16N k;
            k.cn[0] = pack_nib(r0[0], r0[1]); k.cn[1] = pack_nib(r0[2], r0[3]);
            k.c[0] = nzd(pack_nib(0u, S[a * wpr + 4 * cim + 3])); k.c[1] = nzd(k.cn[0]); k.c[2] = nzd(k.cn[1]);
            k.c[3] = nzd(pack_nib(S[a * wpr + 4 * cin], 0u));
            k.m[0] = nzd(pack_nib(rm[0], rm[1])); k.m[1] = nzd(pack_nib(rm[2], rm[3]));
            k.m[2] = nzd(pack_nib(S[am * wpr + 4 * cin], 0u));
            k.p[0] = nzd(pack_nib(0u, S[ap * wpr + 4 * cim + 3])); k.p[1] = nzd(pack_nib(rp[0], rp[1]));
            k.p[2] = nzd(pack_nib(rp[2], rp[3]));
            k.q[0] = nzd(pack_nib(0u, S[ap2 * wpr + 4 * cim + 3])).div; k.q[1] = nzd(pack_nib(rq[0], rq[1])).div;
            k.q[2] = nzd(pack_nib(rq[2], rq[3])).div;
            const size_t base = (size_t)a * d1 + 16 * ci;
            const uint32_t s0 = k.cn[0] - 3u * ((k.cn[0] >> 3) & ONESN), s1 = k.cn[1] - 3u * ((k.cn[1] >> 3) & ONESN);
            const uint32_t F = 0xFFFFFFFFu;
            hc_lut(slots, n, base, 0, 0xFFFFFF00u, F, s0, s1); hc_lut(slots, n, base, 1, 0x01FFFFFFu, F, s0, s1);
            hc_lut(slots, n, base, 2, 0xFF09FF08u, F, s0, s1); hc_lut(slots, n, base, 3, 0xFFFF0908u, F, s0, s1);
            hc_lut(slots, n, base, 4, 0x0BFF0AFFu, F, s0, s1); hc_lut(slots, n, base, 5, 0x0B0AFFFFu, F, s0, s1);
            hc_lut(slots, n, base, 6, 0xFF0C0CFFu, F, s0, s1);
            hc_lut(slots, n, base, 15, F, 0xFFFFFF07u, s0, s1); hc_lut(slots, n, base, 16, F, 0x07FFFFFFu, s0, s1);
            uint32_t mono = 0, anch = 0;
            for (int j = 0; j < 2; j++) {
                const uint32_t x = k.cn[j];
                const uint32_t e0 = x & ONESN, e1 = (x >> 1) & ONESN, e2 = (x >> 2) & ONESN, e3 = (x >> 3) & ONESN;
                mono += __builtin_popcount((e0 ^ e1) & ~(e2 | e3));
                anch += __builtin_popcount((e2 & ~(e0 | e1 | e3)) | (e0 & e1 & e2 & ~e3));
            }
            const uint32_t nz = __builtin_popcount(k.c[1].zero) + __builtin_popcount(k.c[2].zero);
            const uint32_t nd = __builtin_popcount(k.c[1].div) + __builtin_popcount(k.c[2].div);
            NibAcc acc = {0, 0, 0, 0};
            hc_move<0>(k, slots, n, base, acc); hc_move<1>(k, slots, n, base, acc); hc_move<2>(k, slots, n, base, acc);
            hc_move<3>(k, slots, n, base, acc); hc_move<4>(k, slots, n, base, acc); hc_move<5>(k, slots, n, base, acc);
            uint32_t u[2], d[2];
            for (int j = 0; j < 2; j++) {
                const uint32_t both = k.c[j + 1].div & k.q[j + 1];
                u[j] = both & nshr(k.c[j + 1].div, k.c[j + 2].div, 2);
                d[j] = both & nshl(k.q[j], k.q[j + 1], 2);
            }
            hc_lut(slots, n, base, 13, 0xFFFF06FFu, F, u[0], u[1]); hc_lut(slots, n, base, 14, 0xFFFF06FFu, F, d[0], d[1]);
            const uint32_t P = nib_sum(acc.pres), PN = nib_sum(acc.pn), PF = nib_sum(acc.pf), PNF = nib_sum(acc.pnf);
            counts[0] += nz; counts[1] += nd; counts[2] += P - PN - PF + PNF; counts[3] += PN - PNF;
            counts[4] += PF - PNF; counts[5] += PNF;
            counts[6] += __builtin_popcount(u[0]) + __builtin_popcount(u[1]) + __builtin_popcount(d[0]) + __builtin_popcount(d[1]);
            counts[7] += anch; counts[8] += 2 * nz; counts[9] += mono; counts[10] += mono; counts[11] += 2 * nd;
            counts[12] += mono;
        }
    return 0;
}

// eval_site over the whole lattice: out [n][13] class counts per site
extern "C" int kmc_hostcheck_eval_sites(const uint8_t *status, int d0, int d1, uint8_t *out)
{
    struct H { const uint8_t *p; __host__ __device__ int operator()(int i) const { return p[i]; } } S{status};
    for (int a = 0; a < d0; a++)
        for (int b = 0; b < d1; b++) {
            F4 f = eval_site(S, a, b, d0, d1);
            for (int c = 0; c < NC; c++) out[(size_t)(a * d1 + b) * NC + c] = (uint8_t)f4_class(f, c);
        }
    return 0;
}

// the MoveMonovacancy-aware evaluator: out [n][17]
extern "C" int kmc_hostcheck_eval_sites_mm(const uint8_t *status, int d0, int d1, uint8_t *out)
{
    struct H { const uint8_t *p; __host__ __device__ int operator()(int i) const { return p[i]; } } S{status};
    for (int a = 0; a < d0; a++)
        for (int b = 0; b < d1; b++) {
            F4 f = eval_site<true>(S, a, b, d0, d1);
            for (int c = 0; c < NCA; c++) out[(size_t)(a * d1 + b) * NCA + c] = (uint8_t)f4_class(f, c);
        }
    return 0;
}

extern "C" void kmc_hostcheck_draw(uint64_t seed, uint32_t replica, uint64_t step, double *u1, double *u2)
{
    kmc_draw(seed, replica, step, *u1, *u2);
}

extern "C" int kmc_hostcheck_rc_field(int c) { return rc_field(c); }

// the stochastic clock's pieces as the host compiler builds them (same source as the device code)
extern "C" double kmc_hostcheck_log(double x) { return kmc_log(x); }
extern "C" double kmc_hostcheck_draw3(uint64_t seed, uint32_t replica, uint64_t step) { return kmc_draw3(seed, replica, step); }
extern "C" double kmc_hostcheck_time_step(int mode, double total, uint64_t seed, uint32_t replica, uint64_t step)
{
    return kmc_time_step(mode, total, seed, replica, step);
}
