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

extern "C" void kmc_hostcheck_draw(uint64_t seed, uint32_t replica, uint64_t step, double *u1, double *u2)
{
    kmc_draw(seed, replica, step, *u1, *u2);
}

extern "C" int kmc_hostcheck_rc_field(int c) { return rc_field(c); }
