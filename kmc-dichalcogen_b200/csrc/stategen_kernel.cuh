// stategen_kernel.cuh -- initial lattices generated on the device: the reference's
// gen_random_state (demo1/state.py:329-378) with its random source replaced by a Philox stream, so that
// the reference itself, handed a random.Random subclass backed by the same stream, produces the same
// lattice (that is how tests/golden/state_gen_philox.npz was made).
//
// Stream of replica r (a separate domain from the event draws: counter word 3 is 1, not 0):
//     block(q) = Philox4x32-10(counter = (q lo, q hi, seed hi, 1), key = (seed lo, r)),  q = 0, 1, 2, ...
//     random()       = (((x0 << 32) | x1) >> 11) * 2^-53      one block per call
//     getrandbits(k) = x0 >> (32 - k), 1 <= k <= 32            one block per call
//     _randbelow(n)  = CPython's: k = bit_length(n); r = getrandbits(k) until r < n
// 'exact'  (state.py:360-378): rng.shuffle(nodes) = for i = n-1 .. 1: j = _randbelow(i + 1), swap; then runs of
//          the shuffled list become divacancies / monovacancies, each monovacancy drawing rng.choice([1, 2]).
// 'approx' (state.py:342-356): one rng.uniform(0, total) per node through weighted_choice (bisect_left over the
//          sorted cumulative weights), then, in node order, rng.choice([1, 2]) for every monovacancy.
// The walk over the nodes is sequential by construction (rejection sampling makes the number of draws data
// dependent), so one thread generates one replica; replicas run in parallel.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

struct GenStream {
    unsigned long long seed, q;
    uint32_t replica;
    __device__ __forceinline__ void block(uint32_t &x0, uint32_t &x1) {
        uint32_t c0 = (uint32_t)q, c1 = (uint32_t)(q >> 32), c2 = (uint32_t)(seed >> 32), c3 = 1u;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, replica);
        q++;
        x0 = c0; x1 = c1;
    }
    __device__ __forceinline__ double random() {
        uint32_t x0, x1;
        block(x0, x1);
        return (double)((((uint64_t)x0 << 32) | x1) >> 11) * 0x1.0p-53;
    }
    __device__ __forceinline__ uint32_t randbelow(uint32_t n) {
        const int k = 32 - __clz(n);
        uint32_t x0, x1, r;
        do { block(x0, x1); r = x0 >> (32 - k); } while (r >= n);
        return r;
    }
};

struct GenParams {
    int mode;                       // 0 exact, 1 approx
    int n, n_replicas, n_keys;
    int code[3];                    // per key: 3 divacancy, 1 monovacancy (layer drawn), 0 nothing
    long long bound[4];             // exact: interval end points in the shuffled list
    double cumul[3];                // approx: cumulative weights of the sorted live pairs
    unsigned long long seed;
    uint32_t replica0;
    uint8_t *status;                // [R][n], zeroed by the caller
    uint32_t *perm;                 // exact: [n][R] scratch
};

// 'approx', first half (state.py:342-356 through kmc.py:41-50): site i takes draw i of its replica's stream,
// x = 0.0 + (total - 0.0) * random(), category = bisect_left(cumulative weights, x).  One thread per site.
__global__ void kmc_generate_approx_sites_kernel(const GenParams p)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.n_replicas * p.n) return;
    const int r = (int)(gid / p.n);
    const int i = (int)(gid - (long long)r * p.n);
    GenStream g{p.seed, (unsigned long long)i, p.replica0 + (uint32_t)r};
    const double x = __dmul_rn(p.cumul[p.n_keys - 1], g.random());
    int k = 0;
    while (k < p.n_keys - 1 && p.cumul[k] < x) k++;
    p.status[gid] = p.code[k] == 1 ? 0xFF : (uint8_t)p.code[k];
}

__global__ void kmc_generate_state_kernel(const GenParams p)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_replicas) return;
    const int n = p.n, R = p.n_replicas;
    uint8_t *st = p.status + (size_t)r * n;
    GenStream g{p.seed, 0ull, p.replica0 + (uint32_t)r};
    if (p.mode == 0) {
        uint32_t *perm = p.perm + r;
        for (int i = 0; i < n; i++) perm[(size_t)i * R] = (uint32_t)i;
        for (int i = n - 1; i >= 1; i--) {
            const uint32_t j = g.randbelow((uint32_t)i + 1u);
            const uint32_t a = perm[(size_t)i * R], b = perm[(size_t)j * R];
            perm[(size_t)i * R] = b; perm[(size_t)j * R] = a;
        }
        for (int k = 0; k < p.n_keys; k++)
            for (long long q = p.bound[k]; q < p.bound[k + 1]; q++) {
                const uint32_t node = perm[(size_t)q * R];
                if (p.code[k] == 3) st[node] = 3;
                else if (p.code[k] == 1) st[node] = (uint8_t)(1 + g.randbelow(2u));
            }
    } else {
        // 'approx', second half: the per-site categories were drawn by kmc_generate_approx_sites_kernel (draw i
        // belongs to site i, so that part is parallel); the layer of every monovacancy comes next in the stream,
        // in node order, with rejection sampling -- sequential
        g.q = (unsigned long long)n;
        int i = 0;
        if ((n & 15) == 0) {                    // skip 16 sites at a time where no monovacancy is marked
            const uint4 *s4 = reinterpret_cast<const uint4 *>(st);
            for (; i < n; i += 16) {
                const uint4 v = s4[i >> 4];
                const uint32_t any = (v.x & (v.x >> 1)) | (v.y & (v.y >> 1)) | (v.z & (v.z >> 1)) | (v.w & (v.w >> 1));
                if (!(any & 0x40404040u)) continue;             // codes here are 0, 3 or 0xFF: bit 6 set only in 0xFF
                for (int j = i; j < i + 16; j++)
                    if (st[j] == 0xFF) st[j] = (uint8_t)(1 + g.randbelow(2u));
            }
        }
        for (; i < n; i++)
            if (st[i] == 0xFF) st[i] = (uint8_t)(1 + g.randbelow(2u));
    }
}

}  // namespace kmc
