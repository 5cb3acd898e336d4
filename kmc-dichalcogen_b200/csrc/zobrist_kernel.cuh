// zobrist_kernel.cuh -- Zobrist key of every replica's lattice, from scratch.
// The reference XORs one random value per entity (Vacancy(node, layers), Trefoil(nodes)) into an
// incrementally maintained key (demo1/incremental.py:335-380, state.py:199-248) and re-derives it from
// scratch when validating (state.py:136-141).  Its values are drawn on demand from the global `random`
// module; ours come from a fixed Philox table so that keys are reproducible and comparable across runs:
//     value(site, code) = top `bits` bits of (x0 << 32 | x1),
//     (x0, x1, ..) = Philox4x32-10(counter = (site, code, seed >> 32, 2), key = (seed lo, 0))
// code = status byte of the entity's site: 1, 2, 3 (vacancy with that layer set), 4 / 7 (Up / Down trefoil,
// counted once at its anchor).  This kernel is the from-scratch side; the host tracker (demo1/zobrist.py)
// toggles the same values per logged event.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

KMC_HD unsigned long long zobrist_value(uint32_t site, uint32_t code, int bits, unsigned long long seed) {
    uint32_t c0 = site, c1 = code, c2 = (uint32_t)(seed >> 32), c3 = 2u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, 0u);
    return ((((unsigned long long)c0) << 32) | c1) >> (64 - bits);
}

// grid: (n_replicas, blocks per replica) -- replicas on grid.x, which has no 65535 limit; out[rep] must be zero on entry
__global__ void kmc_zobrist_kernel(const uint8_t *status, int n, int bits, unsigned long long seed,
                                   unsigned long long *out)
{
    const int rep = blockIdx.x;
    const uint8_t *st = status + (size_t)rep * n;
    unsigned long long acc = 0;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const uint32_t s = st[i];
        if ((s >= 1 && s <= 4) || s == 7) acc ^= zobrist_value((uint32_t)i, s, bits, seed);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicXor(out + rep, acc);
}

}  // namespace kmc
