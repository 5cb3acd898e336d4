// noinc_kernel.cuh -- K3 for lattices held in HBM: class choice + hierarchical rank search + apply,
// run after a full sweep (K1/K2).  Together they are one `--no-incremental` event of the reference
// (demo1/sim.py:114-143): re-enumerate everything, choose, perform.
//
// One warp per replica.  Inputs are the sweep's per-row class counts and class totals; the site
// search inside the chosen row re-derives class counts from the status plane (no slot planes needed).
#pragma once
#include "replica_kernel.cuh"
#include "sweep_kernel.cuh"

namespace kmc {

struct NoincParams {
    int d0, d1, n, n_replicas;
    uint8_t *status;                     // [R][n]
    const uint32_t *rowcnt;              // [R][d0][rc_stride]
    const unsigned long long *counts;    // [R][NCA]
    int rc_stride;                       // 16 (aligned sweep kernels) or 32 (MoveMonovacancy-aware sweep)
    const uint32_t *rowmm;               // [R][d0][4] MoveMonovacancy row counts of the aligned two-kernel sweep, or nullptr
    int mm;                              // MoveMonovacancy classes are live: 17 classes
    double rate[NCA];
    int order_pos[NCA];
    unsigned long long seed;
    uint32_t replica0;
    int log_mode;
    void *events;                        // [R][nsteps]
    long long nsteps, t;                 // this launch handles local step index t
    unsigned long long *steps_done, *hist, *ties, *hops;
    uint32_t *flags, *any_flag;
    const double *draws;                 // [R][nsteps][2] injected draws or nullptr (Philox)
    double *clock;                       // [R] or nullptr
    int clock_mode;
    uint32_t *tracer;                    // [R][n] or nullptr (see ReplicaParams)
    int stage;                           // 1: stage rows row-1 .. row+2 of the chosen row in shared memory (4 * d1 bytes)
};

// exclusive prefix of `v` over the 256 threads of the block (wsum: 8 words of shared scratch)
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *wsum,
                                                                  int tid, unsigned long long &total)
{
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                       // wsum may still be read from a previous call
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const unsigned long long x = wsum[w]; if (w < warp) base += x; tot += x; }
    total = tot;
    return base + inc - v;
}

// One CTA (256 threads) per replica: warp 0 chooses the class; all threads share the rank search over
// the rows (each owns a contiguous block of rows) and over the sites of the chosen row.
__global__ void __launch_bounds__(256) kmc_noinc_select_apply_kernel(const NoincParams p)
{
    __shared__ unsigned long long wsum[8];
    __shared__ unsigned long long s_rank;
    __shared__ double s_total;
    __shared__ int s_csel, s_zero, s_tie, s_row, s_col;
    __shared__ uint32_t s_rin;
    __shared__ uint32_t s_slotcnt[12];
    const int rep = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int d0 = p.d0, d1 = p.d1;
    const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)p.t;
    if (p.flags[rep] & FLAG_ZERO_RATE) {                // this replica already stopped (uniform): no event
        if (threadIdx.x == 0) {
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = 0xFFFFFFFFu; rec->cls = KMC_CLASS_NONE; rec->slot = KMC_SLOT_NONE;
                rec->flags = 0; rec->total_rate = 0.0; rec->pad = 0;
            } else if (p.log_mode == KMC_LOG_COMPACT) {
                kmc_event_compact e; e.site = 0xFFFFFFFFu; e.cls = KMC_CLASS_NONE; e.slot = KMC_SLOT_NONE;
                e.flags = 0; e.total_rate = 0.0;
                reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
            }
        }
        return;
    }
    uint8_t *st = p.status + (size_t)rep * p.n;
    const int rcs = p.rc_stride;
    const uint32_t *rcg = p.rowcnt + (size_t)rep * d0 * rcs;
    const unsigned long long step = p.steps_done[rep];

    if (tid < 32) {
        const int ncx = p.mm ? NCA : NC;
        const int myc = lane < ncx ? lane : 0;
        const double rate = (lane < ncx && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
        const int pos = (lane < ncx && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
        const long long tot = lane < ncx ? (long long)p.counts[(size_t)rep * NCA + lane] : 0;
        double u1, u2;
        if (p.draws) {
            const double2 u = *reinterpret_cast<const double2 *>(p.draws + ((size_t)rep * (size_t)p.nsteps + (size_t)p.t) * 2);
            u1 = u.x; u2 = u.y;
        } else {
            kmc_draw(p.seed, p.replica0 + (uint32_t)rep, step, u1, u2);
        }
        const double w = __dmul_rn((double)tot, rate);
        ClassPick pick;
        if (p.mm) { PickState pstate = pick_state_init_n<32>(lane); pick = pick_class<32>(w, pos, pstate, u1, lane); }
        else { PickState pstate = pick_state_init_n<16>(lane); pick = pick_class<16>(w, pos, pstate, u1, lane); }
        if (p.log_mode == KMC_LOG_FULL && lane < NCA)
            (reinterpret_cast<kmc_event_full *>(p.events) + idx)->counts[lane] = (uint32_t)tot;
        const long long cnt = __shfl_sync(FULL, tot, pick.zero ? 0 : pick.csel);
        long long rr = (long long)__dmul_rn(u2, (double)cnt);
        if (rr > cnt - 1) rr = cnt - 1;
        if (lane == 0) {
            s_csel = pick.csel; s_zero = pick.zero; s_tie = pick.tie != 0; s_total = pick.total;
            s_rank = (unsigned long long)(rr < 0 ? 0 : rr);
        }
    }
    if (tid < 12) s_slotcnt[tid] = 0;
    __syncthreads();
    if (s_zero) {
        if (tid == 0) {
            p.flags[rep] |= FLAG_ZERO_RATE;
            atomicOr(p.any_flag, FLAG_ZERO_RATE);
            if (p.log_mode == KMC_LOG_FULL) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                rec->site = 0xFFFFFFFFu; rec->cls = KMC_CLASS_NONE; rec->slot = KMC_SLOT_NONE;
                rec->flags = 0; rec->total_rate = 0.0; rec->pad = 0;
            } else if (p.log_mode == KMC_LOG_COMPACT) {
                kmc_event_compact e; e.site = 0xFFFFFFFFu; e.cls = KMC_CLASS_NONE; e.slot = KMC_SLOT_NONE;
                e.flags = 0; e.total_rate = 0.0;
                reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
            }
        }
        return;
    }
    const int csel = s_csel;
    const unsigned long long rank = s_rank;

    // ---- row: each thread owns rows [tid*kr, tid*kr + kr) --------------------------------------------
    {
        const int kr = (d0 + 255) / 256;
        const int r0 = tid * kr, r1 = min(r0 + kr, d0);
        unsigned long long local = 0, total;
        const uint32_t *rmm = p.rowmm ? p.rowmm + (size_t)rep * d0 * 4 : nullptr;
        const bool in_mm = rmm && csel >= C_MOVE_MONO;
        for (int a = r0; a < r1; a++) local += in_mm ? rmm[(size_t)a * 4 + csel - C_MOVE_MONO] : rcg[(size_t)a * rcs + csel];
        const unsigned long long excl = block_exclusive_scan(local, wsum, tid, total);
        if (rank >= excl && rank < excl + local) {
            unsigned long long r = rank - excl;
            int a = r0;
            for (; a < r1; a++) {
                const unsigned long long v = in_mm ? rmm[(size_t)a * 4 + csel - C_MOVE_MONO] : rcg[(size_t)a * rcs + csel];
                if (r < v) break;
                r -= v;
            }
            s_row = a; s_rin = (uint32_t)r;
        }
    }
    __syncthreads();
    const int row = s_row;
    // The two passes below evaluate every site of the chosen row with its neighbours.  Straight from the lattice
    // that is a chain of dependent L2 round trips per site; staging the four rows they can touch (row-1 .. row+2)
    // in shared memory turns it into one coalesced round of loads.  The staged rows form a 4-row lattice in which
    // the chosen row is row 1 and none of the offsets used (-1 .. +2) wraps.
    extern __shared__ uint4 s_stage[];
    const uint8_t *view = st;
    int vrow = row, vd0 = d0;
    if (p.stage) {
        uint8_t *dst = reinterpret_cast<uint8_t *>(s_stage);
        for (int j = 0; j < 4; j++) {
            int a = row - 1 + j;
            a = a < 0 ? a + d0 : (a >= d0 ? a - d0 : a);
            const uint8_t *src = st + (size_t)a * d1;
            if ((d1 & 15) == 0) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
                uint4 *d4 = reinterpret_cast<uint4 *>(dst + (size_t)j * d1);
                for (int i = tid; i < (d1 >> 4); i += 256) d4[i] = s4[i];
            } else {
                for (int i = tid; i < d1; i += 256) dst[(size_t)j * d1 + i] = src[i];
            }
        }
        __syncthreads();
        view = dst; vrow = 1; vd0 = 4;
    }
    // ---- slot: per-slot totals of the class in this row (row, slot, column order) ------------------------
    {
        uint32_t acc[6] = {0u, 0u, 0u, 0u, 0u, 0u};               // twelve 16-bit per-slot counters
        for (int b = tid; b < d1; b += 256) {
            const uint32_t m = site_slot_mask(view, vrow, b, vd0, d1, csel);
#pragma unroll
            for (int q = 0; q < 6; q++) acc[q] += ((m >> (2 * q)) & 1u) | (((m >> (2 * q + 1)) & 1u) << 16);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const uint32_t a = __reduce_add_sync(FULL, acc[q]);
            if (lane == 0) {
                if (a & 0xFFFFu) atomicAdd(&s_slotcnt[2 * q], a & 0xFFFFu);
                if (a >> 16) atomicAdd(&s_slotcnt[2 * q + 1], a >> 16);
            }
        }
    }
    __syncthreads();
    uint32_t r = s_rin;
    int j = 0;
#pragma unroll
    for (int q = 0; q < 11; q++)
        if (j == q && r >= s_slotcnt[q]) { r -= s_slotcnt[q]; j = q + 1; }
    // ---- site: each thread owns columns [tid*kc, tid*kc + kc) ---------------------------------------------
    {
        const int kc = (d1 + 255) / 256;
        const int b0 = tid * kc, b1 = min(b0 + kc, d1);
        unsigned long long local = 0, total;
        for (int b = b0; b < b1; b++) local += (site_slot_mask(view, vrow, b, vd0, d1, csel) >> j) & 1u;
        const unsigned long long excl = block_exclusive_scan(local, wsum, tid, total);
        if ((unsigned long long)r >= excl && (unsigned long long)r < excl + local) {
            uint32_t q = r - (uint32_t)excl;
            int b = b0;
            for (; b < b1; b++) {
                const uint32_t v = (site_slot_mask(view, vrow, b, vd0, d1, csel) >> j) & 1u;
                if (v && q == 0) break;
                q -= v;
            }
            s_col = b;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int col = s_col;
        const int slot = class_slot_base(csel) + j;
        const int site = row * d1 + col;
        const int s0 = st[site];
        const MoveNodes m = move_nodes_st(st, row, col, slot, s0, d0, d1);      // before the lattice changes
        apply_move(st, d0, d1, row, col, slot, s0);
        if (p.tracer) {
            tracer_update(p.tracer + (size_t)rep * p.n, slot, m.na, site, m.ns0, m.a1 * d1 + m.b1, m.ns1,
                          m.a2 * d1 + m.b2, m.ns2, s0);
        }
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
            rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
            rec->flags = s_tie ? 1 : 0; rec->total_rate = s_total; rec->pad = 0;
        } else if (p.log_mode == KMC_LOG_COMPACT) {
            kmc_event_compact e; e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
            e.flags = s_tie ? 1 : 0; e.total_rate = s_total;
            reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
        }
        p.steps_done[rep] = step + 1;
        if (p.clock) p.clock[rep] = __dadd_rn(p.clock[rep], kmc_time_step(p.clock_mode, s_total, p.seed,
                                                                          p.replica0 + (uint32_t)rep, step));
        p.hist[(size_t)rep * NCA + csel] += 1;
        if (slot >= 7 && slot < 13) p.hops[(size_t)rep * 6 + slot - 7] += 1;
        if (s_tie) p.ties[rep] += 1;
    }
}

// KmcSim.perform_given_move (sim.py:145-157) for one explicit move; result[0] = 0 ok, 1 = the move
// does not exist on the current lattice.
__global__ void kmc_apply_given_kernel(uint8_t *st, int d0, int d1, int site, int slot, int *result, uint32_t *tracer)
{
    const int a = site / d1, b = site - a * d1;
    struct G { const uint8_t *p; __device__ int operator()(int k) const { return p[k]; } } S{st};
    const int s0 = st[site];
    bool ok;
    if (slot == 0) ok = s0 == 0;
    else if (slot == 1) ok = s0 == 3;
    else if (slot <= 3) ok = s0 <= 3 && ((s0 | (slot - 1)) != s0);
    else if (slot <= 5) ok = s0 <= 3 && ((s0 & ~(slot - 3)) != s0);
    else if (slot == 6) ok = s0 == 1 || s0 == 2;
    else if (slot <= 12) {
        const int ri = kring(slot - 7);
        const int na = wrap_inc(wrap_dec(a + (int)((RING_DA >> (4 * ri)) & 0xF) - 2, d0), d0);
        const int nb = wrap_inc(wrap_dec(b + (int)((RING_DB >> (4 * ri)) & 0xF) - 2, d1), d1);
        ok = s0 == 3 && st[na * d1 + nb] == 0;
    } else if (slot <= 14) {
        F4 f = eval_site(S, a, b, d0, d1);
        const int ap2 = wrap_inc(a + 2, d0), bp2 = wrap_inc(b + 2, d1), bm2 = wrap_dec(b - 2, d1);
        const bool t20 = st[ap2 * d1 + b] == 3;
        ok = s0 == 3 && t20 && (slot == 13 ? st[a * d1 + bp2] == 3 : st[ap2 * d1 + bm2] == 3);
        (void)f;
    } else if (slot <= 16) ok = (slot == 15 && s0 == 4) || (slot == 16 && s0 == 7);
    else {                                                 // MoveMonovacancy (direction, layer)
        const int ri = kring((slot - MM_SLOT0) >> 1), layer = ((slot - MM_SLOT0) & 1) + 1;
        const int na = wrap_inc(wrap_dec(a + (int)((RING_DA >> (4 * ri)) & 0xF) - 2, d0), d0);
        const int nb = wrap_inc(wrap_dec(b + (int)((RING_DB >> (4 * ri)) & 0xF) - 2, d1), d1);
        const int dst = st[na * d1 + nb];
        ok = s0 >= 1 && s0 <= 3 && (s0 & layer) && dst <= 3 && !(dst & layer);
    }
    if (ok) {
        const MoveNodes m = move_nodes_st(st, a, b, slot, s0, d0, d1);
        apply_move(st, d0, d1, a, b, slot, s0);
        if (tracer) tracer_update(tracer, slot, m.na, site, m.ns0, m.a1 * d1 + m.b1, m.ns1, m.a2 * d1 + m.b2, m.ns2, s0);
    }
    result[0] = ok ? 0 : 1;
}

}  // namespace kmc
