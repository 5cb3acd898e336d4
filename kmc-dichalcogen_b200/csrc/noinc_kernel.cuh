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
    const uint32_t *rowcnt;              // [R][d0][16]
    const unsigned long long *counts;    // [R][13]
    double rate[NC];
    int order_pos[NC];
    unsigned long long seed;
    uint32_t replica0;
    int log_mode;
    void *events;                        // [R][nsteps]
    long long nsteps, t;                 // this launch handles local step index t
    unsigned long long *steps_done, *hist, *ties;
    uint32_t *flags;
};

__global__ void __launch_bounds__(32) kmc_noinc_select_apply_kernel(const NoincParams p)
{
    const int rep = blockIdx.x;
    const int lane = threadIdx.x;
    const int d0 = p.d0, d1 = p.d1;
    if (p.flags[rep] & FLAG_ZERO_RATE) return;          // this replica already stopped
    uint8_t *st = p.status + (size_t)rep * p.n;
    const uint32_t *rcg = p.rowcnt + (size_t)rep * d0 * RCG_STRIDE;

    const int myc = lane < NC ? lane : 0;
    const double rate = (lane < NC && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NC && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    const long long tot = lane < NC ? (long long)p.counts[(size_t)rep * NC + lane] : 0;
    const unsigned long long step = p.steps_done[rep];
    double u1, u2;
    kmc_draw(p.seed, p.replica0 + (uint32_t)rep, step, u1, u2);

    const double w = __dmul_rn((double)tot, rate);
    PickState pstate = pick_state_init(lane);
    const ClassPick pick = pick_class(w, pos, pstate, u1, lane);
    const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)p.t;
    if (p.log_mode == KMC_LOG_FULL && lane < NC)
        (reinterpret_cast<kmc_event_full *>(p.events) + idx)->counts[lane] = (uint32_t)tot;
    if (pick.zero) {
        if (lane == 0) {
            p.flags[rep] |= FLAG_ZERO_RATE;
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
    const int csel = pick.csel;
    const long long cnt = __shfl_sync(FULL, tot, csel);
    long long rr = (long long)__dmul_rn(u2, (double)cnt);
    if (rr > cnt - 1) rr = cnt - 1;

    // row search: 64-bit running rank until it fits a chunk
    int row = 0;
    uint32_t r = 0;
    for (int base = 0; base < d0; base += 32) {
        const int a = base + lane;
        const uint32_t v = a < d0 ? rcg[(size_t)a * RCG_STRIDE + csel] : 0u;
        const uint32_t chunk = __reduce_add_sync(FULL, v);
        if (rr >= (long long)chunk) { rr -= chunk; continue; }
        r = (uint32_t)rr;
        int L;
        warp_rank_search(v, r, L, lane);
        row = base + L;
        break;
    }
    // site and slot inside the row (row, slot, column order)
    int col, slot;
    find_in_row_bytes(st, d0, d1, row, csel, r, lane, col, slot);
    const int site = row * d1 + col;
    const int s0 = st[site];
    __syncwarp();
    if (lane == 0) {
        apply_move(st, d0, d1, row, col, slot, s0);
        if (p.log_mode == KMC_LOG_FULL) {
            kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
            rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
            rec->flags = pick.tie ? 1 : 0; rec->total_rate = pick.total; rec->pad = 0;
        } else if (p.log_mode == KMC_LOG_COMPACT) {
            kmc_event_compact e; e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
            e.flags = pick.tie ? 1 : 0; e.total_rate = pick.total;
            reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
        }
        p.steps_done[rep] = step + 1;
        p.hist[(size_t)rep * NC + csel] += 1;
        if (pick.tie) p.ties[rep] += 1;
    }
}

// KmcSim.perform_given_move (sim.py:145-157) for one explicit move; result[0] = 0 ok, 1 = the move
// does not exist on the current lattice.
__global__ void kmc_apply_given_kernel(uint8_t *st, int d0, int d1, int site, int slot, int *result)
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
    } else ok = (slot == 15 && s0 == 4) || (slot == 16 && s0 == 7);
    if (ok) apply_move(st, d0, d1, a, b, slot, s0);
    result[0] = ok ? 0 : 1;
}

}  // namespace kmc
