// replica_wide2_kernel.cuh -- K3 + K4 on HBM-resident bit planes with a whole CTA (8 warps) per replica: the
// latency build of replica_wide_kernel.cuh for SMALL batches (BASELINE config 5: 1024 x 1024, 128 replicas per GPU,
// i.e. fewer replicas than SMs).
//
// With one warp per replica (replica_wide_kernel.cuh) an event is a single dependent chain of ~1600 warp-instructions
// at ~5 cycles each: the SM is idle 95 % of the time and a batch of 128 replicas cannot hide it.  The chain is long
// because one warp does everything in turn and because its refresh tasks are data-driven (a lane may hold any of the
// six MoveDivacancy directions or a trefoil task, so every operand is decoded from a descriptor at run time).
// Here the work of one event is split over eight warps by ROLE, so that every warp runs straight-line code for its
// role and the roles run side by side:
//   warp k = 0..5   MoveDivacancy direction k: its masks in the site search, its (row, word) refresh tasks
//   warp 6          trefoil shapes: CreateTrefoil site search and refresh tasks
//   warp 7          bookkeeping: event record, statistics, tracer, own-status row counts, the plane rewrite
//   warp 0 also     class choice + row search (it owns the 13 class totals, lane c = class c)
// Six block barriers per event order the phases (choose | site search | pick | before-counts | rewrite |
// after-counts).  Same planes, same row table, same results as the one-warp kernel; selected automatically for
// batches of at most one replica per SM (and for every MoveMonovacancy rule set).
#pragma once
#include "replica_wide_kernel.cuh"

namespace kmc {

// -DKMC_WIDE_TRACE: thread 0 of CTA 0 adds the cycles since the start of the event after each of the six barriers to
// g_wide_trace (slot 15 counts the events); capi.cu prints the averages after a launch.
#ifdef KMC_WIDE_TRACE
__device__ long long g_wide_trace[16];
#define WTRACE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd((unsigned long long *)&g_wide_trace[i], (unsigned long long)(clock64() - wtr0)); } while (0)
#else
#define WTRACE(i) do { } while (0)
#endif

// Row tables: one per class of the reference's own rules; with MoveMonovacancy (MM build) three more -- natural (13),
// merge-divacancy (14), swap-divacancy (15th table, class 16); split-divacancy (class 15) has none: its row count
// is twice the sum of the row's four MoveDivacancy counts (see replica_hw_kernel.cuh).
constexpr int WIDE_NCT_MM = 16;
__host__ __device__ inline int wide_tables(bool mm) { return mm ? WIDE_NCT_MM : NC; }
__host__ __device__ inline size_t wide2_cta_bytes(int d0, bool mm = false) {
    const size_t d0p = ((size_t)d0 + 1) & ~(size_t)1;
    const size_t nct = (size_t)wide_tables(mm);
    return ((nct * d0p * 2 + 15) & ~(size_t)15) + nct * 32 * sizeof(uint32_t);    // row tables + 32 block sums per table
}
// table of a class in the MM build (-1: derived)
__device__ __forceinline__ int wide_table_of(int c) { return c < C_MOVE_MONO + 2 ? c : (c == C_MOVE_MONO + 3 ? 15 : -1); }

// MoveMonovacancy counts of the cell (row a, word w) in direction k: natural | merge << 16, and swap
__device__ __forceinline__ void mm_cell_counts(const WGeom &G, int a, int w, int k, uint32_t &nm, uint32_t &sw) {
    const u64 v = G.valid(w);
    const u64 m1 = window_at(G.rowp(T_M1, a), w, 0) & v, m2 = window_at(G.rowp(T_M2, a), w, 0) & v;
    const u64 dv = window_at(G.rowp(T_D, a), w, 0) & v;
    const int an = wrap1(a + nib(NBR_DA, k), G.d0), db = nib(NBR_DB, k);
    const u64 zn = window_at(G.rowp(T_Z, an), w, db), m1n = window_at(G.rowp(T_M1, an), w, db);
    const u64 m2n = window_at(G.rowp(T_M2, an), w, db);
    nm = (uint32_t)__popcll((m1 | m2) & zn) | ((uint32_t)__popcll((m1 & m2n) | (m2 & m1n)) << 16);
    sw = (uint32_t)__popcll(dv & (m1n | m2n));
}
// the sites of MoveMonovacancy kind q (class 13 + q) in slot (direction k, layer L) of the cell
__device__ __forceinline__ u64 mm_cell_mask(const WGeom &G, int a, int w, int k, int layer, int q) {
    const u64 v = G.valid(w);
    const int an = wrap1(a + nib(NBR_DA, k), G.d0), db = nib(NBR_DB, k);
    const u64 src = window_at(G.rowp(q < 2 ? (layer == 1 ? T_M1 : T_M2) : T_D, a), w, 0) & v;
    const u64 dst = window_at(G.rowp((q & 1) ? (layer == 1 ? T_M2 : T_M1) : T_Z, an), w, db);
    return src & dst;
}
// row tables 13, 14, 15 of the MM build from the planes: one thread per (replica, row)
__global__ void kmc_wide_hier_mm_kernel(const u64 *planes, uint16_t *hier, int R, int d0, int d1, int d0p)
{
    const int Wa = wide_anchor_words(d1), We = Wa + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)R * d0) return;
    const int a = (int)(g % d0);
    const long long rep = g / d0;
    WGeom G{planes + rep * N_TYPES * d0 * We, d0, d1, Wa, We};
    uint32_t nm = 0, sw = 0;
    for (int w = 0; w < Wa; w++)
        for (int k = 0; k < 6; k++) { uint32_t x, y; mm_cell_counts(G, a, w, k, x, y); nm += x; sw += y; }
    uint16_t *H = hier + rep * WIDE_NCT_MM * d0p;
    H[13 * d0p + a] = (uint16_t)nm; H[14 * d0p + a] = (uint16_t)(nm >> 16); H[15 * d0p + a] = (uint16_t)sw;
}

// one refresh task with warp-uniform operand descriptors (decoded once, outside the event loop)
struct TaskOps { int pl[3], da[3], o[3]; };
__device__ __forceinline__ TaskOps task_ops(int k) {
    const uint32_t desc = task_desc(k);
    TaskOps t;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint32_t od = (desc >> (7 * i)) & 0x7Fu;
        t.pl[i] = (int)(od & 1u); t.da[i] = (int)((od >> 1) & 7u) - 2; t.o[i] = (int)((od >> 4) & 7u);
    }
    return t;
}
__device__ __forceinline__ u64 op_word(const u64 *P, int d0, int We, int a, int w, int pl, int da, int o) {
    int ar = a + da;
    ar = ar < 0 ? ar + d0 : (ar >= d0 ? ar - d0 : ar);
    const u64 *rp = P + (uint32_t)(pl * d0 + ar) * (uint32_t)We + w;
    const uint2 x0 = *reinterpret_cast<const uint2 *>(rp);
    const uint32_t x1lo = *reinterpret_cast<const uint32_t *>(rp + 1);
    return (u64)__funnelshift_r(x0.x, x0.y, o) | ((u64)__funnelshift_r(x0.y, x1lo, o) << 32);
}
// counts of the task (row a, word w) of role k: k < 6 -> classes (2,3) in c01 and (4,5) in c23; k == 6 -> class 6 in c01
__device__ __forceinline__ void role_counts(const u64 *P, int d0, int We, int a, int w, int k, const TaskOps &t, u64 valid,
                                            uint32_t &c01, uint32_t &c23) {
    const u64 *dp = P + (uint32_t)(T_D * d0 + a) * (uint32_t)We + w;
    const u64 D0 = window(dp[0], dp[1], 0) & valid;
    const u64 A = op_word(P, d0, We, a, w, t.pl[0], t.da[0], t.o[0]);
    const u64 Bv = op_word(P, d0, We, a, w, t.pl[1], t.da[1], t.o[1]);
    const u64 C = op_word(P, d0, We, a, w, t.pl[2], t.da[2], t.o[2]);
    const u64 first = D0 & A;
    c01 = 0; c23 = 0;
    if (k < 6) kind_counts(first, Bv, C, c01, c23);
    else c01 = __popcll(first & Bv) + __popcll(first & C);
}

template <bool MM = false>
__global__ void __launch_bounds__(256, 1) kmc_replica_wide2_kernel(const WideParams wp)
{
    constexpr int NCT = MM ? WIDE_NCT_MM : NC;             // row tables
    constexpr int NCX = MM ? NCA : NC;                     // classes
    constexpr int NL = MM ? 32 : 16;                       // weight lanes of the class choice
    const ReplicaParams &p = wp.base;
    const int d0 = p.d0, d1 = p.d1, Wa = wp.Wa, We = wp.Wa + 1, d0p = wp.d0p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rep = blockIdx.x;

    extern __shared__ uint4 smem_raw[];
    uint16_t *rc16 = reinterpret_cast<uint16_t *>(smem_raw);
    uint32_t *rc32 = reinterpret_cast<uint32_t *>(rc16);
    uint32_t *bs = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(rc16) + ((NCT * (size_t)d0p * 2 + 15) & ~(size_t)15));
    __shared__ int s_csel, s_row, s_zero, s_tie, s_col, s_s0;
    __shared__ uint32_t s_rin, s_slot_tot[12];
    __shared__ double s_total;
    __shared__ int s_delta[20];                             // class total changes of the refresh (classes 2..6, 13..16)
    // speculative row search: warp 0 publishes the class totals and the in-class draw before it starts choosing the
    // class; warps 1..6 each search the row of one candidate class (the six enabled classes with the highest rates)
    // meanwhile.  If the chosen class is a candidate -- with DFT rates practically always -- its row is there when the
    // class choice ends; otherwise warp 0 searches as before.
    __shared__ int s_pub_tot[NCA], s_hit, s_spec_row[8];
    __shared__ uint32_t s_spec_rin[8];
    __shared__ double s_pub_u2;

    u64 *P = wp.planes + (size_t)rep * N_TYPES * d0 * We;
    WGeom G{P, d0, d1, Wa, We};
    uint16_t *ghier = wp.hier + (size_t)rep * NCT * d0p;
    for (int i = tid; i < (NCT * d0p) / 2; i += 256) rc32[i] = reinterpret_cast<const uint32_t *>(ghier)[i];
    if (tid < 20) s_delta[tid] = 0;
    __syncthreads();
    int bshift = 0;
    while ((32 << bshift) < d0) bshift++;
    const int B = 1 << bshift;
    for (int c = warp; c < NCT; c += 8) {
        uint32_t sum = 0;
        for (int q = 0; q < B; q++) { const int a = lane * B + q; if (a < d0) sum += rc16[c * d0p + a]; }
        bs[c * 32 + lane] = sum;
    }
    __syncthreads();

    // warp 0: class state (lane c owns class c)
    const int myc = lane < NCX ? lane : 0;
    const double rate = (lane < NCX && p.order_pos[myc] >= 0) ? p.rate[myc] : 0.0;
    const int pos = (lane < NCX && p.order_pos[myc] >= 0) ? p.order_pos[myc] : 100 + lane;
    int tot = 0;
    if (warp == 0 && lane < NCX) {
        if (MM && lane == C_MOVE_MONO + 2) {                // split-divacancy: twice the MoveDivacancy hops
            for (int q = 0; q < 4; q++) for (int i = 0; i < 32; i++) tot += 2 * (int)bs[(C_MOVE_DIV + q) * 32 + i];
        } else {
            const int tb = MM ? wide_table_of(lane) : lane;
            for (int i = 0; i < 32; i++) tot += (int)bs[tb * 32 + i];
        }
    }
    const uint32_t my_g1 = (lane < NC && !(lane >= C_MOVE_DIV && lane <= C_CREATE_TREF)) ? g1_const(lane) : 0u;
    PickState pstate = pick_state_init_n<NL>(lane);
    // warp 7: statistics
    unsigned long long hist = 0, hops = 0, n_ties = 0;
    double tclock = p.clock ? p.clock[rep] : 0.0;
    uint32_t *const tracer = p.tracer ? p.tracer + (size_t)rep * p.n : nullptr;
    // warps 0..6: the role's operands
    const TaskOps ops = task_ops(warp < 7 ? warp : 6);
    const int r2 = B > 32 ? B >> 5 : 1;

    int my_cand = -1;                                      // warps 1..6: the class whose row this warp searches ahead
    int cand_warp_of_my_class = -1;                        // warp 0, lane c: the warp that searches class c (or -1)
    {
        auto live = [&](int c) { return p.order_pos[c] >= 0 && p.rate[c] > 0.0; };
        auto rate_rank = [&](int c) {                       // number of enabled classes that come before c by rate
            int rk = 0;
            for (int o = 0; o < NCX; o++)
                if (o != c && live(o) && (p.rate[o] > p.rate[c] || (p.rate[o] == p.rate[c] && o < c))) rk++;
            return rk;
        };
        for (int c = 0; c < NCX; c++)
            if (live(c) && warp >= 1 && warp <= 6 && rate_rank(c) == warp - 1) my_cand = c;
        if (lane < NCX && live(lane)) { const int rk = rate_rank(lane); if (rk < 6) cand_warp_of_my_class = rk + 1; }
    }
    const unsigned long long step_base = p.steps_done[rep];
    double mu1 = 0.0, mu2 = 0.0;
    long long t = 0;
    uint32_t flag = 0;

    // row and rank inside the row of the in-class draw u2 among the `cnt` moves of class c: one warp, two rank searches
    // over the class's block sums and row counts in shared memory
    auto row_search = [&](int c, int cnt, double u2v, int &row, uint32_t &r) {
        long long rr = (long long)__dmul_rn(u2v, (double)cnt);
        if (rr > (long long)cnt - 1) rr = (long long)cnt - 1;
        r = (uint32_t)rr;
        // row count of the class: its table, or (split-divacancy) twice the four MoveDivacancy tables
        const bool split = MM && c == C_MOVE_MONO + 2;
        const int tb = MM ? (split ? C_MOVE_DIV : wide_table_of(c)) : c;
        auto blocksum = [&](int i) -> uint32_t {
            return split ? 2u * (bs[C_MOVE_DIV * 32 + i] + bs[(C_MOVE_DIV + 1) * 32 + i] + bs[(C_MOVE_DIV + 2) * 32 + i] +
                                 bs[(C_MOVE_DIV + 3) * 32 + i]) : bs[tb * 32 + i];
        };
        auto rowcount = [&](int a) -> uint32_t {
            return split ? 2u * ((uint32_t)rc16[C_MOVE_DIV * d0p + a] + rc16[(C_MOVE_DIV + 1) * d0p + a] +
                                 rc16[(C_MOVE_DIV + 2) * d0p + a] + rc16[(C_MOVE_DIV + 3) * d0p + a])
                         : (uint32_t)rc16[tb * d0p + a];
        };
        int L = 0;
        warp_rank_search(blocksum(lane), r, L, lane);
        const int a0 = (L << bshift) + lane * r2;
        uint32_t sum = 0;
        for (int q = 0; q < r2; q++)
            if (lane * r2 + q < B && a0 + q < d0) sum += rowcount(a0 + q);
        int L2 = 0;
        warp_rank_search(sum, r, L2, lane);
        row = (L << bshift) + L2 * r2;
        for (int q = 0; q < r2 - 1; q++) {
            const uint32_t v = rowcount(row);
            if (r < v) break;
            r -= v; row++;
        }
    };
    for (; t < p.nsteps; t++) {
#ifdef KMC_WIDE_TRACE
        const long long wtr0 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd((unsigned long long *)&g_wide_trace[15], 1ull);
#endif
        // ================= phase 1: draws and class choice (warp 0), row search (ahead: warps 1..6) =======
        double u1 = 0.0, u2 = 0.0;
        if (warp == 0) {
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
            u1 = __shfl_sync(FULL, mu1, (int)(t & 31));
            u2 = __shfl_sync(FULL, mu2, (int)(t & 31));
            if (lane < NCX) s_pub_tot[lane] = tot;
            if (lane == 0) s_pub_u2 = u2;
        }
        if (tid < 12) s_slot_tot[tid] = 0;
        // B0 (warps 0..6 only: warp 7 is still writing the previous event's record and statistics, which overlaps the
        // class choice): totals and the in-class draw are public
        if (warp < 7) asm volatile("bar.sync 1, 224;" ::: "memory");
        if (warp == 0) {
            const double w = __dmul_rn((double)tot, rate);
            const ClassPick pick = pick_class<NL>(w, pos, pstate, u1, lane);
            WTRACE(7);
            if (p.log_mode == KMC_LOG_FULL && !pick.zero) {
                kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) +
                                      ((size_t)rep * (size_t)p.nsteps + (size_t)t);
                if (lane < NCA) rec->counts[lane] = lane < NCX ? (uint32_t)tot : 0u;
            }
            int row = 0, hit = -1;
            uint32_t r = 0;
            if (!pick.zero) {
                hit = __shfl_sync(FULL, cand_warp_of_my_class, pick.csel);     // the warp that searched this class ahead
                if (hit < 0) row_search(pick.csel, __shfl_sync(FULL, tot, pick.csel), u2, row, r);
            }
            if (lane == 0) {
                s_csel = pick.csel; s_zero = pick.zero; s_tie = pick.tie != 0; s_total = pick.total;
                s_row = row; s_rin = r; s_hit = hit;
            }
        } else if (my_cand >= 0) {
            const int cnt = s_pub_tot[my_cand];
            int row = 0;
            uint32_t r = 0;
            if (cnt > 0) row_search(my_cand, cnt, s_pub_u2, row, r);
            if (lane == 0) { s_spec_row[warp] = row; s_spec_rin[warp] = r; }
        }
        __syncthreads();                                               // B1
        WTRACE(1);
        if (s_zero) {                                                  // kmc.py:31-32: nothing can happen
            flag |= FLAG_ZERO_RATE;
            if (p.log_mode != KMC_LOG_NONE && tid == 0) {
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
        const int csel = s_csel, row = s_hit >= 0 ? s_spec_row[s_hit] : s_row;
        const double total = s_total;
        const int tie = s_tie;

        // ================= phase 2: the class's slot masks of the chosen row, one role per warp ================
        // (warp 7 meanwhile starts the plane rows of the refresh moving: Z and D of rows row-3 .. row+3)
        const int group = (csel >= C_MOVE_DIV && csel < C_MOVE_DIV + 4) ? 1 : (csel == C_CREATE_TREF ? 2 :
                          ((MM && csel >= C_MOVE_MONO) ? 3 : 0));
        u64 m = 0, m_l2 = 0;                                   // (MM group: the warp's direction, layers 1 and 2)
        if (warp == 7) {
            const int lines = (We * 8 + 127) >> 7;
            const int j = lane >> 2, pl = (lane >> 1) & 1, li = lane & 1;
            if (j < 7 && li < lines) {
                const int a = wrap1(wrap1(row - 3 + j, d0), d0);
                asm volatile("prefetch.global.L1 [%0];" :: "l"(P + ((size_t)pl * d0 + a) * We + min(li * 16, We - 1)));
            }
        } else if (lane < Wa) {
            if (group == 1) {
                if (warp < 6) {
                    Nbhd nb;
                    load_nbhd<false>(G, row, lane, nb);
                    u64 pr, ne, fa;
                    cell_dir_masks(nb, warp, pr, ne, fa);
                    m = kind_select(pr, ne, fa, csel - C_MOVE_DIV);
                }
            } else if (MM && group == 3) {
                if (warp < 6) {
                    m = mm_cell_mask(G, row, lane, warp, 1, csel - C_MOVE_MONO);
                    m_l2 = mm_cell_mask(G, row, lane, warp, 2, csel - C_MOVE_MONO);
                }
            } else if (group == 2) {
                if (warp < 2) {
                    const u64 *ra = G.rowp(T_D, row), *rb2 = G.rowp(T_D, wrap1(row + 2, d0));
                    const u64 a0 = ra[lane], a1 = ra[lane + 1], b0 = rb2[lane], b1 = rb2[lane + 1];
                    const u64 both = window(a0, a1, 0) & window(b0, b1, 0) & G.valid(lane);
                    m = warp == 0 ? (both & window(a0, a1, 2)) : (both & window(b0, b1, -2));
                }
            } else if (warp < 2) {
                int pa = T_M1, pb = -1;
                switch (csel) {
                    case C_CREATE_DIV: pa = T_Z; break;
                    case C_FILL_DIV: pa = T_D; break;
                    case C_DESTROY_TREF: pa = T_A4; pb = T_A7; break;
                    case C_CMONO_FROM_DOUBLE: pa = T_Z; pb = T_Z; break;
                    case C_CMONO_MAKE_EMPTY: pa = T_M2; pb = T_M1; break;
                    case C_FMONO_MAKE_DOUBLE: pa = T_M1; pb = T_M2; break;
                    case C_FMONO_FROM_EMPTY: pa = T_D; pb = T_D; break;
                    default: break;
                }
                const u64 v = G.valid(lane);
                if (warp == 0) {
                    m = window_at(G.rowp(pa, row), lane, 0) & v;
                    if (csel == C_FLIP) m |= window_at(G.rowp(T_M2, row), lane, 0) & v;
                } else if (pb >= 0) {
                    m = window_at(G.rowp(pb, row), lane, 0) & v;
                }
            }
        }
        uint32_t mcnt = __popcll(m);
        const uint32_t mcnt2 = MM ? (uint32_t)__popcll(m_l2) : 0u;
        if (warp < 6) {
            if (MM && group == 3) {                                    // slots (direction, layer) = 2 * warp + (L - 1)
                const uint32_t wt = __reduce_add_sync(FULL, mcnt | (mcnt2 << 16));
                if (lane == 0) { s_slot_tot[2 * warp] = wt & 0xFFFFu; s_slot_tot[2 * warp + 1] = wt >> 16; }
            } else {
                const uint32_t wt = __reduce_add_sync(FULL, mcnt);
                if (lane == 0) s_slot_tot[warp] = wt;
            }
        }
        __syncthreads();                                               // B2
        WTRACE(2);
        uint32_t r = s_hit >= 0 ? s_spec_rin[s_hit] : s_rin;
        int j = 0;
#pragma unroll
        for (int q = 0; q < (MM ? 11 : 5); q++)
            if (j == q && r >= s_slot_tot[q]) { r -= s_slot_tot[q]; j = q + 1; }
        const int jwarp = (MM && group == 3) ? j >> 1 : j;
        if (MM && group == 3 && (j & 1)) { m = m_l2; mcnt = mcnt2; }
        if (warp == jwarp) {
            int L = 0;
            warp_rank_search(mcnt, r, L, lane);                        // r: rank inside word L of slot j
            const u64 mj = __shfl_sync(FULL, m, L);
            const int bit = nth_bit64(mj, r);
            if (lane == 0) {
                s_col = 64 * L + bit;
                int s0v = S_DIVAC;
                if (MM && group == 3) { if (csel < C_MOVE_MONO + 2) s0v = (j & 1) + 1; }    // the hopping layer
                else if (!group) switch (csel) {
                    case C_CREATE_DIV: case C_CMONO_FROM_DOUBLE: s0v = 0; break;
                    case C_FILL_DIV: case C_FMONO_FROM_EMPTY: s0v = 3; break;
                    case C_DESTROY_TREF: s0v = j ? 7 : 4; break;
                    case C_CMONO_MAKE_EMPTY: s0v = j ? 1 : 2; break;
                    case C_FMONO_MAKE_DOUBLE: s0v = j ? 2 : 1; break;
                    default: s0v = ((window_at(G.rowp(T_M1, row), L, 0) >> bit) & 1ull) ? 1 : 2; break;
                }
                s_s0 = s0v;
            }
        }
        __syncthreads();                                               // B3
        WTRACE(3);
        const int col = s_col, s0 = s_s0;
        const int slot = class_slot_base(csel) + j;
        const int site = row * d1 + col;

        // ================= the move (decoded by every thread: uniform, cheap) ====================================
        MoveNodes mv;
        int os1, os2, db1, db2;
        {
            const MoveEntry e = MOVE_TABLE[slot];
            mv.ns0 = (int)((((uint32_t)s0 | (e.w0 & 0xF)) & ((e.w0 >> 4) & 0xF)) ^ ((e.w0 >> 8) & 0xF));
            mv.na = (int)((e.w0 >> 12) & 3);
            mv.da1 = (int)(e.w1 & 0xF) - 2;
            const int da2 = (int)((e.w1 >> 8) & 0xF) - 2;
            mv.a1 = wrap1(row + mv.da1, d0);
            db1 = (int)((e.w1 >> 4) & 0xF) - 2;
            db2 = (int)((e.w1 >> 12) & 0xF) - 2;
            mv.b1 = wrap1(col + db1, d1);
            mv.a2 = wrap1(row + da2, d0);
            mv.b2 = wrap1(col + db2, d1);
            mv.ns1 = (int)((e.w1 >> 16) & 0xF); mv.ns2 = (int)((e.w1 >> 20) & 0xF);
            os1 = (int)((e.w1 >> 24) & 0xF); os2 = (int)(e.w1 >> 28);
            if (MM && slot >= MM_SLOT0) {
                const int layer = ((slot - MM_SLOT0) & 1) + 1;
                const bool onto_mono = (csel - C_MOVE_MONO) & 1;
                os1 = onto_mono ? 3 - layer : 0;
                mv.ns1 = onto_mono ? 3 : layer;
            }
        }
        const int na = mv.na;
        // words holding a column within 2 of a touched column (see replica_wide_kernel.cuh)
        uint32_t wmask;
        {
            const int lo = col + min(0, min(db1, db2)) - 2, hi = col + max(0, max(db1, db2)) + 2;
            if (lo >= 0 && hi < d1) wmask = (1u << (lo >> 6)) | (1u << (hi >> 6));
            else wmask = (1u << ((lo < 0 ? lo + d1 : lo) >> 6)) | (1u << ((hi >= d1 ? hi - d1 : hi) >> 6)) | 1u | (1u << (Wa - 1));
        }
        const int nw = __popc(wmask);
        const int wA = __ffs((int)wmask) - 1, wB = 31 - __clz((int)wmask);
        const int wM = nw == 3 ? __ffs((int)(wmask & (wmask - 1))) - 1 : wA;
        int lo2, n2, n3;
        uint32_t g3;
        if (na == 1) { lo2 = -1; n2 = 3; n3 = 2; g3 = 0x13u; }
        else if (na == 2) {
            lo2 = (mv.da1 < 0 ? mv.da1 : 0) - 1; n2 = (mv.da1 ? 4 : 3);
            n3 = mv.da1 ? 4 : 2;
            g3 = 0x13u | ((uint32_t)(mv.da1 + 3) << 8) | ((uint32_t)(mv.da1 + 1) << 12);
        } else { lo2 = -1; n2 = 5; n3 = 3; g3 = 0x531u; }
        if (n2 > d0) n2 = d0;

        // ================= phase 3: before-counts of the role's tasks; warp 7 does the bookkeeping =================
        // task of lane l: (row index u = l / nw, word iw = l % nw); at most 5 x 3 tasks per role
        int ta = 0, twd = 0;
        bool tact = false;
        if (warp < 7) {
            const int ntk = (warp < 6 ? n2 : n3) * nw;
            const int u = nw == 1 ? lane : (nw == 2 ? lane >> 1 : (lane * 171) >> 9);
            const int iw = lane - u * nw;
            twd = iw == 0 ? wA : (iw == nw - 1 ? wB : wM);
            int a = warp < 6 ? row + lo2 + u : row + (int)((g3 >> (4 * (u & 7))) & 0xF) - 3;
            ta = wrap1(wrap1(a, d0), d0);
            tact = lane < ntk;
        }
        uint32_t b01 = 0, b23 = 0, bnm = 0, bsw = 0;
        if (tact) {
            role_counts(P, d0, We, ta, twd, warp, ops, G.valid(twd), b01, b23);
            if (MM && warp < 6) mm_cell_counts(G, ta, twd, warp, bnm, bsw);
        }
        if (warp == 7) {
            if (p.clock && lane == 0)
                tclock = __dadd_rn(tclock, kmc_time_step(p.clock_mode, total, p.seed, p.replica0 + (uint32_t)rep,
                                                         step_base + (unsigned long long)t));
            if (p.log_mode != KMC_LOG_NONE && lane == 0) {
                const size_t idx = (size_t)rep * (size_t)p.nsteps + (size_t)t;
                if (p.log_mode == KMC_LOG_FULL) {
                    kmc_event_full *rec = reinterpret_cast<kmc_event_full *>(p.events) + idx;
                    rec->site = (uint32_t)site; rec->cls = (uint8_t)csel; rec->slot = (uint8_t)slot;
                    rec->flags = tie ? 1 : 0; rec->total_rate = total; rec->pad = 0;
                } else {
                    kmc_event_compact e;
                    e.site = (uint32_t)site; e.cls = (uint8_t)csel; e.slot = (uint8_t)slot;
                    e.flags = tie ? 1 : 0; e.total_rate = total;
                    reinterpret_cast<kmc_event_compact *>(p.events)[idx] = e;
                }
            }
            if (lane == csel) hist++;
            if (lane == slot - 7 && slot < 13) hops++;
            if (tie && lane == 0) n_ties++;
            if (tracer && lane == 31)
                tracer_update(tracer, slot, na, site, mv.ns0, mv.a1 * d1 + mv.b1, mv.ns1, mv.a2 * d1 + mv.b2, mv.ns2, s0);
            // own-status classes: row table and block sums (owner lanes); the totals are folded by warp 0 below
            if (my_g1) {
                int d = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
                if (d) { rc16[lane * d0p + row] = (uint16_t)(rc16[lane * d0p + row] + d); bs[lane * 32 + (row >> bshift)] += d; }
                if (na > 1) {
                    d = (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
                    if (d) { rc16[lane * d0p + mv.a1] = (uint16_t)(rc16[lane * d0p + mv.a1] + d); bs[lane * 32 + (mv.a1 >> bshift)] += d; }
                }
                if (na > 2) {
                    d = (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
                    if (d) { rc16[lane * d0p + mv.a2] = (uint16_t)(rc16[lane * d0p + mv.a2] + d); bs[lane * 32 + (mv.a2 >> bshift)] += d; }
                }
            }
        }
        if (warp == 0 && my_g1) {
            int dsum = (int)((my_g1 >> (2 * mv.ns0)) & 3u) - (int)((my_g1 >> (2 * s0)) & 3u);
            if (na > 1) dsum += (int)((my_g1 >> (2 * mv.ns1)) & 3u) - (int)((my_g1 >> (2 * os1)) & 3u);
            if (na > 2) dsum += (int)((my_g1 >> (2 * mv.ns2)) & 3u) - (int)((my_g1 >> (2 * os2)) & 3u);
            tot += dsum;
        }
        __syncthreads();                                               // B4: every before-load has its data
        WTRACE(4);

        // ================= phase 4 (warp 7): rewrite the planes; lane pl < 6 owns plane pl ===========================
        if (warp == 7 && lane < N_TYPES) {
            const int as[3] = {row, mv.a1, mv.a2}, bcol[3] = {col, mv.b1, mv.b2}, ns[3] = {mv.ns0, mv.ns1, mv.ns2};
            const int os[3] = {s0, os1, os2};
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < na && (status_plane(os[i]) == lane || status_plane(ns[i]) == lane)) {
                    u64 *rowp = P + (uint32_t)((lane * d0 + as[i]) * We);
                    const bool set = status_plane(ns[i]) == lane;
                    const int q0 = bcol[i] + 2;
                    const int q1 = bcol[i] < 2 ? bcol[i] + d1 + 2 : (bcol[i] >= d1 - 2 ? bcol[i] + 2 - d1 : -1);
                    {
                        u64 *wptr = rowp + (q0 >> 6);
                        const u64 bit = 1ull << (q0 & 63);
                        *wptr = (*wptr & ~bit) | (set ? bit : 0ull);
                    }
                    if (q1 >= 0) {
                        u64 *wptr = rowp + (q1 >> 6);
                        const u64 bit = 1ull << (q1 & 63);
                        *wptr = (*wptr & ~bit) | (set ? bit : 0ull);
                    }
                }
            __threadfence_block();
        }
        __syncthreads();                                               // B5: the new planes are visible
        WTRACE(5);

        // ================= phase 5: after-counts, differences to the row table and the class totals ================
        if (warp < 7) {
            uint32_t a01 = 0, a23 = 0;
            if (tact) {
                // (volatile-free reload: the barrier above orders these loads after warp 7's stores)
                role_counts(P, d0, We, ta, twd, warp, ops, G.valid(twd), a01, a23);
                const int dd[4] = {(int)(a01 & 0xFFFF) - (int)(b01 & 0xFFFF), (int)(a01 >> 16) - (int)(b01 >> 16),
                                   (int)(a23 & 0xFFFF) - (int)(b23 & 0xFFFF), (int)(a23 >> 16) - (int)(b23 >> 16)};
                if (warp < 6) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (dd[q]) {
                            atomicAdd(&rc32[((C_MOVE_DIV + q) * d0p + ta) >> 1], (uint32_t)dd[q] << (16 * (ta & 1)));
                            atomicAdd(&bs[(C_MOVE_DIV + q) * 32 + (ta >> bshift)], (uint32_t)dd[q]);
                        }
                } else if (dd[0]) {
                    atomicAdd(&rc32[(C_CREATE_TREF * d0p + ta) >> 1], (uint32_t)dd[0] << (16 * (ta & 1)));
                    atomicAdd(&bs[C_CREATE_TREF * 32 + (ta >> bshift)], (uint32_t)dd[0]);
                }
            }
            if (MM && warp < 6) {
                uint32_t anm = 0, asw = 0;
                if (tact) {
                    mm_cell_counts(G, ta, twd, warp, anm, asw);
                    const int dm[3] = {(int)(anm & 0xFFFF) - (int)(bnm & 0xFFFF), (int)(anm >> 16) - (int)(bnm >> 16),
                                       (int)asw - (int)bsw};
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        if (dm[q]) {
                            atomicAdd(&rc32[((13 + q) * d0p + ta) >> 1], (uint32_t)dm[q] << (16 * (ta & 1)));
                            atomicAdd(&bs[(13 + q) * 32 + (ta >> bshift)], (uint32_t)dm[q]);
                        }
                }
                int n_lo, n_hi, s_lo, s_hi;
                reduce_packed_diff(anm, bnm, n_lo, n_hi);
                reduce_packed_diff(asw, bsw, s_lo, s_hi);
                if (lane == 0) {
                    if (n_lo) atomicAdd(&s_delta[C_MOVE_MONO], n_lo);
                    if (n_hi) atomicAdd(&s_delta[C_MOVE_MONO + 1], n_hi);
                    if (s_lo) atomicAdd(&s_delta[C_MOVE_MONO + 3], s_lo);
                }
            }
            int d_lo, d_hi, e_lo, e_hi;
            reduce_packed_diff(a01, b01, d_lo, d_hi);
            if (warp < 6) {
                reduce_packed_diff(a23, b23, e_lo, e_hi);
                if (lane == 0) {
                    if (d_lo) atomicAdd(&s_delta[C_MOVE_DIV], d_lo);
                    if (d_hi) atomicAdd(&s_delta[C_MOVE_DIV + 1], d_hi);
                    if (e_lo) atomicAdd(&s_delta[C_MOVE_DIV + 2], e_lo);
                    if (e_hi) atomicAdd(&s_delta[C_MOVE_DIV + 3], e_hi);
                }
            } else if (lane == 0) {
                // (a trefoil count is a plain 32-bit difference in c01: the packed reduction's low half holds it)
                const int dt = d_lo + (d_hi << 16);
                if (dt) atomicAdd(&s_delta[C_CREATE_TREF], dt);
            }
        }
        __syncthreads();                                               // B6
        WTRACE(6);
        if (warp == 0) {
            int d = 0;
            if (lane >= C_MOVE_DIV && lane <= C_CREATE_TREF) d = s_delta[lane];
            if (MM && (lane == C_MOVE_MONO || lane == C_MOVE_MONO + 1 || lane == C_MOVE_MONO + 3)) d = s_delta[lane];
            if (MM && lane == C_MOVE_MONO + 2)    // split-divacancy = 2 x MoveDivacancy hops
                d = 2 * (s_delta[C_MOVE_DIV] + s_delta[C_MOVE_DIV + 1] + s_delta[C_MOVE_DIV + 2] + s_delta[C_MOVE_DIV + 3]);
            tot += d;
            __syncwarp();
            if (lane < 20) s_delta[lane] = 0;     // (written again only after B5 of the next event)
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------------------
    __syncthreads();
    const long long done = t;
    for (int i = tid; i < (NCT * d0p) / 2; i += 256) reinterpret_cast<uint32_t *>(ghier)[i] = rc32[i];
    if (warp == 0) {
        if (lane < NCX) {
            int fresh = 0, blocks = 0;
            if (MM && lane == C_MOVE_MONO + 2) {
                for (int q = 0; q < 4; q++) {
                    for (int a = 0; a < d0; a++) fresh += 2 * rc16[(C_MOVE_DIV + q) * d0p + a];
                    for (int i = 0; i < 32; i++) blocks += 2 * (int)bs[(C_MOVE_DIV + q) * 32 + i];
                }
            } else {
                const int tb = MM ? wide_table_of(lane) : lane;
                for (int a = 0; a < d0; a++) fresh += rc16[tb * d0p + a];
                for (int i = 0; i < 32; i++) blocks += (int)bs[tb * 32 + i];
            }
            if (fresh != tot || blocks != tot) flag |= FLAG_VALIDATE;      // totals and block sums vs the row table
        }
        flag = __reduce_or_sync(FULL, flag);
        if (lane == 0) {
            p.steps_done[rep] = step_base + (unsigned long long)done;
            if (flag) { p.flags[rep] |= flag; atomicOr(p.any_flag, flag); }
        }
    }
    if (warp == 7) {
        if (lane < NCX) p.hist[(size_t)rep * NCA + lane] += hist;
        if (lane < 6) p.hops[(size_t)rep * 6 + lane] += hops;
        if (lane == 0) {
            if (p.clock) p.clock[rep] = tclock;
            p.ties[rep] += n_ties;
        }
    }
}

}  // namespace kmc
