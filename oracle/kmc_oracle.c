/*
 * kmc_oracle.c -- CPU restatement of the kmc-dichalcogen hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA engine: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (kmc-dichalcogen_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement, event by event
 * and site by site, against fixtures under tests/golden/ that were produced by running the
 * unmodified reference (/root/reference/demo1) under injected draws (oracle/ref_oracle.py,
 * oracle/make_golden.py).
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   lattice + PBC stencils .......... demo1/state.py:383-464, hexagonal.py:56-62
 *   site predicates / mutators ...... demo1/state.py:195-299
 *   full enumeration (all_moves) .... demo1/rules.py:71,88,111-124,154-174,189,247-255,266-274,289
 *   incremental refresh ............. demo1/rules.py:22-27,46-51 (pre/post_status_change) and the
 *                                     per-rule moves_dependent_on (:74,:91,:115-117,:159-174,:192-198,
 *                                     :220-227,:292), driven by demo1/sim.py:145-157
 *   move application ................ demo1/rules.py perform(): :62,:79,:96-99,:143-146,:178-181,
 *                                     :204-210,:279-281
 *   class weights and selection ..... demo1/sim.py:117-122, demo1/kmc.py:21-54, demo1/util.py:32-46,65-68
 *   per-class counts ................ demo1/incremental.py:103-110,287-294 (value_counts)
 *   total_rate ...................... demo1/sim.py:138 (math.fsum)
 *
 * Two things the reference leaves to history / hash order are fixed canonically here, exactly as
 * in oracle/ref_oracle.py:
 *   - tie-break among equal class weights: position in `class_order` (rules in config order,
 *     kinds in rule.kinds() order); the reference uses Counter insertion order (sim.py:101);
 *   - in-class pick: rank min((int)(u2*n), n-1) in (row, slot, column) order, i.e. rows ascending, inside
 *     a row the move slots in slot order, inside a slot the sites by column; the reference picks
 *     rng.choice() from a list whose order is a swap-remove artefact (incremental.py:283-326).
 * The probability law is the reference's (uniform within a class).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -pthread  (no -ffast-math: fp64 order matters).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define NC 17          /* event classes (rule:kind): 13 of the reference's rules + 4 of MoveMonovacancy */
#define NS 29          /* move slots per site: 17 + 12 MoveMonovacancy (direction, layer) */
#define MM_SLOT0 17
#define NONE 0xFF

/* status codes: 0 pristine, 1/2 monovacancy layer, 3 divacancy, 4+3*orient+role trefoil member */
enum { PRISTINE = 0, DIVAC = 3, TREF = 4 };

/* class ids */
enum { C_CREATE_DIV = 0, C_FILL_DIV = 1, C_MOVE_DIV = 2 /*..5*/, C_CREATE_TREF = 6, C_DESTROY_TREF = 7,
       C_CMONO_FROM_DOUBLE = 8, C_CMONO_MAKE_EMPTY = 9, C_FMONO_MAKE_DOUBLE = 10, C_FMONO_FROM_EMPTY = 11,
       C_FLIP = 12,
       /* MoveMonovacancy: a stub in the reference (rules.py:295-296; kinds named in config/WSe2.yaml:8-13).
        * Semantics stated in oracle/move_monovacancy.py and pinned by fixtures made from the reference engine
        * running that rule class: kind = 2 * (source is a divacancy) + (target is not pristine). */
       C_MOVE_MONO = 13 /* natural, 14 merge-divacancy, 15 split-divacancy, 16 swap-divacancy */ };

/* Grid.neighbors_and_mutuals order (state.py:443-459): nbr, near, far */
static const int NBR[6][2]  = {{-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {1, -1}};
static const int NEAR[6][2] = {{0, 1}, {-1, 0}, {1, -1}, {-1, 1}, {0, -1}, {1, 0}};
static const int FAR[6][2]  = {{-1, 0}, {1, -1}, {0, 1}, {1, 0}, {-1, 1}, {0, -1}};
/* Grid.neighbors order (state.py:392-394); only set membership matters (distance <= 1) */
static const int RING[6][2] = {{-1, 0}, {0, -1}, {1, -1}, {1, 0}, {0, 1}, {-1, 1}};
/* trefoil = three mutual trefoil_neighbors (state.py:396-401,435-441) = triangle of side 2.
 * Anchored forms: Up(p) = {p, p+(2,0), p+(0,2)}, Down(p) = {p, p+(2,0), p+(2,-2)}. */
static const int TMEM[2][3][2] = {{{0, 0}, {2, 0}, {0, 2}}, {{0, 0}, {2, 0}, {2, -2}}};

typedef struct {
    uint32_t site;        /* linear index a*d1+b of the move's anchor site        */
    uint8_t  cls;         /* class id 0..12 ; 0xFF = total weight zero (kmc.py:32) */
    uint8_t  slot;        /* move slot 0..28                                       */
    uint16_t flags;       /* bit0: x landed exactly on a cumulative-weight boundary */
    double   total_rate;  /* math.fsum of the class weights (sim.py:138)           */
    double   total_naive; /* sequential prefix-sum total used for selection (kmc.py:41-42) */
    uint32_t counts[NC];  /* per-class move counts BEFORE the event               */
    uint32_t pad;
} KmcoEvent;              /* 96 bytes */

typedef struct {
    int d0, d1, n;
    double rate[NC];
    int order[NC];        /* class ids in weighted_choice input order */
    int n_order;
    uint8_t *status;      /* [n]                        */
    uint8_t *slot;        /* [n][NS] class id or NONE   */
    int64_t count[NC];
    int32_t *rowcnt;      /* [d0][NC]                   */
    int64_t hist[NC];
    int64_t n_ties;
    int64_t steps_done;
    /* extensions (the reference has neither, README.md:36-37): KMC clock and divacancy tracers */
    int clock_mode;       /* 0 off, 1 t += 1/R, 2 t += -ln(u3)/R */
    double clock;
    uint64_t clk_seed, clk_step; uint32_t clk_replica;   /* Philox coordinates of the current event's u3 */
    int32_t *trc;         /* [n][2] unwrapped axial displacement of the divacancy at each site, or NULL */
} KmcOracle;

/* ------------------------------------------------------------------------------------------ */
static inline int wrapi(int x, int d) { x %= d; return x < 0 ? x + d : x; }   /* Grid.reduce, state.py:461 */
static inline int at(int d0, int d1, int a, int b) { return wrapi(a, d0) * d1 + wrapi(b, d1); }

/* Full per-site enumeration: which of the 17 move slots exist at site (a,b), and their class.
 * This is the restatement of every rule's all_moves() restricted to moves anchored at one site. */
static void eval_site(const uint8_t *s, int d0, int d1, int a, int b, uint8_t out[NS])
{
    int st = s[at(d0, d1, a, b)];
    for (int j = 0; j < NS; j++) out[j] = NONE;
    /* CreateDivacancy: state.pristine_nodes() (rules.py:71) */
    if (st == PRISTINE) out[0] = C_CREATE_DIV;
    /* FillDivacancy: divacancies (rules.py:88) */
    if (st == DIVAC) out[1] = C_FILL_DIV;
    /* Create/FillMonovacancy: for layer in [1,2], new != old (rules.py:220-227) */
    if (st <= 3) {                                   /* has_defined_layerset */
        for (int layer = 1; layer <= 2; layer++) {
            if ((st | layer) != st)                  /* CreateMonovacancy.new_layerset = old | layer */
                out[1 + layer] = (st == PRISTINE) ? C_CMONO_FROM_DOUBLE : C_CMONO_MAKE_EMPTY;
            if ((st & ~layer) != st)                 /* FillMonovacancy.new_layerset = old & ~layer */
                out[3 + layer] = (st == DIVAC) ? C_FMONO_FROM_EMPTY : C_FMONO_MAKE_DOUBLE;
        }
    }
    /* FlipMonovacancy: monovacancies (rules.py:289) */
    if (st == 1 || st == 2) out[6] = C_FLIP;
    /* MoveDivacancy: moves_from_node (rules.py:119-136) */
    if (st == DIVAC) {
        for (int k = 0; k < 6; k++) {
            if (s[at(d0, d1, a + NBR[k][0], b + NBR[k][1])] == PRISTINE) {
                int near = s[at(d0, d1, a + NEAR[k][0], b + NEAR[k][1])] == DIVAC;
                int far  = s[at(d0, d1, a + FAR[k][0],  b + FAR[k][1])]  == DIVAC;
                out[7 + k] = (uint8_t)(C_MOVE_DIV + near + 2 * far);
            }
        }
        /* CreateTrefoil: three divacancies that are mutual trefoil neighbours (rules.py:159-174) */
        for (int o = 0; o < 2; o++) {
            int ok = 1;
            for (int r = 1; r < 3; r++)
                ok &= s[at(d0, d1, a + TMEM[o][r][0], b + TMEM[o][r][1])] == DIVAC;
            if (ok) out[13 + o] = C_CREATE_TREF;
        }
    }
    /* MoveMonovacancy (oracle/move_monovacancy.py): the vacancy in `layer` hops to neighbour k when the
     * neighbour has a defined layer set without that layer; slot 17 + 2k + (layer - 1) */
    if (st >= 1 && st <= 3) {
        for (int k = 0; k < 6; k++) {
            int dst = s[at(d0, d1, a + NBR[k][0], b + NBR[k][1])];
            if (dst > 3) continue;                   /* trefoil member: no defined layer set */
            for (int layer = 1; layer <= 2; layer++)
                if ((st & layer) && !(dst & layer))
                    out[MM_SLOT0 + 2 * k + (layer - 1)] = (uint8_t)(C_MOVE_MONO + 2 * (st == DIVAC) + (dst != 0));
        }
    }
    /* DestroyTrefoil: one move per existing trefoil (rules.py:189), anchored at its role-0 node */
    if (st == TREF + 0) out[15] = C_DESTROY_TREF;
    if (st == TREF + 3) out[16] = C_DESTROY_TREF;
}

/* Stateless full sweep (the --no-incremental path, sim.py:114-115 -> Rule.initialize_moves). */
void kmco_sweep(const uint8_t *status, int d0, int d1, uint8_t *slots /*[n][17] or NULL*/,
                int64_t counts[NC], int32_t *rowcnt /*[d0][NC] or NULL*/)
{
    uint8_t tmp[NS];
    for (int c = 0; c < NC; c++) counts[c] = 0;
    if (rowcnt) memset(rowcnt, 0, sizeof(int32_t) * (size_t)d0 * NC);
    for (int a = 0; a < d0; a++)
        for (int b = 0; b < d1; b++) {
            eval_site(status, d0, d1, a, b, tmp);
            if (slots) memcpy(slots + ((size_t)a * d1 + b) * NS, tmp, NS);
            for (int j = 0; j < NS; j++)
                if (tmp[j] != NONE) {
                    counts[tmp[j]]++;
                    if (rowcnt) rowcnt[a * NC + tmp[j]]++;
                }
        }
}

/* ------------------------------------------------------------------------------------------ */
KmcOracle *kmco_create(int d0, int d1, const double *rate, const int *order, int n_order)
{
    if (d0 < 1 || d1 < 1 || n_order < 0 || n_order > NC) return NULL;
    KmcOracle *o = (KmcOracle *)calloc(1, sizeof(KmcOracle));
    o->d0 = d0; o->d1 = d1; o->n = d0 * d1;
    memcpy(o->rate, rate, sizeof(double) * NC);
    for (int i = 0; i < n_order; i++) o->order[i] = order[i];
    o->n_order = n_order;
    o->status = (uint8_t *)calloc((size_t)o->n, 1);
    o->slot = (uint8_t *)malloc((size_t)o->n * NS);
    o->rowcnt = (int32_t *)calloc((size_t)d0 * NC, sizeof(int32_t));
    kmco_sweep(o->status, d0, d1, o->slot, o->count, o->rowcnt);
    return o;
}

void kmco_destroy(KmcOracle *o)
{
    if (!o) return;
    free(o->status); free(o->slot); free(o->rowcnt); free(o->trc); free(o);
}

/* ---- extensions: clock and tracers (same definitions as kmc-dichalcogen_b200/csrc/kmc_common.cuh) ---- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1);
/* u3 in (0, 1]: Philox counter domain 3 of the replica's stream */
static double draw3(uint64_t seed, uint32_t replica, uint64_t step)
{
    uint32_t c[4] = {(uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(seed >> 32), 3u};
    philox4x32_10(c, (uint32_t)seed, replica);
    return (double)(((((uint64_t)c[0] << 32) | c[1]) >> 11) + 1ull) * 0x1.0p-53;
}
/* natural logarithm as a fixed sequence of IEEE fp64 operations (no libm: libm's and CUDA's log differ in the
 * last bit; this one is the same everywhere).  Within 2 ulp of the correctly rounded value. */
double kmco_log(double x)
{
    uint64_t bits; memcpy(&bits, &x, 8);
    int e = (int)(bits >> 52) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m; memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double q = 1.0 / 23.0;
    q = q * z + 1.0 / 21.0; q = q * z + 1.0 / 19.0; q = q * z + 1.0 / 17.0; q = q * z + 1.0 / 15.0;
    q = q * z + 1.0 / 13.0; q = q * z + 1.0 / 11.0; q = q * z + 1.0 / 9.0;  q = q * z + 1.0 / 7.0;
    q = q * z + 1.0 / 5.0;  q = q * z + 1.0 / 3.0;  q = q * z + 1.0;
    const double lnm = (s + s) * q;
    const double ed = (double)e;
    return ed * 0x1.62e42fee00000p-1 + (ed * 0x1.a39ef35793c76p-33 + lnm);
}
double kmco_draw3(uint64_t seed, uint32_t replica, uint64_t step) { return draw3(seed, replica, step); }
void kmco_set_clock(KmcOracle *o, int mode) { o->clock_mode = mode; o->clock = 0.0; }
double kmco_get_clock(const KmcOracle *o) { return o->clock; }
void kmco_set_clock_key(KmcOracle *o, uint64_t seed, uint32_t replica) { o->clk_seed = seed; o->clk_replica = replica; }
void kmco_enable_tracers(KmcOracle *o)
{
    if (!o->trc) o->trc = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)o->n);
    memset(o->trc, 0, sizeof(int32_t) * 2 * (size_t)o->n);
}
/* per site (da, db), meaningful where the site holds a divacancy */
void kmco_get_tracers(const KmcOracle *o, int32_t *out) { memcpy(out, o->trc, sizeof(int32_t) * 2 * (size_t)o->n); }
/* sum over divacancies of da^2 + da*db + db^2, and their number */
void kmco_get_msd(const KmcOracle *o, int64_t *sumsq, int64_t *count)
{
    int64_t sq = 0, c = 0;
    for (int i = 0; i < o->n; i++)
        if (o->status[i] == DIVAC) {
            int64_t da = o->trc[2 * i], db = o->trc[2 * i + 1];
            sq += da * da + da * db + db * db; c++;
        }
    *sumsq = sq; *count = c;
}

/* KmcSim.initialize(state): replace the lattice and re-enumerate everything (sim.py:46-68). */
void kmco_set_status(KmcOracle *o, const uint8_t *status)
{
    memcpy(o->status, status, (size_t)o->n);
    if (o->trc) memset(o->trc, 0, sizeof(int32_t) * 2 * (size_t)o->n);      /* every divacancy is a new tracer */
    kmco_sweep(o->status, o->d0, o->d1, o->slot, o->count, o->rowcnt);
}
void kmco_get_status(const KmcOracle *o, uint8_t *out) { memcpy(out, o->status, (size_t)o->n); }
void kmco_get_slots(const KmcOracle *o, uint8_t *out) { memcpy(out, o->slot, (size_t)o->n * NS); }
void kmco_get_counts(const KmcOracle *o, int64_t *out) { memcpy(out, o->count, sizeof(int64_t) * NC); }
void kmco_get_hist(const KmcOracle *o, int64_t *out) { memcpy(out, o->hist, sizeof(int64_t) * NC); }
int64_t kmco_get_ties(const KmcOracle *o) { return o->n_ties; }
int64_t kmco_get_steps(const KmcOracle *o) { return o->steps_done; }
void kmco_reset_stats(KmcOracle *o)
{
    memset(o->hist, 0, sizeof o->hist); o->n_ties = 0; o->steps_done = 0; o->clock = 0.0;
    if (o->trc) memset(o->trc, 0, sizeof(int32_t) * 2 * (size_t)o->n);
}

/* Re-derive slots [j0, j1) at one site and fold the difference into the counts.
 * (clear_move before the mutation + add_move after it, collapsed: sim.py:149-157.) */
static void refresh_slots(KmcOracle *o, int a, int b, int j0, int j1)
{
    a = wrapi(a, o->d0); b = wrapi(b, o->d1);
    uint8_t now[NS];
    eval_site(o->status, o->d0, o->d1, a, b, now);
    uint8_t *old = o->slot + ((size_t)a * o->d1 + b) * NS;
    for (int j = j0; j < j1; j++) {
        if (old[j] == now[j]) continue;
        if (old[j] != NONE) { o->count[old[j]]--; o->rowcnt[a * NC + old[j]]--; }
        if (now[j] != NONE) { o->count[now[j]]++; o->rowcnt[a * NC + now[j]]++; }
        old[j] = now[j];
    }
}

/* Neighbourhood refresh after the status of `nodes` changed: every rule's moves_dependent_on. */
static void refresh_after(KmcOracle *o, const int (*nodes)[2], int n_nodes)
{
    for (int i = 0; i < n_nodes; i++) {
        int a = nodes[i][0], b = nodes[i][1];
        /* single-site rules: Create/FillDivacancy, Create/Fill/FlipMonovacancy (rules.py:74,91,220,292)
         * and DestroyTrefoil slots of a trefoil anchored here */
        refresh_slots(o, a, b, 0, 7);
        /* MoveDivacancy: all nodes within distance <= 1 (rules.py:115-117) */
        refresh_slots(o, a, b, 7, 13);
        for (int k = 0; k < 6; k++) refresh_slots(o, a + RING[k][0], b + RING[k][1], 7, 13);
        /* MoveMonovacancy: likewise all nodes within distance <= 1 (move_monovacancy.py:moves_dependent_on) */
        refresh_slots(o, a, b, MM_SLOT0, NS);
        for (int k = 0; k < 6; k++) refresh_slots(o, a + RING[k][0], b + RING[k][1], MM_SLOT0, NS);
        /* Create/DestroyTrefoil: every triangle containing this node (rules.py:159-174,192-198):
         * anchors p, p-(2,0), p-(0,2) [Up] and p, p-(2,0), p-(2,-2) [Down] */
        refresh_slots(o, a, b, 13, 17);
        refresh_slots(o, a - 2, b, 13, 17);
        refresh_slots(o, a, b - 2, 13, 17);
        refresh_slots(o, a - 2, b + 2, 13, 17);
    }
}

/* Rule.perform for the move in (site, slot); returns the affected nodes (nodes_affected_by). */
static int perform_move(KmcOracle *o, int site, int slot, int nodes[3][2]);
/* + tracers: a site that BECOMES a divacancy by anything but a hop starts a new tracer at displacement 0 */
static int perform(KmcOracle *o, int site, int slot, int nodes[3][2])
{
    int nn = perform_move(o, site, slot, nodes);
    int carried = nn < 0;
    if (carried) nn = -nn;
    if (o->trc && !(slot >= 7 && slot <= 12) && !carried)
        for (int i = 0; i < nn; i++) {
            int j = nodes[i][0] * o->d1 + nodes[i][1];
            if (o->status[j] == DIVAC) { o->trc[2 * j] = 0; o->trc[2 * j + 1] = 0; }
        }
    return nn;
}
static int perform_move(KmcOracle *o, int site, int slot, int nodes[3][2])
{
    int d0 = o->d0, d1 = o->d1;
    int a = site / d1, b = site % d1;
    uint8_t *s = o->status;
    nodes[0][0] = a; nodes[0][1] = b;
    if (slot == 0) { s[site] = DIVAC; return 1; }                  /* new_divacancy   (rules.py:62)  */
    if (slot == 1) { s[site] = PRISTINE; return 1; }               /* pop_divacancy   (rules.py:79)  */
    if (slot == 2 || slot == 3) { s[site] |= (uint8_t)(slot - 1); return 1; }   /* (rules.py:204-210) */
    if (slot == 4 || slot == 5) { s[site] &= (uint8_t)~(slot - 3); return 1; }
    if (slot == 6) { s[site] ^= 3; return 1; }                     /* flip layer      (rules.py:279)  */
    if (slot >= 7 && slot <= 12) {                                 /* MoveDivacancy   (rules.py:96-99) */
        int k = slot - 7;
        int na = wrapi(a + NBR[k][0], d0), nb = wrapi(b + NBR[k][1], d1);
        s[site] = PRISTINE; s[na * d1 + nb] = DIVAC;
        nodes[1][0] = na; nodes[1][1] = nb;
        if (o->trc) {                                              /* the hop carries the tracer along */
            o->trc[2 * (na * d1 + nb)] = o->trc[2 * site] + NBR[k][0];
            o->trc[2 * (na * d1 + nb) + 1] = o->trc[2 * site + 1] + NBR[k][1];
        }
        return 2;
    }
    if (slot >= MM_SLOT0) {                                        /* MoveMonovacancy (move_monovacancy.py:perform) */
        int k = (slot - MM_SLOT0) >> 1, layer = ((slot - MM_SLOT0) & 1) + 1;
        int na = wrapi(a + NBR[k][0], d0), nb = wrapi(b + NBR[k][1], d1);
        int was_div = s[site] == DIVAC;
        s[site] &= (uint8_t)~layer; s[na * d1 + nb] |= (uint8_t)layer;
        nodes[1][0] = na; nodes[1][1] = nb;
        /* tracers: a divacancy that trades places with a monovacancy (swap) carries its displacement along;
         * a merge starts a new tracer (handled by the caller: the target became a divacancy) */
        if (o->trc && was_div && s[na * d1 + nb] == DIVAC) {
            o->trc[2 * (na * d1 + nb)] = o->trc[2 * site] + NBR[k][0];
            o->trc[2 * (na * d1 + nb) + 1] = o->trc[2 * site + 1] + NBR[k][1];
            return -2;                                             /* 2 nodes, tracer already carried */
        }
        return 2;
    }
    int o_ = (slot - 13) & 1, create = slot < 15;                  /* trefoils (rules.py:143-146,178-181) */
    for (int r = 0; r < 3; r++) {
        int na = wrapi(a + TMEM[o_][r][0], d0), nb = wrapi(b + TMEM[o_][r][1], d1);
        s[na * d1 + nb] = create ? (uint8_t)(TREF + 3 * o_ + r) : DIVAC;
        nodes[r][0] = na; nodes[r][1] = nb;
    }
    return 3;
}

/* math.fsum: exactly rounded sum (Shewchuk partials + half-even fix-up), used by sim.py:138. */
static double exact_sum(const double *v, int n)
{
    double p[2 * NC + 2];
    int np = 0;
    for (int k = 0; k < n; k++) {
        double x = v[k];
        int i = 0;
        for (int j = 0; j < np; j++) {
            double y = p[j];
            if (fabs(x) < fabs(y)) { double t = x; x = y; y = t; }
            volatile double hi = x + y;
            volatile double lo = y - (hi - x);
            if (lo != 0.0) p[i++] = lo;
            x = hi;
        }
        np = i; p[np++] = x;
    }
    double hi = 0.0;
    if (np > 0) {
        int m = np;
        hi = p[--m];
        double lo = 0.0;
        while (m > 0) {
            double x = hi, y = p[--m];
            volatile double h2 = x + y;
            volatile double yr = h2 - x;
            hi = h2; lo = y - yr;
            if (lo != 0.0) break;
        }
        if (m > 0 && ((lo < 0.0 && p[m - 1] < 0.0) || (lo > 0.0 && p[m - 1] > 0.0))) {
            double y = lo * 2.0;
            volatile double x = hi + y;
            volatile double yr = x - hi;
            if (y == yr) hi = x;
        }
    }
    return hi;
}

/* One KMC event under injected draws (sim.py:106-143).  Returns 0, or 1 if the total weight is
 * zero (the reference raises ValueError, kmc.py:31-32; the lattice is left untouched). */
int kmco_step(KmcOracle *o, double u1, double u2, KmcoEvent *ev)
{
    double w[NC]; int cls[NC]; int n = 0;
    double allw[NC];
    /* k_w_pairs in canonical order; weight = count * rate (sim.py:120-121); drop zeros (kmc.py:30) */
    for (int i = 0; i < o->n_order; i++) {
        int c = o->order[i];
        double wi = (double)o->count[c] * o->rate[c];
        allw[i] = wi;
        if (wi != 0.0) { w[n] = wi; cls[n] = c; n++; }
    }
    if (ev) {
        for (int c = 0; c < NC; c++) ev->counts[c] = (uint32_t)o->count[c];
        ev->pad = 0; ev->flags = 0;
        ev->total_rate = exact_sum(allw, o->n_order);
    }
    if (n == 0) {
        if (ev) { ev->cls = NONE; ev->slot = NONE; ev->site = 0xFFFFFFFFu; ev->total_naive = 0.0; }
        return 1;
    }
    /* stable sort ascending by weight (kmc.py:35) -- insertion sort is stable */
    for (int i = 1; i < n; i++) {
        double wi = w[i]; int ci = cls[i]; int j = i - 1;
        while (j >= 0 && w[j] > wi) { w[j + 1] = w[j]; cls[j + 1] = cls[j]; j--; }
        w[j + 1] = wi; cls[j + 1] = ci;
    }
    /* sequential partial sums starting from integer 0 (util.py:65-68) */
    double cum[NC] = {0}; double acc = 0.0;
    for (int i = 0; i < n; i++) { acc = acc + w[i]; cum[i] = acc; }
    double total = cum[n - 1];
    double x = 0.0 + (total - 0.0) * u1;               /* random.uniform(0., total) (kmc.py:49) */
    int sel = 0;                                        /* bisect_left (kmc.py:50) */
    while (sel < n && cum[sel] < x) sel++;
    if (sel >= n) sel = n - 1;                          /* unreachable: x <= total */
    int tie = 0;
    for (int i = 0; i < n; i++) tie |= (cum[i] == x);
    int c = cls[sel];

    /* in-class pick: canonical rank order (row, slot, column) */
    int64_t cnt = o->count[c];
    int64_t r = (int64_t)(u2 * (double)cnt);
    if (r > cnt - 1) r = cnt - 1;
    int row = 0;
    while (r >= o->rowcnt[row * NC + c]) { r -= o->rowcnt[row * NC + c]; row++; }
    int site = -1, slot = -1;
    for (int j = 0; j < NS && site < 0; j++)
        for (int b = 0; b < o->d1; b++)
            if (o->slot[((size_t)row * o->d1 + b) * NS + j] == c) {
                if (r == 0) { site = row * o->d1 + b; slot = j; break; }
                r--;
            }
    if (ev) {
        ev->site = (uint32_t)site; ev->cls = (uint8_t)c; ev->slot = (uint8_t)slot;
        ev->flags = (uint16_t)tie; ev->total_naive = total;
    }
    o->n_ties += tie;
    o->hist[c]++;
    if (o->clock_mode == 1) o->clock = o->clock + 1.0 / total;
    else if (o->clock_mode == 2)
        o->clock = o->clock + (0.0 - kmco_log(draw3(o->clk_seed, o->clk_replica, o->clk_step))) / total;
    o->steps_done++;

    int nodes[3][2];
    int nn = perform(o, site, slot, nodes);
    refresh_after(o, (const int (*)[2])nodes, nn);
    return 0;
}

/* Same pick, but by a linear scan of the whole slot table (no row hierarchy): cross-check only. */
int kmco_rank_to_move_scan(const KmcOracle *o, int c, int64_t r, int *site, int *slot)
{
    for (int a = 0; a < o->d0; a++)
        for (int j = 0; j < NS; j++)
            for (int b = 0; b < o->d1; b++) {
                const int i = a * o->d1 + b;
                if (o->slot[(size_t)i * NS + j] == c) { if (r == 0) { *site = i; *slot = j; return 0; } r--; }
            }
    return 1;
}

/* Apply an explicit move (KmcSim.perform_given_move, sim.py:145-157). */
int kmco_perform_given(KmcOracle *o, int site, int slot)
{
    if (site < 0 || site >= o->n || slot < 0 || slot >= NS) return -1;
    if (o->slot[(size_t)site * NS + slot] == NONE) return -2;
    int nodes[3][2];
    int nn = perform(o, site, slot, nodes);
    refresh_after(o, (const int (*)[2])nodes, nn);
    return 0;
}

/* KmcSim.validate (sim.py:159-168): incremental tables == from-scratch enumeration. */
int kmco_validate(const KmcOracle *o)
{
    uint8_t *slots = (uint8_t *)malloc((size_t)o->n * NS);
    int32_t *rowcnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->d0 * NC);
    int64_t counts[NC];
    kmco_sweep(o->status, o->d0, o->d1, slots, counts, rowcnt);
    int bad = 0;
    if (memcmp(slots, o->slot, (size_t)o->n * NS)) bad |= 1;
    if (memcmp(counts, o->count, sizeof counts)) bad |= 2;
    if (memcmp(rowcnt, o->rowcnt, sizeof(int32_t) * (size_t)o->d0 * NC)) bad |= 4;
    /* trefoil integrity: every member code has its two partners */
    for (int i = 0; i < o->n && !(bad & 8); i++) {
        int st = o->status[i];
        if (st > 9) { bad |= 8; break; }
        if (st >= TREF) {
            int o_ = (st - TREF) / 3, r = (st - TREF) % 3;
            int a = i / o->d1 - TMEM[o_][r][0], b = i % o->d1 - TMEM[o_][r][1];
            for (int q = 0; q < 3; q++)
                if (o->status[at(o->d0, o->d1, a + TMEM[o_][q][0], b + TMEM[o_][q][1])] != TREF + 3 * o_ + q)
                    bad |= 8;
        }
    }
    free(slots); free(rowcnt);
    return bad;
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. SC'11).  Draw recipe: see oracle/philox.py. */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void kmco_draw(uint64_t seed, uint32_t replica, uint64_t step, double *u1, double *u2)
{
    uint32_t c[4] = {(uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(seed >> 32), 0u};
    philox4x32_10(c, (uint32_t)seed, replica);
    *u1 = (double)((((uint64_t)c[0] << 32) | c[1]) >> 11) * 0x1.0p-53;
    *u2 = (double)((((uint64_t)c[2] << 32) | c[3]) >> 11) * 0x1.0p-53;
}

/* Run `nsteps` events from a buffer of draws u[nsteps][2]; returns the number of events done
 * (stops early when the total weight is zero). */
int64_t kmco_run_draws(KmcOracle *o, const double *u, int64_t nsteps, KmcoEvent *ev)
{
    for (int64_t t = 0; t < nsteps; t++) {
        o->clk_step = (uint64_t)o->steps_done;       /* injected draws: u3 from (seed, replica) set by kmco_set_clock_key */
        if (kmco_step(o, u[2 * t], u[2 * t + 1], ev ? ev + t : NULL)) return t;
    }
    return nsteps;
}

int64_t kmco_run_philox(KmcOracle *o, uint64_t seed, uint32_t replica, uint64_t step0, int64_t nsteps,
                        KmcoEvent *ev)
{
    for (int64_t t = 0; t < nsteps; t++) {
        double u1, u2;
        kmco_draw(seed, replica, step0 + (uint64_t)t, &u1, &u2);
        o->clk_seed = seed; o->clk_replica = replica; o->clk_step = step0 + (uint64_t)t;
        if (kmco_step(o, u1, u2, ev ? ev + t : NULL)) return t;
    }
    return nsteps;
}

/* Batch of independent replicas (one trajectory each, as the reference runs them one per process).
 * status: [R][n] in/out; hist: [NC] summed over replicas; returns total events done.
 * Threads: pthreads pulling replica indices from a shared counter. */
typedef struct {
    int d0, d1, n_order, n_replicas;
    const double *rate; const int *order;
    uint8_t *status; uint32_t replica0; uint64_t seed, step0; int64_t nsteps;
    int next;                       /* guarded by mu */
    int64_t total, hist[NC];        /* guarded by mu */
    pthread_mutex_t mu;
} BatchJob;

static void *batch_worker(void *arg)
{
    BatchJob *jb = (BatchJob *)arg;
    size_t n = (size_t)jb->d0 * jb->d1;
    KmcOracle *o = kmco_create(jb->d0, jb->d1, jb->rate, jb->order, jb->n_order);
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int r = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (r >= jb->n_replicas) break;
        kmco_reset_stats(o);
        kmco_set_status(o, jb->status + (size_t)r * n);
        int64_t done = kmco_run_philox(o, jb->seed, jb->replica0 + (uint32_t)r, jb->step0, jb->nsteps, NULL);
        kmco_get_status(o, jb->status + (size_t)r * n);
        pthread_mutex_lock(&jb->mu);
        jb->total += done;
        for (int c = 0; c < NC; c++) jb->hist[c] += o->hist[c];
        pthread_mutex_unlock(&jb->mu);
    }
    kmco_destroy(o);
    return NULL;
}

int64_t kmco_run_replicas(int d0, int d1, const double *rate, const int *order, int n_order,
                          uint8_t *status, int n_replicas, uint32_t replica0, uint64_t seed,
                          uint64_t step0, int64_t nsteps, int64_t *hist, int n_threads)
{
    BatchJob jb;
    memset(&jb, 0, sizeof jb);
    jb.d0 = d0; jb.d1 = d1; jb.rate = rate; jb.order = order; jb.n_order = n_order;
    jb.status = status; jb.n_replicas = n_replicas; jb.replica0 = replica0;
    jb.seed = seed; jb.step0 = step0; jb.nsteps = nsteps;
    pthread_mutex_init(&jb.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &jb);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&jb.mu);
    if (hist) for (int c = 0; c < NC; c++) hist[c] = jb.hist[c];
    return jb.total;
}

int kmco_event_size(void) { return (int)sizeof(KmcoEvent); }
