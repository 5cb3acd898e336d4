/*
 * kmc_b200.h -- C ABI of the B200-native KMC engine for the kmc-dichalcogen hot path.
 *
 * The reference (ExpHP/kmc-dichalcogen) is pure Python and has NO plugin / FFI interface: the seam
 * this library replaces is the Python object `KmcSim` as driven by `demo1/__main__.py:run`
 * (reference demo1/__main__.py:123,129,141,148) and the callbacks it fans out to.  Every entry
 * point below names the reference interface it stands in for (paths relative to the reference
 * root).  The Python host in kmc-dichalcogen_b200/demo1/ binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns 0 on success and a
 *     negative kmc_status on failure; kmc_last_error() gives the message (thread-local).
 *   - host buffers belong to the caller; device memory belongs to the handle.
 *   - one host thread per handle; work is ordered on the handle's own CUDA stream.
 *   - lattice: axial hex coordinates (a, b), 0 <= a < dim0, 0 <= b < dim1, periodic; site index
 *     i = a*dim1 + b (reference demo1/state.py:388-390,461-464).  dim0, dim1 >= 5 and != 6
 *     (smaller lattices alias the trefoil geometry under PBC; 6 makes collinear "triangles").
 *   - site status byte: 0 pristine, 1 / 2 monovacancy in layer 1 / 2, 3 divacancy (these equal the
 *     reference's vacant layer set, demo1/state.py:22-23,171-177), 4 + 3*orientation + role for a
 *     trefoil member (orientation 0: {p, p+(2,0), p+(0,2)}, 1: {p, p+(2,0), p+(2,-2)}; role = index
 *     in that list; role 0 is the anchor p).
 *   - event classes (rule:kind) 0..16 and per-site move slots 0..28: see the tables at the end.  Classes 0..12 /
 *     slots 0..16 are the rules the reference implements; classes 13..16 / slots 17..28 are MoveMonovacancy, which
 *     the reference's config/WSe2.yaml:8-13 names but its code only stubs (demo1/rules.py:295-296) -- an extension
 *     with stated semantics, pinned by fixtures made from the reference engine running that
 *     rule as a MultiKindRule subclass (oracle/move_monovacancy.py).
 */
#ifndef KMC_B200_H
#define KMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMC_N_CLASSES 17
#define KMC_N_SLOTS   29
#define KMC_SLOT_NONE 0xFF
#define KMC_CLASS_NONE 0xFF      /* event.cls when the total rate is zero (reference kmc.py:31-32) */

typedef enum {
    KMC_OK = 0,
    KMC_ERR_ARG = -1,            /* bad argument (RuntimeError('config: ...') territory)              */
    KMC_ERR_CUDA = -2,           /* CUDA runtime failure (message carries cudaGetErrorString)         */
    KMC_ERR_STATE = -3,          /* invalid lattice (bad status byte / broken trefoil)                */
    KMC_ERR_VALIDATE = -4,       /* incremental tables != from-scratch sweep (AssertionError)         */
    KMC_ERR_NO_DEVICE = -5,      /* no usable CUDA device: there is NO CPU fallback                   */
    KMC_ERR_ZERO_RATE = -6       /* some replica reached total rate 0 (ValueError, kmc.py:32)         */
} kmc_status;

/* Event log records.  total_rate is the sequential fp64 prefix-sum total the selection used
 * (reference kmc.py:41-42); the reference *logs* math.fsum of the same weights (sim.py:138), which
 * the host recomputes exactly from `counts` (full records) -- the two agree to <= 1e-12 relative. */
typedef struct {
    uint32_t site;               /* anchor site of the move                                          */
    uint8_t  cls;                /* event class 0..16, or KMC_CLASS_NONE                             */
    uint8_t  slot;               /* move slot 0..28                                                  */
    uint16_t flags;              /* bit0: the class draw landed exactly on a cumulative boundary     */
    double   total_rate;
} kmc_event_compact;             /* 16 bytes */

typedef struct {
    uint32_t site;
    uint8_t  cls;
    uint8_t  slot;
    uint16_t flags;
    double   total_rate;
    uint32_t counts[KMC_N_CLASSES];   /* moves per class BEFORE the event (sim.py:90-104)          */
    uint32_t pad;
} kmc_event_full;                /* 88 bytes */

enum { KMC_LOG_NONE = 0, KMC_LOG_COMPACT = 1, KMC_LOG_FULL = 2 };

typedef struct kmc_engine kmc_engine;

const char *kmc_last_error(void);
int kmc_device_count(void);

/* KmcSim.__init__ (reference demo1/sim.py:39-44) for `n_replicas` independent trajectories on
 * CUDA device `device`.  class_rate[c] is the fp64 rate of class c (the host computes Arrhenius
 * rates with math.exp exactly as sim.py:179-181 so bits match); classes of disabled rules get 0.
 * class_order lists the enabled class ids in weighted_choice input order (rules in config order,
 * kinds in Rule.kinds() order): it is the tie-break among equal weights (kmc.py:35 is a stable
 * sort).  The lattice starts pristine. */
int kmc_create(kmc_engine **out, int device, int dim0, int dim1, int n_replicas,
               const double *class_rate /*[17]*/, const int *class_order, int n_order);
int kmc_destroy(kmc_engine *h);

/* KmcSim.initialize(state) (sim.py:46-54): replace the lattices.  host: [n_replicas][dim0*dim1]. */
int kmc_upload_state(kmc_engine *h, const uint8_t *host_status);
int kmc_download_state(kmc_engine *h, uint8_t *host_status);
/* gen_random_state (state.py:329-378) on the device, one Philox stream per replica (key = (seed, replica0 + r),
 * counter domain 1; see csrc/stategen_kernel.cuh for the draw recipe).  The caller passes the plan the
 * reference derives on the host: mode 0 'exact' -- n_keys runs of the shuffled node list, run k = positions
 * [bounds[k], bounds[k+1]) gets codes[k]; mode 1 'approx' -- the sorted live (key, weight) pairs of
 * weighted_choice as cumulative weights cumul[n_keys].  codes: 3 divacancy, 1 monovacancy (layer drawn per
 * site), 0 leave pristine.  n_keys <= 3.  Replaces all lattices of the handle. */
int kmc_generate_state(kmc_engine *h, int mode, int n_keys, const int *codes, const int64_t *bounds,
                       const double *cumul, uint64_t seed, uint32_t replica0);
/* same, but the source/destination is a device pointer on the handle's device */
int kmc_set_state_device(kmc_engine *h, const void *dev_status);
int kmc_get_state_device(kmc_engine *h, void *dev_status);

/* Full enumeration = Rule.initialize_moves for all rules (sim.py:204-206, rules.py:19-21,43-45), the
 * `--no-incremental` sweep (sim.py:114-115): status planes -> 17 slot planes (class id or 0xFF per
 * site), per-row and per-replica class counts.  Everything stays on the device. */
int kmc_sweep(kmc_engine *h);
/* IncrementalMoveCache.undecided_counts for every rule (incremental.py:103-110): [n_replicas][17],
 * from the last kmc_sweep(). */
int kmc_get_counts(kmc_engine *h, int64_t *out);
/* IncrementalMoveCache.undecided_sets (incremental.py:112-123) as a table: [n_replicas][n][29]
 * class id or 0xFF, from the last kmc_sweep().  Parity / debug. */
int kmc_get_slots(kmc_engine *h, uint8_t *out);
/* per-row class counts of the last kmc_sweep(): [n_replicas][dim0][17] */
int kmc_get_row_counts(kmc_engine *h, uint32_t *out);

/* KmcSim.perform_random_move x nsteps (sim.py:106-143) for every replica, incremental tables held
 * in shared memory, many events per launch.  Draws: Philox4x32-10, key = (seed lo32, replica0 + r),
 * counter = (step lo32, step hi32, seed hi32, 0); u1 = class draw, u2 = in-class draw (53-bit).
 * Step numbering continues across calls.  events: host buffer [n_replicas][nsteps] of
 * kmc_event_compact / kmc_event_full according to log_mode, or NULL with KMC_LOG_NONE.
 * validate != 0 re-enumerates every lattice at the end of the launch and compares with the
 * incrementally maintained tables (KmcSim.validate, sim.py:159-168) -> KMC_ERR_VALIDATE.
 * A replica whose total rate reaches zero stops (reference: ValueError, kmc.py:31-32; see kmc_get_steps_done).
 * With an event log or validate != 0 the call synchronises and returns KMC_ERR_ZERO_RATE / KMC_ERR_VALIDATE
 * itself.  A statistics-only launch (KMC_LOG_NONE, validate == 0) is ASYNCHRONOUS and returns KMC_OK once
 * queued: the stop is then reported by the next kmc_check_status() on the handle (it stays raised until
 * kmc_reset_stats or until the lattice is replaced).  The batch kernel (two replicas per warp) takes at most
 * 2^30 events per launch (KMC_ERR_ARG beyond that): split longer runs, the step numbering continues. */
int kmc_run(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, int log_mode,
            void *events, int validate);
/* Same, with the caller's uniform draws u[n_replicas][nsteps][2] (host buffer): the seam the
 * reference offers through weighted_choice(..., rng=) (kmc.py:21) and get_random_key(..., rng=)
 * (incremental.py:283). */
int kmc_step_draws(kmc_engine *h, const double *u, int64_t nsteps, int log_mode, void *events,
                   int validate);
/* Wait for the handle's stream, then report what the launches since the last kmc_reset_stats / lattice
 * replacement raised: KMC_ERR_ZERO_RATE if some replica stopped on a total rate of zero (kmc.py:31-32),
 * KMC_ERR_VALIDATE if a validating launch found a mismatch, else KMC_OK.  The synchronous counterpart of an
 * asynchronous kmc_run. */
int kmc_check_status(kmc_engine *h);
/* KmcSim.perform_given_move (sim.py:145-157) on one replica: apply move (site, slot). */
int kmc_perform_given(kmc_engine *h, int replica, uint32_t site, int slot);

/* The `--no-incremental` event (sim.py:114-143) on a lattice of any size held in HBM: full sweep,
 * class selection + hierarchical in-class rank search, apply.  Same draws and records as kmc_run.  At most 2^30
 * events per call (KMC_ERR_ARG beyond that).  Lattices with rows of at most 16384 sites run on bit planes, one kernel
 * per event -- one cooperative launch for the whole call when all its CTAs can be resident; the status bytes are
 * brought up to date when they are asked for. */
int kmc_run_noinc(kmc_engine *h, int64_t nsteps, uint64_t seed, uint32_t replica0, int log_mode,
                  void *events);
/* the same event with the caller's draws u[n_replicas][nsteps][2] (the kmc_step_draws seam on this path) */
int kmc_step_draws_noinc(kmc_engine *h, const double *u, int64_t nsteps, int log_mode, void *events);

/* statistics: events per class summed over replicas [17]; per replica [n_replicas][17];
 * events done per replica [n_replicas]; boundary ties summed over replicas. */
int kmc_get_histogram(kmc_engine *h, int64_t *out);
int kmc_get_histogram_per_replica(kmc_engine *h, int64_t *out);
int kmc_get_steps_done(kmc_engine *h, int64_t *out);
/* Divacancy hop statistics (extension, SURVEY 8(f) rank 4 "diffusion observables"): per replica, the number of
 * MoveDivacancy events per direction k = 0..5 (slot 7 + k; hop vector = neighbour k of demo1/state.py:392-394
 * as listed in `neighbors_and_mutuals`, state.py:443-459) since the last kmc_reset_stats.  The net
 * displacement of the divacancy population is sum_k hops[k] * nbr_k.  out: [n_replicas][6]. */
int kmc_get_hops(kmc_engine *h, int64_t *out);
/* Zobrist key of every replica's current lattice, from scratch (State.zobrist_key / __validate_zobrist_key,
 * demo1/state.py:136-141,306-317; ZobristKey, demo1/incremental.py:335-380): XOR over the entities on the
 * lattice of a `bits`-wide value per (site, status code) from a fixed Philox table keyed by `seed`
 * (csrc/zobrist_kernel.cuh; the reference draws its values on demand from `random`, so only the structure
 * -- equal keys iff equal lattices -- carries over).  out: [n_replicas]. */
int kmc_zobrist(kmc_engine *h, int bits, uint64_t seed, uint64_t *out);
int kmc_get_ties(kmc_engine *h, int64_t *out);
int kmc_reset_stats(kmc_engine *h);
/* Optional KMC clock (extension; the reference draws no time step, SURVEY F6, but logs total_rate so that time
 * can be reconstructed, sim.py:138, README.md:36-37).  Per replica, accumulated in fp64 in event order across
 * launches:  mode 1: t += 1 / total_rate (the mean residence time);  mode 2: the BKL / Gillespie residence time
 * t += -ln(u3) / total_rate with a third Philox draw per event, u3 in (0, 1] from counter domain 3 of the
 * replica's stream (csrc/kmc_common.cuh: kmc_draw3, kmc_log -- a fixed sequence of IEEE operations, so the value
 * is bit-identical on the device, the host and the oracle).  mode 0 switches it off.  out: [n_replicas]. */
int kmc_enable_clock(kmc_engine *h, int mode);
int kmc_get_clock(kmc_engine *h, double *out);
/* Optional divacancy tracers (extension, SURVEY 8(f) rank 4 "diffusion observables"): every divacancy carries its
 * unwrapped displacement (axial da, db; int16 each) since it appeared (CreateDivacancy, CreateMonovacancy
 * make-empty, DestroyTrefoil), since the lattice was replaced, or since kmc_reset_stats; a MoveDivacancy hop in
 * direction k adds neighbour k of neighbors_and_mutuals (state.py:443-459).  kmc_get_msd: per replica the sum
 * over its divacancies of da^2 + da*db + db^2 (squared length in lattice constants) and their number, so that
 * MSD = sum / count and, with the clock, D = MSD / (4 t).  [n_replicas] each.  kmc_get_tracers: the raw table
 * of one replica, [dim0*dim1] words da | db << 16 (entries of sites without a divacancy are stale). */
int kmc_enable_tracers(kmc_engine *h, int on);
int kmc_get_msd(kmc_engine *h, int64_t *sum_sq, int64_t *n_tracers);
int kmc_get_tracers(kmc_engine *h, int replica, uint32_t *out);

/* measurement helpers: CUDA events recorded on the handle's stream */
int kmc_sync(kmc_engine *h);
/* read `bytes` (0 = 256 MiB, > the 126 MB L2) of scratch on the handle's stream, asynchronously: leaves
 * L2 full of clean foreign lines, so the next launch starts cold without a host round trip and without
 * having to write back dirty flush data inside its timed region */
int kmc_flush_l2(kmc_engine *h, size_t bytes);
/* measurement helper: writes the sweep's 17 slot planes with the sweep's store pattern and nothing else
 * (no loads, no arithmetic): the practical ceiling of the sweep's write traffic on this device */
int kmc_probe_store_pattern(kmc_engine *h);
int kmc_timer_start(kmc_engine *h);
int kmc_timer_stop(kmc_engine *h, float *elapsed_ms);   /* synchronises */
/* number of kernels this library has launched on the handle so far */
int kmc_get_launch_count(kmc_engine *h, int64_t *out);
/* raw device pointers (for callers that share device memory, e.g. torch tensors / NCCL) */
int kmc_get_device_pointers(kmc_engine *h, void **status, void **slots, void **row_counts,
                            void **histogram);
/* tuning knobs: warps (replicas) per CTA of the replica kernel, 0 = automatic (max 8); and which
 * replica kernel runs: 0 = automatic (bit-sliced lattice, two replicas per warp, when dim1 <= 64; byte
 * lattice otherwise), 1 = byte lattice in shared memory, 2 = bit-sliced with one warp per replica,
 * 4 = bit-sliced with a half-warp per replica, 3 = byte lattice left in HBM/L2 with only the
 * row-count hierarchy in shared memory (chosen automatically when the lattice does not fit and variant 5
 * does not apply), 5 = bit planes of rows wider than 64 sites kept in HBM/L2 with the row table in shared
 * memory (64 < dim1 <= 2048; chosen automatically when the byte lattices of the batch cannot all be resident
 * in shared memory, e.g. 1024x1024 or a thousand 256x256 replicas; a whole CTA per replica when the batch has no
 * more replicas than the device has SMs, else one warp per replica).  All produce identical results. */
int kmc_set_warps_per_block(kmc_engine *h, int warps);
int kmc_set_replica_kernel(kmc_engine *h, int variant);
/* aligned-shape sweep kernel: 0 = automatic (rolling window over rows when a row fills a CTA, i.e. 4096
 * sites; else nibble-parallel + PRMT tables), 1 = nibble-parallel + PRMT, 2 = byte-parallel (with MoveMonovacancy
 * enabled: the same kernel as 1; the byte-parallel kernel has no build for that rule).  Same results. */
int kmc_set_sweep_kernel(kmc_engine *h, int variant);
/* measurement helper: successive kmc_sweep calls write two alternating slot-plane buffers (steady-state
 * bandwidth: no launch rewrites lines its predecessor left dirty in L2).  Getters read the latest one. */
int kmc_set_sweep_double_buffer(kmc_engine *h, int on);

/*
 * Event classes (reference demo1/rules.py):           Move slots per site (move key in the reference):
 *   0  CreateDivacancy:natural      (:60-75)            0      CreateDivacancy        node
 *   1  FillDivacancy:natural        (:77-92)            1      FillDivacancy          node
 *   2  MoveDivacancy:natural        (:94-136)           2,3    CreateMonovacancy      (node, layer 1/2, old|layer)
 *   3  MoveDivacancy:missing-metal                      4,5    FillMonovacancy        (node, layer 1/2, old&~layer)
 *   4  MoveDivacancy:missing-comb                       6      FlipMonovacancy        node
 *   5  MoveDivacancy:missing-both                       7..12  MoveDivacancy          (node, nbr_k), k = slot-7, in
 *   6  CreateTrefoil:natural        (:141-174)                                        neighbors_and_mutuals order
 *   7  DestroyTrefoil:natural       (:176-198)                                        (state.py:443-459)
 *   8  CreateMonovacancy:from-double(:238-255)          13,14  CreateTrefoil          {p,p+(2,0),p+(0,2)} / {p,p+(2,0),p+(2,-2)}
 *   9  CreateMonovacancy:make-empty                     15,16  DestroyTrefoil         same two triangles
 *   10 FillMonovacancy:make-double  (:257-274)
 *   11 FillMonovacancy:from-empty
 *   12 FlipMonovacancy:natural      (:277-293)
 *   13 MoveMonovacancy:natural          (stub :295-296; kinds from    17..28 MoveMonovacancy  (node, nbr_k, layer L): 17 + 2k + (L-1)
 *   14 MoveMonovacancy:merge-divacancy   config/WSe2.yaml:8-13)              the vacancy in layer L of `node` hops to nbr_k, whose
 *   15 MoveMonovacancy:split-divacancy                                       layer set must be defined and lack L;
 *   16 MoveMonovacancy:swap-divacancy                                        kind = 2*(node is a divacancy) + (nbr_k not pristine)
 * While any of classes 13..16 is enabled the engine runs the MoveMonovacancy-aware build of the same kernels (17 class
 * weights, 29 slot planes); rule sets without the rule run the 13-class builds.  The one-warp-per-replica form of
 * variant 5 has no such build (the CTA-per-replica form serves those batches).
 * In-class order (the rank the in-class draw indexes): row ascending, then slot ascending, then column.
 */

#ifdef __cplusplus
}
#endif
#endif /* KMC_B200_H */
