#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 KMC engine (contract: see the task's bench section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2]): config/WSe2.yaml barriers at T = 1500 K (stub rule
MoveMonovacancy disabled), 64x64 lattice with 5 % divacancies + 5 % monovacancies, 16384 independent
replicas PER GPU (two replicas per warp; weak scaling), Philox draws keyed by (seed, global replica).
One "step" = one launch of the fused replica kernel = `--events` KMC events on every replica.

One JSON line on stdout (rank 0):
  value      whole-job events/s, lattices resident in HBM, CUDA-event timed on the engine's stream
  e2e        same metric through the C ABI with HOST buffers: per step H2D of all lattices from
             pinned memory, the events, D2H of the final lattices + per-replica event histograms
  roofline   the HBM-bound kernel of the path: the 4096x4096 full rate sweep (BASELINE configs[3]),
             18 algorithmic bytes per site, timed live here with CUDA events, L2 flushed
  replica_kernel   SM-bound figures for the fused kernel (events/s per SM, SM cycles per event)
  cpu_baseline     the CPU oracle port (oracle/kmc_oracle.c) on the host cores, bounded sample
`--impl reference` times that CPU port alone (the reference itself is pure Python that cannot travel
to the GPU box) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))

SEED = 20261019
DIM = (64, 64)
TEMPERATURE = 1500.0
KB_EV = 1.38064852e-23 * 6.24150913e18            # reference demo1/sim.py:27-29
WSE2_BARRIERS = {2: 2.25, 3: 2.11199155, 4: 2.52009312, 5: 2.25, 6: 3.12293677, 7: 4.43051273,
                 8: 6.72584072, 9: 6.72584072, 12: 3.41217502}      # config/WSe2.yaml
WSE2_ORDER = [8, 9, 12, 2, 3, 4, 5, 6, 7]
# config/WSe2.yaml with EVERY rule it names: MoveMonovacancy (:8-13) is a stub in the reference (rules.py:295-296) and was
# left out in round 1; it is built now, so the headline workload is the whole file
WSE2_FULL_BARRIERS = dict(WSE2_BARRIERS, **{"13": 2.56175687, "14": 3.03662851, "15": 2.33673358, "16": 2.56175687})
WSE2_FULL_BARRIERS = {int(k): v for k, v in WSE2_FULL_BARRIERS.items()}
WSE2_FULL_ORDER = [8, 9, 12, 13, 14, 15, 16, 2, 3, 4, 5, 6, 7]
SWEEP_DIM = (4096, 4096)
SWEEP_BYTES_PER_SITE = 18                         # 1 B status read + 17 B slot codes written
SWEEP_BYTES_PER_SITE_FULL = 30                    # + 12 B MoveMonovacancy slot codes


def _rates(barriers):
    import math
    r = [0.0] * 17
    for c, e in barriers.items():
        r[c] = math.exp(-e / (TEMPERATURE * KB_EV))
    return r


def wse2_rates():
    """config/WSe2.yaml without MoveMonovacancy: the round-1 headline workload (kept for continuity and probes)."""
    return _rates(WSE2_BARRIERS)


def wse2_full_rates():
    return _rates(WSE2_FULL_BARRIERS)


def synthetic_lattices(n_replicas, replica0, dim=DIM, frac_div=0.05, frac_mono=0.05):
    """'exact'-mode lattices (reference demo1/state.py:360-378 semantics: exact defect counts at
    shuffled positions), vectorised: one independent permutation per replica."""
    n = dim[0] * dim[1]
    nd, nm = round(frac_div * n), round((frac_div + frac_mono) * n)
    rng = np.random.Generator(np.random.Philox(key=SEED + 7919 * replica0))
    base = np.zeros((n_replicas, n), dtype=np.uint8)
    base[:, :nd] = 3
    base[:, nd:nm] = rng.integers(1, 3, size=(n_replicas, nm - nd), dtype=np.uint8)
    return rng.permuted(base, axis=1)


def iid_lattice(dim, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.choice(np.arange(4, dtype=np.uint8), size=dim[0] * dim[1],
                      p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def recorded_replica_counters():
    """ncu counters of the fused replica kernel from the committed capture (instructions and issue-slot use
    per event: the quantities that bound it, since it moves ~0 HBM bytes)."""
    p = os.path.join(ROOT, "profiles", "replica_counters.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def recorded_traffic():
    """dram bytes read + written per sweep launch, from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# -------------------------------------------------------------------------------------------------
def cpu_port_events_per_sec(n_threads, target_seconds, rates, order):
    """The CPU oracle port on the host cores: one independent trajectory per thread at a time
    (the reference runs one trajectory per process).  bench.py may execute oracle/ only here."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kmc_oracle as ko
    probe_r, probe_e = n_threads, 20000
    st = synthetic_lattices(probe_r, 0)
    t0 = time.perf_counter()
    ko.run_replicas(DIM, rates, order, st, seed=SEED, nsteps=probe_e, threads=n_threads)
    rate = probe_r * probe_e / (time.perf_counter() - t0)
    n_rep = 4 * n_threads
    n_ev = int(max(20000, min(5_000_000, rate * target_seconds / n_rep)))
    st = synthetic_lattices(n_rep, 0)
    t0 = time.perf_counter()
    done, hist, _ = ko.run_replicas(DIM, rates, order, st, seed=SEED, nsteps=n_ev, threads=n_threads)
    dt = time.perf_counter() - t0
    return done / dt, "%d replicas x %d events, %d threads, %.1f s" % (n_rep, n_ev, n_threads, dt), dt


def python_reference_events_per_sec(n_procs, nsteps=6000):
    """The ACTUAL Python reference (unmodified, staged under baseline/_ref/reference by __graft_entry__.build(); the
    py3.12 compat shim oracle/ref_shim on PYTHONPATH), one process per core -- it has no threads -- on the headline
    workload: config/WSe2.yaml with the stub MoveMonovacancy disabled by an overlay file, -T 1500, 64x64, 5 % + 5 %
    seeded defects.  events/s = processes * nsteps / (wall(nsteps) - wall(1 step)): initialisation subtracted."""
    import tempfile
    ref = os.path.join(ROOT, "baseline", "_ref", "reference")
    if not os.path.isdir(os.path.join(ref, "demo1")):
        return None
    tmp = tempfile.mkdtemp(prefix="kmc_ref_")
    with open(os.path.join(tmp, "overlay.yaml"), "w") as f:
        f.write("rules:\n  MoveMonovacancy: {enabled: false}\nstate:\n  random:\n    mode: exact\n"
                "    divacancy: 5%\n    monovacancy: 5%\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "ref_shim"), PYTHONDONTWRITEBYTECODE="1")

    def wall(n):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, "-m", "demo1", "config/WSe2.yaml", os.path.join(tmp, "overlay.yaml"),
                                   "-T", "1500", "-d", "64,64", "-n", str(n)], cwd=ref, env=env,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.PIPE) for _ in range(n_procs)]
        errs = [p.communicate()[1] for p in procs]
        if any(p.returncode != 0 for p in procs):
            raise RuntimeError("reference run failed: %s" % errs[0][-500:].decode("utf-8", "replace"))
        return time.perf_counter() - t0
    try:
        wall(1)                                   # warm the page cache / bytecode-less imports
        t1 = min(wall(1), wall(1))
        tn = wall(nsteps)
    except (RuntimeError, OSError) as e:
        return {"unavailable": str(e)[:300]}
    v = n_procs * (nsteps - 1) / max(tn - t1, 1e-9)
    return {"value": v, "unit": "events/s", "cores": n_procs, "kind": "reference",
            "sample": "%d processes x %d events (python -m demo1 config/WSe2.yaml + overlay, -T 1500 -d 64,64), "
                      "init subtracted; %.1f s" % (n_procs, nsteps, t1 + tn),
            "events_per_s_per_core": v / n_procs}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, as its C port (the Python
    reference cannot travel to the GPU box), on all host threads, same workload/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates, order = wse2_full_rates(), WSE2_FULL_ORDER
    n_threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kmc_oracle as ko
    n_rep = 4 * n_threads
    # calibrate one step to ~ (120 s / (steps + warmup)), capped at 10 s
    st = synthetic_lattices(n_threads, 0)
    t0 = time.perf_counter()
    ko.run_replicas(DIM, rates, order, st, seed=SEED, nsteps=20000, threads=n_threads)
    rate = n_threads * 20000 / (time.perf_counter() - t0)
    per_step_s = min(10.0, 120.0 / (args.steps + args.warmup))
    n_ev = int(max(5000, rate * per_step_s / n_rep))
    st = synthetic_lattices(n_rep, 0)
    for w in range(args.warmup):
        ko.run_replicas(DIM, rates, order, st, seed=SEED, nsteps=n_ev, step0=w * n_ev, threads=n_threads)
    t0 = time.perf_counter()
    total = 0
    for k in range(args.steps):
        done, _, _ = ko.run_replicas(DIM, rates, order, st, seed=SEED, nsteps=n_ev,
                                     step0=(args.warmup + k) * n_ev, threads=n_threads)
        total += done
    dt = time.perf_counter() - t0
    value = total / dt
    sample = "%d replicas x %d events per step (64x64, config/WSe2.yaml all rules, 1500 K), %d host threads" % (n_rep, n_ev, n_threads)
    line = {
        "impl": "reference", "metric": "kmc_events_per_sec", "value": value, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 lattice / f64 rates",
        "data": "synthetic",
        "config": {"workload": "config/WSe2.yaml (every rule, MoveMonovacancy included) 64x64 lattice, independent replicas, "
                               "CPU port of the reference hot path (oracle/kmc_oracle.c), bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=16384, help="replicas per GPU")
    ap.add_argument("--strong-replicas", type=int, default=16384,
                    help="replicas in TOTAL for the strong-scaling block (BASELINE configs[2] as worded)")
    ap.add_argument("--events", type=int, default=10000, help="events per replica per step (launch)")
    ap.add_argument("--warps-per-block", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--e2e-in-flight", type=int, default=2,
                    help="steps kept in flight by the end-to-end measurement (1 = sequential, 2 = copies overlap)")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--sweep-iters", type=int, default=20)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from demo1 import parallel
    from demo1.engine import Engine

    rank, world, local = parallel.world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    cuda = torch.device("cuda", local)
    # NCCL prints its version banner on stdout when the communicator is created; the contract is ONE JSON
    # line on stdout, so stdout points at stderr while the process group comes up
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        parallel.init(backend="nccl", device=cuda)
        parallel.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    rates, order = wse2_full_rates(), WSE2_FULL_ORDER          # the headline: config/WSe2.yaml, every rule
    rates13, order13 = wse2_rates(), WSE2_ORDER                 # round-1 workload (MoveMonovacancy left out)
    E = args.events

    def measure_batch(R, sample_clocks):
        """Device-resident and end-to-end events/s of R replicas per GPU (global replica index = rank * R + r)."""
        replica0 = rank * R                 # global replica index -> results independent of GPU count
        lattices = synthetic_lattices(R, replica0)
        host = torch.from_numpy(lattices).pin_memory()
        host_out = torch.empty_like(host).pin_memory()
        eng = Engine(DIM, rates, order, n_replicas=R, device=local)
        if args.warps_per_block:
            eng.set_warps_per_block(args.warps_per_block)
        eng.upload_state(host.numpy())
        # ---- device-resident throughput --------------------------------------------------------
        for _ in range(args.warmup):
            eng.run(E, SEED, replica0=replica0)
        eng.sync()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = eng.launch_count()
        total_ms = 0.0
        for _ in range(args.steps):
            eng.flush_l2()                  # on the engine's stream: L2 is cold, the GPU never idles
            eng.timer_start()
            eng.run(E, SEED, replica0=replica0)
            total_ms += eng.timer_stop()
        n_launch = eng.launch_count() - launches0
        clk = sampler.stop() if sampler else None
        barrier()
        eng.check_status()                  # a statistics-only launch is asynchronous: this is where a stop shows
        max_ms = parallel.max_over_ranks(total_ms, device=cuda)
        value = world * R * E * args.steps / (max_ms * 1e-3)
        steps_done = eng.steps_done()
        assert int(steps_done.min()) == (args.warmup + args.steps) * E, "a replica stopped early"
        # the one collective of the path: event histogram summed over GPUs (NCCL over NVLink)
        hist = parallel.reduce_histogram(eng.histogram(), device=cuda)
        assert int(hist.sum()) == world * R * E * (args.warmup + args.steps)

        # ---- end to end through the C ABI with host buffers --------------------------------------
        # Every step: H2D of that step's lattices from pinned memory, the launch, D2H of the final lattices and
        # of the per-replica statistics.  Two handles (two streams), each driven by its own host thread, keep two
        # steps in flight so that one step's copies run under the other's kernel (ctypes releases the GIL in the
        # calls); `--e2e-in-flight 1` gives the strictly sequential figure.
        n_flight = max(1, min(2, args.e2e_in_flight))
        engines = [eng] + [Engine(DIM, rates, order, n_replicas=R, device=local) for _ in range(n_flight - 1)]
        if n_flight == 2 and 1024 <= R < 4096:
            # two handles in flight put 2 R replicas on the GPU at once: enough for the two-replicas-per-warp kernel,
            # which one handle alone would not pick below 4096 replicas (measured at R = 2048: 1.58e9 vs 1.28e9 e2e)
            for e in engines:
                e.set_replica_kernel(4)
        hosts_in = [host] + [host.clone().pin_memory() for _ in range(n_flight - 1)]
        hosts_out = [host_out] + [torch.empty_like(host).pin_memory() for _ in range(n_flight - 1)]
        checks = [0] * n_flight

        def e2e_step(i):
            e = engines[i]
            e.upload_state(hosts_in[i].numpy())                # H2D from pinned memory (+ device-side validation)
            e.run(E, SEED, replica0=replica0)
            e.download_state_into(hosts_out[i].numpy())        # D2H final lattices
            checks[i] = int(e.histogram_per_replica().sum())   # D2H statistics

        def e2e_run(nsteps):
            shares = [nsteps // n_flight + (1 if i < nsteps % n_flight else 0) for i in range(n_flight)]

            def work(i):
                for _ in range(shares[i]):
                    e2e_step(i)
            if n_flight == 1:
                work(0)
                return
            ts = [threading.Thread(target=work, args=(i,)) for i in range(n_flight)]
            for th in ts:
                th.start()
            for th in ts:
                th.join()

        for e in engines:
            e.reset_stats()
        e2e_run(2 * n_flight)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert sum(checks) > 0
        for e in engines:
            e.check_status()
        e2e_value = world * R * E * args.steps / parallel.max_over_ranks(e2e_s, device=cuda)
        for e in engines:
            e.close()
        return {"value": value, "max_ms": max_ms, "e2e": e2e_value, "launches": n_launch, "clocks": clk,
                "hist": hist, "h2d": R * DIM[0] * DIM[1], "d2h": R * DIM[0] * DIM[1] + R * 17 * 8,
                "n_flight": n_flight}

    R = args.replicas
    weak = measure_batch(R, True)
    value, max_ms, e2e_value, launches, clocks, hist = (weak[k] for k in ("value", "max_ms", "e2e", "launches",
                                                                          "clocks", "hist"))
    h2d, d2h, n_flight = weak["h2d"], weak["d2h"], weak["n_flight"]
    # BASELINE configs[2] as worded: 16384 replicas IN TOTAL, sharded over the GPUs (strong scaling).  At one GPU
    # that is the headline batch itself.
    total_strong = args.strong_replicas
    if world == 1 and total_strong == R:
        strong = weak
    else:
        strong = measure_batch(max(1, total_strong // world), False)
        launches += strong["launches"]
    strong_block = {"replicas_total": (total_strong // world) * world, "replicas_per_gpu": total_strong // world,
                    "value": strong["value"], "unit": "events/s", "ms_per_step": strong["max_ms"] / args.steps,
                    "e2e": {"value": strong["e2e"], "unit": "events/s", "h2d_bytes_per_step": strong["h2d"],
                            "d2h_bytes_per_step": strong["d2h"]},
                    "scaling": "strong", "events_per_replica_per_step": E}

    # ---- HBM roofline: 4096^2 full rate sweep (single lattice, "replicas only": rank 0 measures it) -----
    roofline = None
    if rank == 0 and not args.no_sweep:
        peak, peak_src = measured_peaks()
        # the 17-plane build of kmc_sweep_roll_kernel (rule sets without MoveMonovacancy: the round-1 kernel) is timed on
        # an engine without the rule, the 29-plane build (every rule of config/WSe2.yaml in one pass) below
        sw = Engine(SWEEP_DIM, rates13, order13, n_replicas=1, device=local)
        sw.upload_state(iid_lattice(SWEEP_DIM, SEED))
        for _ in range(3):
            sw.sweep()
        sw.sync()
        l0 = sw.launch_count()
        ms = []
        for _ in range(args.sweep_iters):
            sw.flush_l2()                   # queued ahead of the events: no launch gap inside the timed region
            sw.timer_start()
            sw.sweep()
            ms.append(sw.timer_stop())
        # steady state: back-to-back sweeps over two alternating slot-plane buffers, ONE event pair, no L2 flush
        # in between -- each launch starts with the previous launch's last ~60 MB still dirty in L2 and pays for
        # their write-back, exactly as it leaves its own behind: DRAM bytes inside the window = algorithmic bytes
        sw.set_sweep_double_buffer(True)
        for _ in range(4):
            sw.sweep()
        sw.sync()
        n_steady = max(20, args.sweep_iters)
        sw.timer_start()
        for _ in range(n_steady):
            sw.sweep()
        steady_ms = sw.timer_stop() / n_steady
        sweep_launches = sw.launch_count() - l0
        sw.close()
        fs = Engine(SWEEP_DIM, rates, order, n_replicas=1, device=local)
        fs.upload_state(iid_lattice(SWEEP_DIM, SEED))
        for _ in range(3):
            fs.sweep()
        fs.sync()
        l0 = fs.launch_count()
        ms_full = []
        for _ in range(args.sweep_iters):
            fs.flush_l2()
            fs.timer_start()
            fs.sweep()
            ms_full.append(fs.timer_stop())
        fs.set_sweep_double_buffer(True)
        for _ in range(4):
            fs.sweep()
        fs.sync()
        fs.timer_start()
        for _ in range(n_steady):
            fs.sweep()
        full_steady_ms = fs.timer_stop() / n_steady
        sweep_launches += fs.launch_count() - l0
        fs.close()
        full_ms = float(np.mean(ms_full))
        avg_ms = float(np.mean(ms))
        nsites = SWEEP_DIM[0] * SWEEP_DIM[1]
        achieved = nsites * SWEEP_BYTES_PER_SITE / (avg_ms * 1e-3) / 1e9
        roofline = {"kernel": "kmc_sweep_roll_kernel (K1+K2, 4096x4096 full rate sweep)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (recorded_traffic() or {}).get("dram_bytes_per_launch"),
                    "traffic_source": (recorded_traffic() or {}).get("source"), "peak_source": peak_src, "bytes_per_site": SWEEP_BYTES_PER_SITE,
                    "ms_per_sweep": avg_ms, "ms_min": float(np.min(ms)), "sites_per_s": nsites / (avg_ms * 1e-3),
                    "ms_per_sweep_steady": steady_ms,
                    "achieved_steady": nsites * SWEEP_BYTES_PER_SITE / (steady_ms * 1e-3) / 1e9,
                    "frac_steady": nsites * SWEEP_BYTES_PER_SITE / (steady_ms * 1e-3) / 1e9 / peak,
                    "steady_note": "%d back-to-back launches alternating two slot-plane buffers, one CUDA event pair, no flush" % n_steady,
                    "full_rule_set": {"what": "config/WSe2.yaml with MoveMonovacancy: 29 slot planes and 17 classes in one launch "
                                              "of the same kernel's MoveMonovacancy build (kmc_sweep_roll_kernel<true, true>)",
                                      "bytes_per_site": SWEEP_BYTES_PER_SITE_FULL, "ms_per_sweep": full_ms,
                                      "achieved": nsites * SWEEP_BYTES_PER_SITE_FULL / (full_ms * 1e-3) / 1e9,
                                      "frac": nsites * SWEEP_BYTES_PER_SITE_FULL / (full_ms * 1e-3) / 1e9 / peak,
                                      "ms_per_sweep_steady": full_steady_ms,
                                      "frac_steady": nsites * SWEEP_BYTES_PER_SITE_FULL / (full_steady_ms * 1e-3) / 1e9 / peak},
                    "note": "timed region = exactly one launch (slot planes + row counts + class totals in one kernel), CUDA events on the engine's stream, L2 flushed before each launch"}
        launches += sweep_launches

    # ---- the other BASELINE configs, briefly (parity-test shapes; reported for context, rank 0 only) ------
    other = None
    if rank == 0 and not args.no_other_configs:
        other = {}
        FULL, R1 = "config/WSe2.yaml, every rule (MoveMonovacancy live)", "config/WSe2.yaml without MoveMonovacancy (round-1 workload)"

        def timed_batch(rts, odr, nrep, warm, nev, dim=DIM, lattices=None):
            e = Engine(dim, rts, odr, n_replicas=nrep, device=local)
            e.upload_state(lattices if lattices is not None else synthetic_lattices(nrep, 0))
            e.run(warm, SEED)
            e.sync()
            e.timer_start()
            e.run(nev, SEED)
            ms_ = e.timer_stop()
            e.check_status()
            h_ = e.histogram()
            e.close()
            return nrep * nev / (ms_ * 1e-3), [int(x) for x in h_]

        # the round-1 headline workload (the same batch without MoveMonovacancy), for continuity with BENCH_r01
        v, hh = timed_batch(rates13, order13, R, 1000, E)
        other["64x64_x%d_wse2_without_move_monovacancy" % R] = {"events_per_s": v, "replicas": R, "events_per_replica": E,
                                                               "config": R1, "kernel": "kmc_replica_hw_kernel<64,64,PLAIN>",
                                                               "event_histogram": hh}
        # the headline batch with every implemented rule of the reference live at comparable rates
        # (config/mixed_rates.yaml: 13 classes; the two halves of a warp mostly choose different classes)
        mx_rates = [10.0, 20.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 10.0, 300.0, 100.0, 100.0, 55.0]
        mx_order = [8, 9, 10, 11, 0, 2, 3, 4, 5, 6, 7, 12, 1]
        v, hh = timed_batch(mx_rates, mx_order, R, 2000, E)
        other["64x64_x%d_mixed_rates" % R] = {"events_per_s": v, "replicas": R, "events_per_replica": E,
                                              "config": "kmc-dichalcogen_b200/config/mixed_rates.yaml (13 live classes)",
                                              "event_histogram": hh}
        # configs[1]: 64x64, ONE trajectory, incremental (latency-bound: one warp on one SM)
        v, _ = timed_batch(rates, order, 1, 20000, 300000)
        v13, _ = timed_batch(rates13, order13, 1, 20000, 500000)
        other["64x64_single_trajectory_incremental"] = {"events_per_s": v, "config": FULL,
                                                        "without_move_monovacancy_events_per_s": v13}
        # the same trajectory through the reference-facing API with the reference's JSON event log produced for
        # every event (KmcSim.iter_json_lines: kernel launches of 4096 events + host formatting), wall clock
        from demo1.sim import KmcSim, RuleSpec
        from demo1.state import State
        specs = [RuleSpec("CreateMonovacancy", {"from-double": 6.72584072, "make-empty": 6.72584072}, True),
                 RuleSpec("FlipMonovacancy", {"natural": 3.41217502}, True),
                 RuleSpec("MoveMonovacancy", {"natural": 2.56175687, "merge-divacancy": 3.03662851,
                                              "split-divacancy": 2.33673358, "swap-divacancy": 2.56175687}, True),
                 RuleSpec("MoveDivacancy", {"natural": 2.25, "missing-metal": 2.11199155, "missing-comb": 2.52009312,
                                            "missing-both": 2.25}, True),
                 RuleSpec("CreateTrefoil", {"natural": 3.12293677}, True),
                 RuleSpec("DestroyTrefoil", {"natural": 4.43051273}, True)]
        sim = KmcSim(State(DIM, synthetic_lattices(1, 0)[0]), specs, temperature=TEMPERATURE, seed=SEED, device=local)
        nlog, nchar = 200000, 0
        t0 = time.perf_counter()
        for lines in sim.iter_json_lines(nlog):
            nchar += sum(map(len, lines))
        dt = time.perf_counter() - t0
        sim.close()
        other["64x64_single_trajectory_json_event_log"] = {"events_per_s": nlog / dt, "events": nlog, "config": FULL,
                                                           "json_bytes_per_event": nchar / nlog}
        # configs[0]: the reference's own CPU-runnable case (`python3 -m demo1 test-cfg.yaml`: 200x200, pristine
        # start, test-cfg.yaml rates, one trajectory); the Python reference does ~2.5e3 events/s on it
        tc_rates = [10.0, 0.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 10.0, 300.0, 100.0, 100.0, 0.0]
        tc_order = [8, 9, 10, 11, 0, 2, 3, 4, 5, 6, 7]
        v, _ = timed_batch(tc_rates, tc_order, 1, 10000, 200000, dim=(200, 200), lattices=np.zeros((1, 200 * 200), dtype=np.uint8))
        other["200x200_testcfg_single_trajectory"] = {"events_per_s": v, "events": 200000}
        # configs[3] end to end: one `--no-incremental` event on 4096^2 = full re-enumeration + select + apply, on the
        # bit-sliced lattice (noinc_planes_kernel.cuh): one kernel per event; without MoveMonovacancy all CTAs fit on
        # the device at once and the whole call is one cooperative launch
        for key, rts, odr in (("4096x4096_no_incremental_event", rates, order),
                              ("4096x4096_no_incremental_event_without_move_monovacancy", rates13, order13)):
            big = Engine(SWEEP_DIM, rts, odr, n_replicas=1, device=local)
            big.upload_state(iid_lattice(SWEEP_DIM, SEED))
            big.run_noinc(5, SEED)
            big.sync()
            nev = 400
            big.timer_start()
            big.run_noinc(nev, SEED)
            msn = big.timer_stop()
            big.close()
            other[key] = {"events_per_s": nev / (msn * 1e-3), "ms_per_event": msn / nev, "events": nev}
        # configs[4] shape: 1024x1024 lattices stay in HBM/L2 as bit planes, row tables in shared memory, one CTA per
        # replica (128 replicas = one GPU's share of the 1024 at 8 GPUs)
        nrep5, ev5 = 128, 4000
        rng = np.random.Generator(np.random.Philox(SEED))
        base = rng.choice(np.arange(4, dtype=np.uint8), size=1024 * 1024, p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
        lat5 = np.tile(base, (nrep5, 1))
        v, _ = timed_batch(rates, order, nrep5, 200, ev5, dim=(1024, 1024), lattices=lat5)
        v13, _ = timed_batch(rates13, order13, nrep5, 200, ev5, dim=(1024, 1024), lattices=lat5)
        del lat5
        other["1024x1024_x128_replicas_incremental"] = {
            "events_per_s": v, "config": FULL, "replicas": nrep5, "events_per_replica": ev5,
            "without_move_monovacancy_events_per_s": v13,
            "note": "bit planes + row tables stay in HBM/L2 between launches (wide bit-plane kernel, CTA per replica)"}
        # the initial lattices generated on the device (reference gen_random_state 'exact' under
        # a Philox stream per replica) instead of uploaded
        gen = Engine(DIM, rates, order, n_replicas=R, device=local)
        gen.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, SEED)
        gen.timer_start()
        gen.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, SEED)
        msg = gen.timer_stop()
        gen.close()
        other["device_state_generation_exact_64x64"] = {"ms": msg, "replicas": R, "lattices_per_s": R / (msg * 1e-3)}

    cpu = cpu_py = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_threads = os.cpu_count() or 1
        v, sample, _ = cpu_port_events_per_sec(n_threads, 12.0, rates, order)
        cpu = {"value": v, "unit": "events/s", "cores": n_threads, "kind": "port", "sample": sample}
        cpu_py = python_reference_events_per_sec(min(n_threads, 32))

    if rank == 0:
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        per_sm = value / world / 148.0
        line = {
            "metric": "kmc_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 lattice / f64 rates", "data": "synthetic",
            "config": {"workload": "config/WSe2.yaml, every rule it names (MoveMonovacancy -- a stub in the reference -- "
                                   "built as an extension with stated semantics), T=1500K, 64x64 lattice, "
                                   "5%% divacancy + 5%% monovacancy, %d independent replicas per GPU, "
                                   "two replicas per warp (half-warp kernel, MoveMonovacancy-aware build); the round-1 "
                                   "workload (same file without that rule) is other_configs.64x64_x%d_wse2_without_move_monovacancy" % (R, R),
                       "replicas_per_gpu": R, "events_per_replica_per_step": E,
                       "l2": "flushed (256 MiB read on the engine's stream, leaves clean lines) before every timed launch", "seed": SEED,
                       "parallelism": "replicas sharded by contiguous global index; one NCCL all-reduce of "
                                      "the event histogram at the end"},
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps_in_flight": n_flight},
            "strong_scaling": strong_block,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "replica_kernel": {"bound": "sm-issue / shared-memory (lattice resident in smem, ~0 HBM bytes per event)",
                               "events_per_s_per_sm": per_sm, "sm_cycles_per_event": sm_clock / per_sm,
                               "hbm_bytes_per_event_algorithmic": 2.0 * DIM[0] * DIM[1] / E,
                               # instruction-issue roofline: warp-instructions per event (ncu, recorded) x measured
                               # events/s against 4 issue slots per SM per cycle at the sampled SM clock
                               "issue_roofline": (lambda c: None if not c else {
                                   "achieved_warp_inst_per_s": c["warp_instructions_per_event"] * value / world,
                                   "peak_warp_inst_per_s": 148 * 4 * sm_clock,
                                   "frac": c["warp_instructions_per_event"] * value / world / (148 * 4 * sm_clock)})(
                                       recorded_replica_counters()),
                               "ncu": recorded_replica_counters()},
            "cpu_baseline": cpu,
            # the unmodified Python reference itself on the same host cores (one process per core), so that the
            # comparison does not hinge on how fast the C port of it is
            "cpu_baseline_python": cpu_py,
            "other_configs": other,
            "event_histogram": [int(x) for x in hist],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
