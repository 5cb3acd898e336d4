"""`python3 -m demo1 CONFIG...` -- the reference's command line (reference demo1/__main__.py:14-157)
over the CUDA engine: same flags, same YAML, same streamed JSON.  Extra flags: --seed, --replicas,
--stats-only, --device."""
import argparse
import json
import sys

from . import config
from .sim import KmcSim, class_table

PROG = "demo1"


def main(argv=None):
    import logging
    parser = argparse.ArgumentParser("python -m " + PROG, description="")
    parser.add_argument("CONFIG", nargs="+", type=argparse.FileType("r"), help="paths to config files")
    parser.add_argument("-d", "--dimensions", metavar="ARM,ZAG", type=delim_parser(positive_int, n=2, sep=","),
                        default=[200, 200], help="PBC grid dimensions as a number of unit cells in each dimension")
    parser.add_argument("-n", "--steps", type=positive_int, default=10000, help="stop after this many events")
    parser.add_argument("-P", "--output-pstats", help="record profiling data, readable by the pstats module")
    parser.add_argument("-p", "--profile", action="store_true", help="display profiling data")
    parser.add_argument("-D", "--debug", action="store_true", help="display debug logs, hide standard output")
    parser.add_argument("-T", "--temperature", type=float, default=None,
                        help="temperature in kelvin. Only meaningful if config files specify barriers instead of rates.")
    parser.add_argument("--write-initial", metavar="PATH", help="write initial state to PATH")
    parser.add_argument("--embed-initial", action="store_true", help="embed initial state in output")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--no-incremental", dest="incremental", action="store_false",
                       help="disable incremental updates (debugging flag)")
    group.add_argument("--validate-every", metavar="NSTEP", type=nonnegative_int, default=0,
                       help="do incremental updates, but perform an expensive validation of all objects "
                            "every NSTEP steps. (debugging flag)")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--zobrist-bits", metavar="NBITS", type=positive_int, default=None,
                       help="Experimental flag. Computes and prints zobrist keys with this many bits (1..64). "
                            "Values come from a fixed Philox table keyed by --seed, so keys are reproducible.")
    group.add_argument("--zobrist-hash", action="store_true",
                       help="Experimental flag. Deterministic zobrist keys: 64 bits from the table keyed by 0.")
    # new flags
    parser.add_argument("--seed", type=int, default=None, help="Philox seed (default: from the OS)")
    parser.add_argument("--state-seed", type=int, default=None, help="seed of the initial-state generator")
    parser.add_argument("--replicas", type=positive_int, default=1,
                        help="run this many independent trajectories (statistics only unless 1)")
    parser.add_argument("--stats-only", action="store_true", help="no event log, only the class histogram")
    parser.add_argument("--independent-states", action="store_true",
                        help="with --replicas: every replica starts from its own lattice, generated on the GPU "
                             "from the config's state.random section (Philox stream (state seed, replica))")
    parser.add_argument("--device", type=int, default=0, help="CUDA device index")
    args = parser.parse_args(argv)

    if args.debug:
        logging.getLogger().setLevel(10)
    if args.zobrist_bits is not None and args.zobrist_bits > 64:
        die("--zobrist-bits: at most 64")
    config_dict = config.load_all(args.CONFIG)

    def myrun():
        run(ofile=DevNull() if args.debug else sys.stdout, nsteps=args.steps, dims=args.dimensions,
            config_dict=config_dict, validate_every=args.validate_every, incremental=args.incremental,
            save_initial_path=args.write_initial, embed_initial=args.embed_initial,
            temperature=args.temperature, seed=args.seed, state_seed=args.state_seed,
            replicas=args.replicas, stats_only=args.stats_only, device=args.device,
            independent_states=args.independent_states,
            zobrist=(64, 0) if args.zobrist_hash else ((args.zobrist_bits, None) if args.zobrist_bits else None))

    if args.output_pstats or args.profile:
        with_profiling(myrun, args.output_pstats, sys.stderr if args.profile else None)
    else:
        myrun()


class DevNull:
    def write(self, *a, **kw):
        pass


def run(ofile, nsteps, dims, config_dict, validate_every, incremental, save_initial_path, embed_initial,
        temperature, seed=None, state_seed=None, replicas=1, stats_only=False, device=0,
        independent_states=False, zobrist=None):
    import random
    cfg = config.from_dict(config_dict)        # pops every key: "config" is echoed as {} (reference quirk Q1)
    rng = random.Random(state_seed) if state_seed is not None else random
    init_state = cfg["state_gen_func"](tuple(dims), rng=rng)
    if save_initial_path:
        with open(save_initial_path, "w") as f:
            json.dump(init_state.to_dict(), f)

    if replicas > 1 or stats_only:
        return run_batch(ofile, nsteps, dims, cfg, init_state, temperature, seed, replicas, device,
                         state_seed, independent_states)
    if independent_states:
        die("--independent-states needs --replicas N or --stats-only")

    chunk = validate_every if validate_every else min(nsteps, 4096)
    sim = KmcSim(init_state, cfg["rule_specs"], incremental=incremental, temperature=temperature,
                 seed=seed, device=device, chunk=chunk, validate_launches=bool(validate_every),
                 zobrist_bits=zobrist[0] if zobrist else None, zobrist_seed=zobrist[1] if zobrist else None)

    with write_enclosing("{", "\n}", ofile):
        def write_key_val(key, val, end=",\n", pretty=True, **kw):
            ofile.write('"%s": ' % key)
            json.dump(val, ofile, indent=2 if pretty else None, **kw)
            ofile.write(end)

        write_key_val("grid", grid_info(dims))
        write_key_val("config", config_dict)
        write_key_val("temperature", temperature, pretty=False)
        write_key_val("rates", sim.rates_info(), sort_keys=True)
        if embed_initial:
            write_key_val("initial_state", init_state.to_dict(), pretty=False)
        with write_enclosing(' "events": [\n  ', "\n ]", ofile):
            first = True
            # with --validate-every N the kernel validates at the end of every N-event launch
            # (KmcSim.validate semantics); a mismatch raises AssertionError from the iterator
            for lines in sim.iter_json_lines(nsteps):
                if lines:
                    ofile.write(("" if first else ",\n  ") + ",\n  ".join(lines))
                    first = False
    ofile.write("\n")
    sim.close()


def run_batch(ofile, nsteps, dims, cfg, init_state, temperature, seed, replicas, device, state_seed=None,
              independent_states=False, zobrist=None):
    """--replicas / --stats-only: many independent trajectories; output is the per-class event histogram and
    the divacancy hop statistics instead of an event log.  Under torchrun (one process per GPU) the replicas are
    sharded by contiguous global index; replica r's trajectory depends only on (seed, r), so the result does
    not depend on the number of GPUs.  Rank 0 writes the output."""
    import os
    import numpy as np
    from . import parallel
    from . import tables as T
    from .engine import Engine
    rules_and_rates, rates13, order = class_table(cfg["rule_specs"], temperature)
    info = {"%s:%s" % (rule, kind): rate for rule, rates in rules_and_rates.items() for kind, rate in rates.items()}
    rank, size, local = parallel.world()
    if size > 1:
        if seed is None or (state_seed is None and not independent_states):
            die("--seed (and --state-seed, unless --independent-states) are required when several processes "
                "share a batch")
        import torch
        device = local
        torch.cuda.set_device(local)
        parallel.init_quiet(device=torch.device("cuda", local))
    seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
    lo, hi = parallel.replica_range(replicas, size, rank)
    if hi == lo:
        die("more processes than replicas")
    eng = Engine(init_state.dim, rates13, order, n_replicas=hi - lo, device=device)
    if independent_states:
        gen = cfg["state_gen_func"].keywords          # partial(gen_random_state, mode=..., params=...)
        eng.generate_state(gen["mode"], gen["params"], seed if state_seed is None else state_seed, replica0=lo)
    else:
        eng.upload_state(np.tile(init_state.status, (hi - lo, 1)))
    # simulated time per replica: the BKL clock t += -ln(u3) / total_rate (a third Philox draw per event; the
    # reference draws no time step but logs total_rate for this reconstruction, README.md:36-37), and one tracer
    # per divacancy carrying its unwrapped displacement
    eng.enable_clock("stochastic")
    eng.enable_tracers(True)
    try:
        eng.run(nsteps, seed, replica0=lo, validate=True)
    except ValueError:
        pass
    dev = None
    if size > 1:
        import torch
        dev = torch.device("cuda", local)
    sums = np.concatenate([eng.histogram(), eng.hops().sum(axis=0), [int(eng.steps_done().sum())]]).astype(np.int64)
    sums = parallel.reduce_histogram(sums, device=dev)
    # sum of squared net displacements as exact integers (a^2 + a*b + b^2 per replica)
    d = eng.displacement()
    sq = parallel.reduce_histogram(np.array([int((d[:, 0] ** 2 + d[:, 0] * d[:, 1] + d[:, 1] ** 2).sum())],
                                            dtype=np.int64), device=dev)
    # per-replica times gathered and summed exactly (fsum), so the result does not depend on the sharding
    import math
    time_sum = math.fsum(parallel.gather_float64(eng.clock(), device=dev).tolist())
    tsq, tcnt = eng.msd()
    tr = parallel.reduce_histogram(np.array([int(tsq.sum()), int(tcnt.sum())], dtype=np.int64), device=dev)
    eng.close()
    parallel.shutdown()
    if rank != 0:
        return
    nc = T.N_CLASSES
    hist, hops, done = sums[:nc], sums[nc:nc + 6], sums[nc + 6]
    out = {"grid": grid_info(dims), "temperature": temperature, "rates": info, "replicas": replicas,
           "seed": seed, "events_done": int(done), "processes": size,
           "histogram": {"%s:%s" % T.CLASSES[c]: int(hist[c]) for c in order},
           # divacancy hops per direction (neighbour order of state.py:443-459), summed over replicas, and the
           # mean squared net displacement of a replica's divacancy population in axial lattice units
           "divacancy_hops": [int(x) for x in hops],
           "mean_squared_net_displacement": float(sq[0]) / replicas,
           # simulated time per replica, averaged: sum over events of -ln(u3) / total_rate
           "mean_time": time_sum / replicas,
           # tracer diffusion: mean squared displacement (lattice constants^2) of the divacancies present at the
           # end, each measured from where it appeared (or from the initial lattice), and D = MSD / (4 t)
           "tracers": int(tr[1]),
           "msd": (float(tr[0]) / float(tr[1])) if tr[1] else None,
           "D": (float(tr[0]) / float(tr[1]) / (4.0 * time_sum / replicas)) if tr[1] and time_sum > 0 else None}
    json.dump(out, ofile, indent=2, sort_keys=True)
    ofile.write("\n")


def axial_norm2(d):
    """|a * e1 + b * e2|^2 for axial hex coordinates (unit lattice constant): a^2 + a*b + b^2."""
    import numpy as np
    d = np.asarray(d, dtype=np.float64)
    return d[:, 0] ** 2 + d[:, 0] * d[:, 1] + d[:, 1] ** 2


class write_enclosing:
    """Writes `start` on entry and `end` on exit, also on exceptions, so that an interrupted run
    leaves a parseable file (reference __main__.py:159-174)."""

    def __init__(self, start, end, file):
        self.file, self.start, self.end = file, start, end

    def __enter__(self):
        self.file.write(self.start)

    def __exit__(self, *args):
        self.file.write(self.end)


def grid_info(dims):
    return {"lattice_type": "hexagonal", "coord_format": "axial", "dim": list(dims)}


def with_profiling(job, stats_out=None, text_out=None):
    """Run ``job()`` under cProfile (``-P PATH`` dumps pstats data, ``-p`` prints the 20 most expensive calls by
    cumulative time).  ``python -m demo1`` cannot be combined with ``-m cProfile``, hence the wrapper."""
    import cProfile
    import pstats
    profiler = cProfile.Profile()
    try:
        return profiler.runcall(job)
    finally:
        if stats_out:
            try:
                profiler.dump_stats(stats_out)
            except OSError as err:
                warn("could not write pstats ({})", err)
        if text_out:
            pstats.Stats(profiler, stream=text_out).sort_stats("cumtime").print_stats(20)


def _bounded_int(lowest, what):
    def parse(text):
        value = int(text)
        if value < lowest:
            raise argparse.ArgumentTypeError("%s: %d" % (what, value))
        return value
    return parse


positive_int = _bounded_int(1, "Not a positive integer")
nonnegative_int = _bounded_int(0, "Not a non-negative integer")


def delim_parser(typ, n=None, sep=","):
    """argparse type for ``A<sep>B<sep>...``: a list of ``typ`` values, exactly ``n`` of them if given."""
    def parse(text):
        pieces = text.split(sep)
        if n is not None and len(pieces) != n:
            raise argparse.ArgumentTypeError("Expected %d arguments separated by %r; got %d" % (n, sep, len(pieces)))
        try:
            return [typ(piece) for piece in pieces]
        except ValueError:
            raise argparse.ArgumentTypeError("Bad value in %r" % (text,))
    return parse


def say_err(msg, *args, **kw):
    sys.stderr.write("%s: %s\n" % (PROG, msg.format(*args, **kw)))


def warn(msg, *args, **kw):
    say_err("Warning: " + msg, *args, **kw)


def die(msg, *args, **kw):
    say_err("FATAL: " + msg, *args, **kw)
    say_err("Aborting.")
    raise SystemExit(1)


if __name__ == "__main__":
    main()
