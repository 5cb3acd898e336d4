"""Small profiling target for ncu: 3 full sweeps of a 4096x4096 lattice and 3 launches of the fused
replica kernel (16384 replicas x 64x64, `--events` events each).  Run plain first, then under ncu."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))
import bench  # noqa: E402
from demo1.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=500)
ap.add_argument("--replicas", type=int, default=16384)
ap.add_argument("--what", default="both", choices=["both", "sweep", "replica"])
ap.add_argument("--rules", default="wse2", choices=["wse2", "full", "mixed"],
                help="wse2: config/WSe2.yaml minus MoveMonovacancy (the headline); full: every rule; mixed: 13 live classes")
ap.add_argument("--variant", type=int, default=0)
args = ap.parse_args()
rates, order = bench.wse2_rates(), bench.WSE2_ORDER
if args.rules == "full":
    import math
    bar = dict(bench.WSE2_BARRIERS)
    bar.update({13: 2.56175687, 14: 3.03662851, 15: 2.33673358, 16: 2.56175687})
    rates = [(math.exp(-bar[c] / (bench.TEMPERATURE * bench.KB_EV)) if c in bar else 0.0) for c in range(17)]
    order = [8, 9, 12, 13, 14, 15, 16, 2, 3, 4, 5, 6, 7]
elif args.rules == "mixed":
    rates = [10.0, 20.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 10.0, 300.0, 100.0, 100.0, 55.0]
    order = [8, 9, 10, 11, 0, 2, 3, 4, 5, 6, 7, 12, 1]
if args.what in ("both", "sweep"):
    sw = Engine(bench.SWEEP_DIM, rates, order, n_replicas=1)
    sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
    for _ in range(3):
        sw.sweep()
    sw.sync()
    print("sweep counts", sw.counts()[0].tolist())
    sw.close()
if args.what in ("both", "replica"):
    eng = Engine(bench.DIM, rates, order, n_replicas=args.replicas)
    eng.set_replica_kernel(args.variant)
    eng.upload_state(bench.synthetic_lattices(args.replicas, 0))
    for _ in range(3):
        eng.run(args.events, bench.SEED)
    eng.sync()
    print("replica hist", eng.histogram().tolist())
    eng.close()
