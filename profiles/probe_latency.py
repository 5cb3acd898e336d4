"""Latency-bound regimes of the warp-per-replica kernel: one trajectory, and 2048 replicas (strong scaling at 8 GPUs)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
for R, E, variant in ((1, 200000, 0), (148, 50000, 0), (2048, 5000, 2), (2048, 5000, 4), (4096, 5000, 0)):
    lat = bench.synthetic_lattices(R, 0)
    eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
    eng.set_replica_kernel(variant)
    eng.upload_state(lat); eng.run(1000, bench.SEED); eng.sync()
    ts = []
    for _ in range(3):
        eng.upload_state(lat); eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
    eng.close()
    print("R %5d variant %d: %.4e events/s  (%.3f us per event per replica)" % (R, variant, R * E / (min(ts) * 1e-3), min(ts) * 1e3 / E), flush=True)
