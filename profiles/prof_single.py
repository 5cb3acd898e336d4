"""Profiling target: ONE 64x64 WSe2 trajectory (latency-bound: one warp on one SM)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))
import bench
from demo1.engine import Engine
ev = 20000
one = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
one.upload_state(bench.synthetic_lattices(1, 0))
one.run(2000, bench.SEED); one.sync()
one.timer_start(); one.run(ev, bench.SEED); ms = one.timer_stop()
print("single trajectory: %.3e events/s, %.3f us per event" % (ev / (ms * 1e-3), ms * 1e3 / ev))
