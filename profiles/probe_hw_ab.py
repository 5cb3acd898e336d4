"""A/B of a library build (KMC_LIB=...) on the three 64x64 x 16384 workloads of the bench (half-warp kernel)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
from util_cases import WSE2_FULL_ORDER, wse2_full_rates
mixed = [10.0, 20.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 10.0, 300.0, 100.0, 100.0, 55.0]
mixed_order = [8, 9, 10, 11, 0, 2, 3, 4, 5, 6, 7, 12, 1]
for name, rates, order in (("full rules", wse2_full_rates(1500.0), WSE2_FULL_ORDER), ("round-1 rules", bench.wse2_rates(), bench.WSE2_ORDER),
                           ("mixed rates", mixed, mixed_order)):
    eng = Engine(bench.DIM, rates, order, n_replicas=16384)
    eng.upload_state(bench.synthetic_lattices(16384, 0))
    eng.run(1000, bench.SEED); eng.sync()
    best = 0.0
    for _ in range(3):
        eng.timer_start(); eng.run(5000, bench.SEED); ms = eng.timer_stop()
        best = max(best, 16384 * 5000 / (ms * 1e-3))
    eng.check_status(); eng.close()
    print("%-14s %.4e events/s" % (name, best), flush=True)
