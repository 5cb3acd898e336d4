"""Small batches on the warp-per-replica kernel: KMC_BP_LAT = 0 / 1 forces the 64- / 96-register build (unset =
automatic: latency build up to 2368 replicas), warps per block 0 = automatic (one block per SM first)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
for R in (1, 148, 1184, 2368):
    for lat in ("0", "1", None):
        if lat is None: os.environ.pop("KMC_BP_LAT", None)
        else: os.environ["KMC_BP_LAT"] = lat
        eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
        eng.set_replica_kernel(2)
        eng.upload_state(bench.synthetic_lattices(R, 0)); eng.run(2000, 1); eng.sync()
        eng.timer_start(); eng.run(50000, 1); ms = eng.timer_stop()
        print("R %4d lat %s: %.3f us/event, %.3e events/s" % (R, lat or "auto", ms * 1e3 / 50000, R * 50000 / ms * 1e3), flush=True)
        eng.close()
