"""Warp-per-replica kernel in the strong-scaling regime (2048 replicas per GPU): warps per block."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
E = 5000
for R in (1024, 2048, 3072):
    lat = bench.synthetic_lattices(R, 0)
    for lat_build in ("1", "0"):
        os.environ["KMC_BP_LAT"] = lat_build
        row = []
        for W in (3, 4, 5, 6, 7, 8):
            eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
            eng.set_replica_kernel(2); eng.set_warps_per_block(W)
            eng.upload_state(lat); eng.run(1000, bench.SEED); eng.sync()
            ts = []
            for _ in range(3):
                eng.upload_state(lat); eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
            eng.close()
            row.append("W=%d %.3e" % (W, R * E / (min(ts) * 1e-3)))
        print("R %5d lat=%s: %s" % (R, lat_build, "  ".join(row)), flush=True)
