"""Strong-scaling batch sizes with the full rule set: warp-per-replica (2) vs half-warp (4) kernel."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from demo1.engine import Engine
from util_cases import WSE2_FULL_ORDER, wse2_full_rates
rates, order = wse2_full_rates(1500.0), WSE2_FULL_ORDER
for R in (1024, 2048, 3072, 4096, 8192):
    out = []
    for variant, env in ((2, {}), (4, {}), (4, {"KMC_HW_REPLICAS_PER_BLOCK": "14"}), (4, {"KMC_HW_REPLICAS_PER_BLOCK": "28"})):
        for k, v in env.items(): os.environ[k] = v
        eng = Engine(bench.DIM, rates, order, n_replicas=R)
        eng.set_replica_kernel(variant)
        eng.upload_state(bench.synthetic_lattices(R, 0))
        eng.run(500, bench.SEED); eng.sync()
        eng.timer_start(); eng.run(5000, bench.SEED); ms = eng.timer_stop()
        eng.check_status(); eng.close()
        for k in env: os.environ.pop(k, None)
        out.append("v%d%s %.3e" % (variant, (" hpb=" + env["KMC_HW_REPLICAS_PER_BLOCK"]) if env else "", R * 5000 / (ms * 1e-3)))
    print("R %5d: %s" % (R, "   ".join(out)), flush=True)
