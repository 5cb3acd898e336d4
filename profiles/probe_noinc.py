"""One --no-incremental event on 4096^2 (BASELINE configs[3] end to end): byte-lattice path vs the plane path."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
for planes in ("0", "1"):
    os.environ["KMC_NOINC_PLANES"] = planes
    big = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
    big.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
    big.run_noinc(5, bench.SEED); big.sync()
    big.timer_start(); big.run_noinc(int(os.environ.get("NOINC_EVENTS", "100")), bench.SEED); ms = big.timer_stop()
    print("planes=%s: %.2f us per event" % (planes, 1e3 * ms / int(os.environ.get("NOINC_EVENTS", "100"))), flush=True)
    big.close()
