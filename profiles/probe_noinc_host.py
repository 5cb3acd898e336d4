"""Is the no-incremental loop bound by the host's launch rate?  Enqueue time vs device time for N events."""
import os, sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
N = int(os.environ.get("NOINC_EVENTS", "2000"))
big = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
big.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
big.run_noinc(5, bench.SEED); big.sync()
big.timer_start()
t0 = time.perf_counter(); big.run_noinc(N, bench.SEED); t1 = time.perf_counter()
ms = big.timer_stop(); t2 = time.perf_counter()
print("events %d: enqueue %.2f us/event, device %.2f us/event, wall %.2f us/event" % (N, 1e6 * (t1 - t0) / N, 1e3 * ms / N, 1e6 * (t2 - t0) / N))
big.close()
