"""The no-incremental event at 4096^2 with every rule of config/WSe2.yaml (MoveMonovacancy live)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
sys.path.insert(0, "tests")
from util_cases import WSE2_FULL_ORDER, wse2_full_rates
N = int(os.environ.get("NOINC_EVENTS", "1000"))
big = Engine(bench.SWEEP_DIM, wse2_full_rates(1500.0), WSE2_FULL_ORDER, n_replicas=1)
big.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
big.run_noinc(5, bench.SEED); big.sync()
big.timer_start(); big.run_noinc(N, bench.SEED); ms = big.timer_stop()
print("full rules: %.2f us per event" % (1e3 * ms / N))
big.close()
