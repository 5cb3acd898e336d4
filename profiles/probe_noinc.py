import sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
big = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
big.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
big.run_noinc(5, bench.SEED); big.sync()
big.timer_start(); big.run_noinc(100, bench.SEED); ms = big.timer_stop()
print("4096^2 no-incremental event: %.1f us" % (ms * 1e3 / 100))
