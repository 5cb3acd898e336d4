"""One launch of every production kernel at its benchmark shape, for a consolidated `ncu --set full` capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))
import numpy as np
import bench
from demo1.engine import Engine
rates, order = bench.wse2_rates(), bench.WSE2_ORDER
# K1+K2 sweep with slot planes, and the count-only sweep + selection of the no-incremental path (4096^2)
sw = Engine(bench.SWEEP_DIM, rates, order, n_replicas=1)
sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
sw.sweep(); sw.sweep()
sw.run_noinc(2, bench.SEED)
sw.close()
# K3+K4: half-warp kernel (16384 replicas), warp-per-replica kernel (one trajectory), wide kernel (1024^2 x 128)
eng = Engine(bench.DIM, rates, order, n_replicas=16384)
eng.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, bench.SEED)
eng.run(2000, bench.SEED); eng.run(2000, bench.SEED); eng.sync(); eng.close()
one = Engine(bench.DIM, rates, order, n_replicas=1)
one.upload_state(bench.synthetic_lattices(1, 0)); one.run(20000, bench.SEED); one.run(20000, bench.SEED); one.sync(); one.close()
big = Engine((1024, 1024), rates, order, n_replicas=128)
big.generate_state("approx", {"divacancy": 0.05, "monovacancy": 0.05}, bench.SEED)
big.run(500, bench.SEED); big.run(1000, bench.SEED); big.sync()
big.zobrist(64, 1); big.download_state(); big.close()
print("done")
