import sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
for dim, R, mode in (((1024, 1024), 128, "approx"), ((64, 64), 16384, "approx"), ((64, 64), 16384, "exact"), ((1024, 1024), 8, "exact")):
    eng = Engine(dim, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
    eng.generate_state(mode, {"divacancy": 0.05, "monovacancy": 0.05}, 1)
    t0 = time.perf_counter(); eng.generate_state(mode, {"divacancy": 0.05, "monovacancy": 0.05}, 1); dt = time.perf_counter() - t0
    print("%s x %d %s: %.1f ms" % (dim, R, mode, dt * 1e3), flush=True)
    eng.close()
