"""Strong-scaling regime of BASELINE configs[2]: 16384 replicas in total = 2048 .. 8192 per GPU.
Which fused kernel (warp per replica = 2, two replicas per warp = 4) and which block shape."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
E = 5000


def run(R, variant, env):
    for k in ("KMC_HW_REPLICAS_PER_BLOCK", "KMC_BP_LAT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    lat = bench.synthetic_lattices(R, 0)
    eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
    eng.set_replica_kernel(variant)
    eng.upload_state(lat); eng.run(1000, bench.SEED); eng.sync()
    ts = []
    for _ in range(3):
        eng.upload_state(lat); eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
    eng.close()
    return R * E / (min(ts) * 1e-3)


for R in (1024, 2048, 4096, 8192):
    print("R %5d: bp lat=1 %.3e  bp lat=0 %.3e" % (R, run(R, 2, {"KMC_BP_LAT": "1"}), run(R, 2, {"KMC_BP_LAT": "0"})), flush=True)
    for hpb in (8, 14, 16, 24, 28, 32, 56):
        if R / hpb < 100:
            continue
        print("        hw hpb=%2d (%4d blocks) %.3e" % (hpb, (R + hpb - 1) // hpb, run(R, 4, {"KMC_HW_REPLICAS_PER_BLOCK": str(hpb)})), flush=True)
