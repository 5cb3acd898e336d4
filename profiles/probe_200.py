"""BASELINE configs[0] (200x200, pristine, test-cfg rates, one trajectory): byte lattice in shared memory vs bit planes."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench
from demo1.engine import Engine
from util_cases import TEST_RATES, ALL_ORDER
for dim in ((128, 128), (160, 160), (192, 192), (200, 200), (70, 200), (200, 70), (224, 224), (256, 256), (512, 512)):
    out = []
    for variant, env in ((0, {}), (1, {}), (5, {"KMC_WIDE_CTA": "0"}), (5, {"KMC_WIDE_CTA": "1"})):
        for k, v in env.items(): os.environ[k] = v
        try:
            eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=1)
            eng.set_replica_kernel(variant)
            eng.upload_state(np.zeros((1, dim[0] * dim[1]), np.uint8))
            eng.run(20000, 5); eng.sync()
            eng.timer_start(); eng.run(100000, 5); ms = eng.timer_stop()
            eng.close()
            out.append("v%d%s %.3e" % (variant, (" cta=" + env["KMC_WIDE_CTA"]) if env else "", 100000 / (ms * 1e-3)))
        except Exception as e:
            out.append("v%d: %s" % (variant, str(e)[:40]))
        for k in env: os.environ.pop(k, None)
    print(dim, "   ".join(out), flush=True)
