"""Throughput of the incremental path on 1024x1024 x 128 replicas (BASELINE config 5): the wide bit-plane
kernel (variant 5) against the byte lattice kept in HBM (variant 3).  Usage: python profiles/probe_wide.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import importlib
import numpy as np
from util_cases import TEST_RATES, ALL_ORDER, wse2_rates, WSE2_ORDER, iid_state
Engine = importlib.import_module("kmc-dichalcogen_b200.demo1.engine").Engine

for dim, nrep in (((1024, 1024), 128), ((1024, 1024), 1024), ((128, 128), 4096), ((256, 256), 2048)):
    st = iid_state(dim, 3)
    st0 = np.broadcast_to(st, (nrep, st.size)).copy()
    for variant in (3, 5):
        for wpb in (0, 1, 2, 4, 8):
            if variant == 3 and wpb not in (0,): continue
            eng = Engine(dim, wse2_rates(1500.0), WSE2_ORDER, n_replicas=nrep)
            try:
                eng.set_replica_kernel(variant)
                eng.set_warps_per_block(wpb)
                eng.upload_state(st0)
                eng.run(200, 1)
                eng.sync()
                n = 2000
                eng.timer_start(); eng.run(n, 1); ms = eng.timer_stop()
                print("dim=%s R=%d variant=%d wpb=%d: %.3f ms for %d events/replica -> %.3e events/s, %.2f us/event"
                      % (dim, nrep, variant, wpb, ms, n, nrep * n / ms * 1e3, ms * 1e3 / n), flush=True)
            except Exception as e:
                print("dim=%s R=%d variant=%d wpb=%d: %s" % (dim, nrep, variant, wpb, e), flush=True)
            eng.close()
