"""BASELINE config 5 as stated: 1024x1024 lattices, 1024 replicas over the GPUs of one box (128 per GPU at 8 GPUs),
WSe2 rates at 1500 K, lattices generated on the device, hop statistics reduced with NCCL at the end.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/prof_config5_multi.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))
import numpy as np
import torch
import bench
from demo1 import parallel
from demo1.engine import Engine

rank, world, local = parallel.world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
parallel.init_quiet(backend="nccl" if world > 1 else None, device=dev)
TOTAL, EVENTS = 1024, 20000
lo, hi = parallel.replica_range(TOTAL, world, rank)
eng = Engine((1024, 1024), bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=hi - lo, device=local)
eng.generate_state("approx", {"divacancy": 0.05, "monovacancy": 0.05}, bench.SEED, replica0=lo)
eng.enable_clock(True)
eng.run(500, bench.SEED, replica0=lo); eng.sync()
eng.reset_stats()
parallel.barrier(); torch.cuda.synchronize()
eng.timer_start(); eng.run(EVENTS, bench.SEED, replica0=lo); ms = eng.timer_stop()
ms = parallel.max_over_ranks(ms, device=dev)
hist = parallel.reduce_histogram(eng.histogram(), device=dev)
hops = parallel.reduce_histogram(eng.hops().sum(axis=0), device=dev)
d = eng.displacement()
sq = parallel.sum_over_ranks(float((d[:, 0] ** 2 + d[:, 0] * d[:, 1] + d[:, 1] ** 2).sum()), device=dev)
tsum = parallel.sum_over_ranks(float(eng.clock().sum()), device=dev)
if rank == 0:
    print(json.dumps({"workload": "1024x1024 WSe2 1500 K, %d replicas over %d GPUs (%d per GPU), %d events per replica"
                      % (TOTAL, world, hi - lo, EVENTS), "n_gpus": world, "events_per_s": TOTAL * EVENTS / (ms * 1e-3),
                      "ms": ms, "event_histogram": [int(x) for x in hist], "divacancy_hops": [int(x) for x in hops],
                      "mean_squared_net_displacement": sq / TOTAL, "mean_time": tsum / TOTAL}))
eng.close()
parallel.shutdown()
