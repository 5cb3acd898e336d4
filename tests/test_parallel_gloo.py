"""The N>1 host path on CPU: world_size-2 gloo processes, replicas sharded by global index, the one
all-reduce of the event histogram.  The per-rank engine here is the CPU oracle (as a stand-in for a
GPU); what is under test is the sharding and reduction logic of demo1/parallel.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from demo1 import parallel
from util_cases import ko, WSE2_ORDER, wse2_rates, exact_state

N_TOTAL, NSTEPS, SEED, DIM = 7, 300, 11, (16, 16)


def test_replica_ranges_partition_the_batch():
    for n in (1, 7, 16, 16384, 1000):
        for w in (1, 2, 3, 4, 8):
            spans = [parallel.replica_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _lattices():
    return np.stack([exact_state(DIM, 0.08, 0.05, 50 + r) for r in range(N_TOTAL)])


def _worker(rank, world_size, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, size, _ = parallel.init(backend="gloo")
    lo, hi = parallel.replica_range(N_TOTAL, size, r)
    st = _lattices()[lo:hi]
    done, hist, st = ko.run_replicas(DIM, wse2_rates(4000.0), WSE2_ORDER, st, seed=SEED, nsteps=NSTEPS,
                                     replica0=lo, threads=1)
    total = parallel.reduce_histogram(hist)
    slowest = parallel.max_over_ranks(float(rank + 1))
    both = parallel.sum_over_ranks(0.5 + rank)
    # uneven shards (4 + 3 replicas) gathered in rank order: what the batch CLI sums exactly with fsum
    gathered = parallel.gather_float64(np.arange(lo, hi, dtype=np.float64) * 0.1)
    parallel.barrier()
    np.savez(os.path.join(out_dir, "rank%d.npz" % r), hist=total, slowest=slowest, both=both, st=st, lo=lo, hi=hi,
             gathered=gathered)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
def test_two_ranks_reduce_to_the_single_process_result(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    single_st = _lattices()
    done, single_hist, single_st = ko.run_replicas(DIM, wse2_rates(4000.0), WSE2_ORDER, single_st, seed=SEED,
                                                   nsteps=NSTEPS, replica0=0, threads=2)
    assert done == N_TOTAL * NSTEPS
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    for p in parts:
        assert (p["hist"] == single_hist).all()          # every rank holds the global histogram
        assert float(p["slowest"]) == 2.0                # max over ranks
        assert float(p["both"]) == 2.0                   # sum over ranks: 0.5 + 1.5
        assert (p["gathered"] == np.arange(N_TOTAL, dtype=np.float64) * 0.1).all()
        assert (p["st"] == single_st[int(p["lo"]):int(p["hi"])]).all()   # results independent of sharding
    assert [int(p["lo"]) for p in parts] == [0, 4] and [int(p["hi"]) for p in parts] == [4, 7]
