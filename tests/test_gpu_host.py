"""The reference-facing host objects on a GPU: KmcSim log dictionaries against the golden fixtures,
and the `python -m demo1` command line (same JSON layout as the reference's)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from demo1.sim import KmcSim, RuleSpec
from demo1.state import State
from util_golden import load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kmc-dichalcogen_b200")


def specs_from_meta(meta):
    out = []
    for name in meta["rule_order"]:
        body = meta["rules"][name]
        key = "barrier" if "barrier" in body else "rate"
        val = body[key]
        if not isinstance(val, dict):
            val = {"natural": val}
        out.append(RuleSpec(name, {k: float(v) for k, v in val.items()}, rate_is_barrier=(key == "barrier")))
    return out


@pytest.mark.parametrize("name,incremental", [("allrules_8x8", True), ("wse2_16x16", True),
                                              ("ties_reordered_7x11", True), ("testcfg_5x7", False),
                                              ("wse2_full_16x16", True), ("allrules_mm_8x8", True),
                                              ("allrules_mm_5x7", False)])
def test_kmcsim_log_matches_reference(name, incremental):
    g = load(name)
    m = g["meta"]
    nsteps = m["nsteps"] if incremental else 150
    sim = KmcSim(State(m["dim"], g["status0"]), specs_from_meta(m), temperature=m["temperature"],
                 incremental=incremental, seed=m["seed"], replica=m["replica"], chunk=257,
                 validate_launches=incremental)
    assert sim.rates_info() == {"%s:%s" % (r, k): v for r in m["rule_order"]
                                for k, v in sim.rules_and_rates[r].items()}
    for t in range(nsteps):
        info = sim.perform_random_move()
        want = g["events"][t]
        assert sorted(info) == ["kind", "move", "rate", "rule", "total_rate"]
        assert info["total_rate"] == want["total_rate"]              # exactly rounded fsum, bit-equal
        assert info["rate"] == g["rates"][want["cls"]]
        if t < len(m["infos"]):
            ref = m["infos"][t]
            assert (info["rule"], info["kind"]) == (ref["rule"], ref["kind"])
            if "nodes" in info["move"]:
                assert sorted(map(tuple, info["move"]["nodes"])) == sorted(map(tuple, ref["move"]["nodes"]))
            else:
                assert info["move"] == ref["move"]
    assert sim.validate() is True
    if nsteps == m["nsteps"]:
        # events are generated ahead in chunks of 257: run to the end of the last chunk on the oracle side
        pass
    sim.close()


def test_input_state_is_not_modified_and_zero_rate_raises():
    st = State((5, 5))
    st.new_divacancy((1, 1))
    before = st.status.copy()
    sim = KmcSim(st, [RuleSpec("FillDivacancy", {"natural": 2.0}, False)], seed=1, chunk=8)
    info = sim.perform_random_move()
    assert info == {"rule": "FillDivacancy", "kind": "natural", "move": {"node": [1, 1]}, "rate": 2.0,
                    "total_rate": 2.0}
    with pytest.raises(ValueError, match="total weight of zero"):
        sim.perform_random_move()
    assert (st.status == before).all()
    assert (sim.current_state().status == 0).all()
    sim.close()


def run_cli(args):
    env = dict(os.environ, PYTHONPATH=PKG)
    p = subprocess.run([sys.executable, "-m", "demo1"] + args, capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_cli_json_layout_matches_reference():
    out = run_cli([os.path.join(PKG, "config", "mixed_rates.yaml"), "-d", "8,8", "-n", "200", "--seed", "7",
                   "--embed-initial", "--validate-every", "50"])
    doc = json.loads(out)
    assert list(doc) == ["grid", "config", "temperature", "rates", "initial_state", "events"]
    assert doc["grid"] == {"lattice_type": "hexagonal", "coord_format": "axial", "dim": [8, 8]}
    assert doc["config"] == {} and doc["temperature"] is None
    assert doc["initial_state"] == {"dim": [8, 8], "vacancies": [], "trefoils": []}
    assert len(doc["events"]) == 200
    # pristine 8x8 with these rates: 64*(2*10 + 10) (SURVEY 8c known answer)
    assert doc["events"][0]["total_rate"] == 1920.0
    assert out.startswith('{"grid": {\n  "lattice_type"') and out.endswith("\n ]\n}\n")
    # same seed -> same log; the pristine start is the allrules_8x8 fixture's trajectory
    g = load("allrules_8x8")
    assert [e["total_rate"] for e in doc["events"]] == list(g["events"]["total_rate"][:200])


def test_cli_event_log_replays_to_the_final_lattice(tmp_path):
    """the JSON log written by the CLI, replayed animate.py-style from the embedded
    initial state, reproduces the lattice the device ends with."""
    from demo1 import tables as T
    cfg = os.path.join(PKG, "config", "mixed_rates.yaml")
    doc = json.loads(run_cli([cfg, "-d", "12,10", "-n", "1500", "--seed", "11", "--embed-initial"]))
    dim = tuple(doc["grid"]["dim"])
    st = State.from_dict(doc["initial_state"]).status
    for ev in doc["events"]:
        T.replay_event(st, ev, dim)
    # same trajectory through KmcSim, then read the device lattice
    from demo1 import config
    with open(cfg) as f:
        c = config.from_dict(config.load_all([f]))
    sim = KmcSim(State(dim), c["rule_specs"], seed=11, chunk=1500)
    for _ in range(1500):
        sim.perform_random_move()
    assert (sim.current_state().status == st).all()
    assert {e["rule"] for e in doc["events"]} >= {"CreateTrefoil", "DestroyTrefoil", "MoveDivacancy"}
    sim.close()


def test_cli_runs_the_full_wse2_rule_set_without_warnings(tmp_path):
    """the reference's config/WSe2.yaml names MoveMonovacancy (a stub there, so the reference
    cannot run its own file); here the full rule set runs with no skip warning, the log carries the rule's
    payload {was, now, layer}, replays to the device's final lattice, and --validate-every holds."""
    import subprocess
    import sys
    from demo1 import tables as T
    cfg = os.path.join(PKG, "config", "wse2_full.yaml")
    args = [cfg, "-T", "1500", "-d", "16,16", "-n", "1200", "--seed", "20261019", "--state-seed", "5",
            "--embed-initial", "--validate-every", "300"]
    env = dict(os.environ, PYTHONPATH=PKG)
    proc = subprocess.run([sys.executable, "-m", "demo1"] + args, capture_output=True, text=True, env=env, cwd=ROOT,
                          timeout=300)
    assert proc.returncode == 0, proc.stderr
    assert "stub" not in proc.stderr and "skipped" not in proc.stderr
    doc = json.loads(proc.stdout)
    assert len(doc["rates"]) == 13 and "MoveMonovacancy:split-divacancy" in doc["rates"]
    mm = [e for e in doc["events"] if e["rule"] == "MoveMonovacancy"]
    assert len(mm) > 600 and all(sorted(e["move"]) == ["layer", "now", "was"] for e in mm)
    assert {e["kind"] for e in mm} >= {"natural", "split-divacancy"}
    dim = tuple(doc["grid"]["dim"])
    st = State.from_dict(doc["initial_state"]).status
    for ev in doc["events"]:
        T.replay_event(st, ev, dim)
    # the same trajectory on the oracle
    from util_cases import ko, WSE2_FULL_ORDER, wse2_full_rates
    o = ko.Oracle(dim, wse2_full_rates(1500.0), WSE2_FULL_ORDER, State.from_dict(doc["initial_state"]).status)
    _, want = o.run_philox(20261019, 0, 0, 1200)
    assert (o.status() == st).all()
    assert [e["total_rate"] for e in doc["events"]] == list(want["total_rate"])


def test_perform_given_move_accepts_reference_keys():
    st = State((8, 8))
    st.new_divacancy((2, 2))
    sim = KmcSim(st, [RuleSpec("MoveDivacancy", {"natural": 1.0, "missing-metal": 1.0, "missing-comb": 1.0,
                                                  "missing-both": 1.0}, False)], seed=1, chunk=1)
    sim.perform_given_move("MoveDivacancy", ((2, 2), (1, 3)))
    now = sim.current_state()
    assert now.is_divacancy((1, 3)) and now.is_pristine((2, 2))
    with pytest.raises(RuntimeError, match="does not exist"):
        sim.perform_given_move("MoveDivacancy", ((2, 2), (1, 3)))
    sim.close()


def test_cli_no_incremental_agrees_with_incremental():
    cfg = os.path.join(PKG, "config", "wse2_1500K.yaml")
    common = [cfg, "-T", "1500", "-d", "16,16", "-n", "120", "--seed", "3", "--state-seed", "4"]
    a = json.loads(run_cli(common))
    b = json.loads(run_cli(common + ["--no-incremental"]))
    assert a["events"] == b["events"]
    assert a["temperature"] == 1500.0 and len(a["rates"]) == 9


def test_cli_replica_batch_statistics():
    cfg = os.path.join(PKG, "config", "wse2_1500K.yaml")
    out = json.loads(run_cli([cfg, "-T", "1500", "-d", "64,64", "-n", "500", "--seed", "3", "--state-seed", "4",
                              "--replicas", "64"]))
    assert out["events_done"] == 64 * 500 and sum(out["histogram"].values()) == 64 * 500
    assert sum(out["divacancy_hops"]) == sum(v for k, v in out["histogram"].items() if k.startswith("MoveDivacancy"))
    # simulated time = sum of 1 / total_rate: ~200 divacancies x ~5 free directions at 2.8e-8 .. 8e-8 per hop give a
    # total rate of a few 1e-5, so a 500-event run lasts of the order of 1e7 time units
    assert 1e6 < out["mean_time"] < 1e9 and out["mean_squared_net_displacement"] > 0


def test_cli_independent_replica_states_are_generated_on_the_device():
    """--independent-states: replica r starts from gen_random_state under the Philox stream (state seed, r),
    so the histogram equals that of oracle runs started from oracle/state_gen.py lattices."""
    from util_cases import ko, WSE2_ORDER, wse2_rates         # puts oracle/ on sys.path
    import state_gen
    cfg = os.path.join(PKG, "config", "wse2_1500K.yaml")
    nrep, nsteps = 6, 300
    out = json.loads(run_cli([cfg, "-T", "1500", "-d", "16,16", "-n", str(nsteps), "--seed", "3", "--state-seed", "4",
                              "--replicas", str(nrep), "--independent-states"]))
    with open(cfg) as f:
        import yaml
        params = yaml.safe_load(f)["state"]["random"]
    mode = params.pop("mode")
    frac = {k: (float(v[:-1]) / 100 if isinstance(v, str) else float(v)) for k, v in params.items()}
    hist = np.zeros(17, dtype=np.int64)
    for r in range(nrep):
        st0 = state_gen.generate((16, 16), mode, frac, 4, r)
        o = ko.Oracle((16, 16), wse2_rates(1500.0), WSE2_ORDER, st0)
        o.run_philox(3, r, 0, nsteps, log=False)
        hist += np.asarray(o.hist())
    from demo1 import tables as T
    assert out["histogram"] == {"%s:%s" % T.CLASSES[c]: int(hist[c]) for c in WSE2_ORDER}


def test_kmcsim_zobrist_keys_follow_the_lattice():
    """KmcSim(zobrist_bits=...): every event dict carries the key of the lattice after it (reference
    sim.py:141-142); validate() checks the incrementally toggled key against a device pass from scratch."""
    from util_cases import ko  # noqa: F401  (puts oracle/ on sys.path)
    import zobrist as oz
    g = load("allrules_8x8")
    m = g["meta"]
    sim = KmcSim(State(m["dim"], g["status0"]), specs_from_meta(m), seed=m["seed"], chunk=100, zobrist_bits=48,
                 zobrist_seed=11)
    assert sim.is_zobrist_enabled() and sim.zobrist_key() == oz.key(g["status0"], 48, 11)
    keys = [sim.perform_random_move()["zobrist"] for _ in range(300)]
    assert sim.validate() is True
    assert keys[-1] == sim.zobrist_key() == oz.key(sim.current_state().status, 48, 11)
    assert all(0 <= k < 2 ** 48 for k in keys)
    sim.close()


def test_cli_zobrist_bits_adds_keys_to_the_log():
    cfg = os.path.join(PKG, "config", "mixed_rates.yaml")
    out = json.loads(run_cli([cfg, "-d", "8,8", "-n", "50", "--seed", "3", "--state-seed", "4", "--zobrist-bits", "20"]))
    keys = [e["zobrist"] for e in out["events"]]
    assert len(keys) == 50 and all(0 <= k < 2 ** 20 for k in keys)
    out2 = json.loads(run_cli([cfg, "-d", "8,8", "-n", "50", "--seed", "3", "--state-seed", "4", "--zobrist-bits", "20"]))
    assert [e["zobrist"] for e in out2["events"]] == keys


def test_bulk_json_lines_equal_the_per_event_api_across_launches():
    """KmcSim.iter_json_lines (prefetching the next launch on a helper thread) against perform_random_move,
    same seed, chunk boundaries inside the run; then the zero-rate stop."""
    g = load("allrules_8x8")
    m = g["meta"]
    a = KmcSim(State(m["dim"], g["status0"]), specs_from_meta(m), seed=m["seed"], chunk=257, validate_launches=True)
    b = KmcSim(State(m["dim"], g["status0"]), specs_from_meta(m), seed=m["seed"], chunk=100)
    fast = [line for lines in a.iter_json_lines(700) for line in lines]
    fast += [line for lines in a.iter_json_lines(333) for line in lines]          # resumes inside a chunk
    slow = [json.dumps(b.perform_random_move(), sort_keys=True) for _ in range(1033)]
    assert fast == slow
    assert json.loads(fast[0])["total_rate"] == g["events"][0]["total_rate"]
    a.close()
    b.close()
    # only FillDivacancy on a lattice with two divacancies: two events, then ValueError
    st = State((5, 5))
    st.status[3] = st.status[17] = 3
    sim = KmcSim(st, [RuleSpec("FillDivacancy", {"natural": 5.0}, False)], seed=1, chunk=10)
    got = []
    with pytest.raises(ValueError, match="total weight of zero"):
        for lines in sim.iter_json_lines(8):
            got += lines
    assert len(got) == 2 and all(json.loads(x)["rule"] == "FillDivacancy" for x in got)
    sim.close()


def test_cli_batch_is_independent_of_the_number_of_gpus():
    """The replica batch under torchrun (one process per GPU, replicas sharded by global index, NCCL all-reduce
    of the statistics) gives the same output as one process: replica r depends only on (seed, r)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = os.path.join(PKG, "config", "wse2_1500K.yaml")
    args = [cfg, "-T", "1500", "-d", "64,64", "-n", "400", "--seed", "3", "--state-seed", "4", "--replicas", "101",
            "--independent-states"]
    one = json.loads(run_cli(args))
    env = dict(os.environ, PYTHONPATH=PKG)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "demo1"] + args,
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    two = json.loads(p.stdout)
    assert two.pop("processes") == 2 and one.pop("processes") == 1
    assert two == one
