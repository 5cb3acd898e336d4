"""Host-side logic of the demo1 package (no GPU): config schema and errors, state generators against
fixtures produced by the reference's gen_random_state, (de)serialisation, rate computation."""
import io
import json
import math
import os
import random
import warnings

import numpy as np
import pytest
import yaml

from demo1 import config, tables as T
from demo1.kmc import weighted_choice
from demo1.sim import RuleSpec, BOLTZMANN__eV_PER_K
from demo1.state import State, gen_random_state
from util_golden import GOLDEN, load

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmc-dichalcogen_b200")


def test_state_generators_match_reference_fixtures():
    z = np.load(os.path.join(GOLDEN, "state_gen.npz"))
    metas = json.loads(str(z["meta"]))
    for i, m in enumerate(metas):
        st = gen_random_state(tuple(m["dim"]), m["mode"], m["params"], rng=random.Random(m["seed"]))
        assert (st.status == z["status_%d" % i]).all(), m


def test_philox_backed_generators_match_the_reference_driven_by_the_same_stream():
    """tests/golden/state_gen_philox.npz: the reference's gen_random_state with a Philox-backed random.Random.
    The host twin (demo1.state.PhiloxRandom) must give the same lattices and consume the same number of
    draws; the plan handed to the device generator must describe the same runs / weights."""
    from demo1.state import PhiloxRandom, generator_plan
    z = np.load(os.path.join(GOLDEN, "state_gen_philox.npz"))
    for i, m in enumerate(json.loads(str(z["meta"]))):
        rng = PhiloxRandom(m["seed"], m["replica"])
        st = gen_random_state(tuple(m["dim"]), m["mode"], m["params"], rng=rng)
        want = z["status_%d" % i]
        assert (st.status == want).all(), m
        assert rng._q == m["draws"], m
        mode_id, codes, bounds, cumul = generator_plan(want.size, m["mode"], m["params"])
        if m["mode"] == "exact":
            assert mode_id == 0 and bounds[0] == 0
            for c, a, b in zip(codes, bounds[:-1], bounds[1:]):
                assert b - a == ((want == 3).sum() if c == 3 else ((want == 1) | (want == 2)).sum()), m
        else:
            assert mode_id == 1 and len(cumul) == len(codes) and all(x > 0 for x in cumul)


def test_state_roundtrip_and_validation():
    g = load("allrules_8x8")
    st = State((8, 8), g["status1"])
    st.validate()
    d = st.to_dict()
    assert set(d) == {"dim", "vacancies", "trefoils"}
    back = State.from_dict(json.loads(json.dumps(d)))
    assert (back.status == st.status).all()
    assert len(d["trefoils"]) == int(((st.status == 4) | (st.status == 7)).sum())
    bad = st.clone()
    bad.status[np.nonzero(bad.status >= 4)[0][0]] = 0
    with pytest.raises(AssertionError):
        bad.validate()


def test_shipped_configs_load():
    for name, nrules in [("wse2_1500K.yaml", 5), ("mixed_rates.yaml", 8)]:
        with open(os.path.join(PKG, "config", name)) as f:
            cfg = config.from_dict(config.load_all([f]))
        assert len(cfg["rule_specs"]) == nrules
        st = cfg["state_gen_func"]((16, 16), rng=random.Random(1))
        assert st.status.size == 256


def test_reference_style_configs_are_accepted():
    # test-cfg.yaml style: stale 'Rule' prefixes; WSe2.yaml style: MoveMonovacancy, which the
    # reference only stubs (rules.py:295-296), is a real rule here and loads without a warning
    d = yaml.safe_load("""
rules:
  RuleCreateDivacancy: {rate: 10}
  MoveMonovacancy: {barrier: {natural: 2.5, merge-divacancy: 3.0, split-divacancy: 2.3, swap-divacancy: 2.5}}
  FlipMonovacancy: {barrier: 3.4}
""")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cfg = config.from_dict(d)
    assert [s.name for s in cfg["rule_specs"]] == ["CreateDivacancy", "MoveMonovacancy", "FlipMonovacancy"]
    assert not w
    assert d == {}                      # consumed, as the reference does (quirk Q1)
    from demo1.sim import class_table
    _, rates, order = class_table(cfg["rule_specs"], 1500.0)
    assert order == [0, 13, 14, 15, 16, 12] and len(rates) == 17 and rates[15] > rates[13] > rates[14] > 0
    # a kind the rule does not have is still an error, a missing one too (reference sim.py:71-78)
    bad = yaml.safe_load("rules: {MoveMonovacancy: {rate: {natural: 1, merge-divacancy: 1, split-divacancy: 1}}}")
    with pytest.raises((RuntimeError, TypeError)):
        class_table(config.from_dict(bad)["rule_specs"], None)


@pytest.mark.parametrize("text,msg", [
    ("state: {}", "missing required key: 'rules'"),
    ("rules: {}\nextra: 1", "unrecognized section"),
    ("rules: {Nope: {rate: 1}}", "unknown rule: Nope"),
    ("rules: {CreateDivacancy: {rate: 1, barrier: 2}}", "exactly one of"),
    ("rules: {CreateDivacancy: {}}", "exactly one of"),
    ("rules: {CreateDivacancy: {rate: -1}}", "negative"),
    ("rules: {CreateDivacancy: {rate: abc}}", "real number or a mapping"),
    ("rules: {CreateDivacancy: {rate: {natural: x}}}", "non-numeric"),
    ("rules: {CreateDivacancy: {rate: 1, colour: red}}", "invalid attribute"),
    ("rules: {CreateDivacancy: {rate: 1, enabled: 'yes'}}", "enabled must be a boolean"),
    ("rules: {}\nstate: {random: {divacancy: 1%}}", "missing required key: 'mode'"),
    ("rules: {}\nstate: {random: {mode: fuzzy}}", "invalid choice for 'mode'"),
    ("rules: {}\nstate: {random: {mode: exact, divacancy: 70%, monovacancy: 70%}}", "sum of rates"),
    ("rules: {}\nstate: {random: {mode: exact, trefoil: 1%}}", "unrecognized parameter"),
    ("rules: {}\nstate: {path: x}", "invalid attribute"),
])
def test_config_errors(text, msg):
    with pytest.raises(RuntimeError, match="config: .*" + msg):
        config.from_dict(yaml.safe_load(text))


def test_merge_later_file_wins_and_keeps_first_appearance_order():
    a = io.StringIO("rules:\n  CreateDivacancy: {rate: 1}\n  FillDivacancy: {rate: 2}\n")
    b = io.StringIO("rules:\n  FillDivacancy: {rate: 5}\n  FlipMonovacancy: {rate: 3}\n")
    merged = config.load_all([a, b])
    assert list(merged["rules"]) == ["CreateDivacancy", "FillDivacancy", "FlipMonovacancy"]
    assert merged["rules"]["FillDivacancy"]["rate"] == 5


def test_rates_from_barriers_match_reference_formula():
    spec = RuleSpec("MoveDivacancy", {"natural": 2.25, "missing-metal": 2.11199155, "missing-comb": 2.52009312,
                                      "missing-both": 2.25}, rate_is_barrier=True)
    with pytest.raises(ValueError, match="temperature"):
        spec.rates(None)
    r = spec.rates(1500.0)
    assert r["natural"] == math.exp(-2.25 / (1500.0 * BOLTZMANN__eV_PER_K))
    g = load("wse2_16x16")              # rates computed by the reference itself
    assert r["natural"] == g["rates"][2] and r["missing-metal"] == g["rates"][3]
    assert r["missing-comb"] == g["rates"][4] and r["missing-both"] == g["rates"][5]


def test_weighted_choice_matches_reference_semantics():
    class Fixed:
        def __init__(self, u): self.u = u
        def uniform(self, a, b): return a + (b - a) * self.u
    pairs = [("a", 3.0), ("z", 0.0), ("b", 1.0), ("c", 1.0)]
    # sorted stably: b(1) c(1) a(3); cumulative 1, 2, 5
    assert weighted_choice(pairs, rng=Fixed(0.0)) == "b"
    assert weighted_choice(pairs, rng=Fixed(0.2)) == "b"        # x = 1.0 -> bisect_left -> index 0
    assert weighted_choice(pairs, rng=Fixed(0.21)) == "c"
    assert weighted_choice(pairs, rng=Fixed(0.41)) == "a"
    with pytest.raises(ValueError, match="total weight of zero"):
        weighted_choice([("a", 0.0)])
    with pytest.raises(ValueError, match="Negative weight"):
        weighted_choice([("a", -1.0), ("b", 2.0)])


def test_move_info_formats():
    dim = (8, 8)
    assert T.move_info(10, 0, dim) == {"node": [1, 2]}
    assert T.move_info(10, 3, dim) == {"node": [1, 2], "layer": 2}
    assert T.move_info(10, 4, dim) == {"node": [1, 2], "layer": 1}
    assert T.move_info(0, 7, dim) == {"was": [0, 0], "now": [7, 1]}
    assert T.move_info(62, 14, dim) == {"nodes": [[7, 6], [1, 6], [1, 4]]}
    g = load("allrules_8x8")
    for rec, info in zip(g["events"], g["meta"]["infos"]):
        mine = T.move_info(int(rec["site"]), int(rec["slot"]), dim)
        assert T.CLASSES[int(rec["cls"])] == (info["rule"], info["kind"])
        if "nodes" in mine:             # the reference lists a frozenset: order is arbitrary
            assert sorted(map(tuple, mine["nodes"])) == sorted(map(tuple, info["move"]["nodes"]))
        else:
            assert mine == info["move"]


def test_reference_move_keys_map_to_slots():
    dim = (8, 8)
    assert T.slot_of_move("CreateDivacancy", (1, 2), dim) == (10, 0)
    assert T.slot_of_move("FlipMonovacancy", (1, 2), dim) == (10, 6)
    assert T.slot_of_move("CreateMonovacancy", ((1, 2), 2, 2), dim) == (10, 3)
    assert T.slot_of_move("FillMonovacancy", ((1, 2), 1, 2), dim) == (10, 4)
    assert T.slot_of_move("MoveDivacancy", ((0, 0), (7, 1)), dim) == (0, 7)
    assert T.slot_of_move("MoveDivacancy", (10, 9), dim) == (10, 9)              # (site, slot) passes through
    assert T.slot_of_move("CreateTrefoil", frozenset([(7, 6), (1, 6), (1, 4)]), dim) == (62, 14)
    assert T.slot_of_move("DestroyTrefoil", [(0, 0), (2, 0), (0, 2)], dim) == (0, 15)
    with pytest.raises(ValueError):
        T.slot_of_move("MoveDivacancy", ((0, 0), (3, 3)), dim)
    # every (site, slot) round-trips through move_info for the single-node and pair rules
    for site in (0, 9, 63):
        for slot in range(7, 13):
            info = T.move_info(site, slot, dim)
            assert T.slot_of_move("MoveDivacancy", (tuple(info["was"]), tuple(info["now"])), dim) == (site, slot)


def test_replay_of_a_golden_log_reproduces_the_reference_lattice():
    """The event log alone determines the lattice (animate.py-style replay) -- on the reference's fixtures."""
    for name in ("allrules_8x8", "wse2_hot_24x24", "allrules_40x70"):
        g = load(name)
        dim = tuple(g["meta"]["dim"])
        st = g["status0"].copy()
        for rec in g["events"]:
            rule, _ = T.CLASSES[int(rec["cls"])]
            ev = {"rule": rule, "move": T.move_info(int(rec["site"]), int(rec["slot"]), dim)}
            T.replay_event(st, ev, dim)
        assert (st == g["status1"]).all(), name


def _labels_by_first_occurrence(keys):
    first, out = {}, []
    for k in keys:
        out.append(first.setdefault(int(k), len(first)))
    return out


def test_zobrist_keys_partition_the_trajectory_like_the_reference():
    """tests/golden/zobrist_ref.npz: the reference's own 64-bit Zobrist key after every event of a 5x5
    trajectory with revisits.  Its values are unreproducible (drawn on demand from `random`), but which steps
    share a key is a property of the lattices alone -- the host tracker must induce the same partition, and
    equal the from-scratch key (oracle/zobrist.py) of the replayed lattice."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    import zobrist as oz
    from demo1.zobrist import ZobristTracker
    z = np.load(os.path.join(GOLDEN, "zobrist_ref.npz"))
    m = json.loads(str(z["meta"]))
    dim = tuple(m["dim"])
    for bits, seed in ((64, 5), (40, (9 << 32) | 1)):
        tr = ZobristTracker(z["status0"], dim, bits, seed)
        keys = [tr.key]
        assert tr.key == oz.key(z["status0"], bits, seed) == ZobristTracker.from_scratch(z["status0"], bits, seed)
        for t, ev in enumerate(z["events"]):
            rule, _kind = T.CLASSES[int(ev["cls"])]
            keys.append(tr.apply({"rule": rule, "move": T.move_info(int(ev["site"]), int(ev["slot"]), dim)}))
            if t % 97 == 0:
                assert keys[-1] == oz.key(tr.status, bits, seed)
        assert (np.array(tr.status, dtype=np.uint8) == z["status1"]).all()
        assert keys[-1] == oz.key(z["status1"], bits, seed)
        assert all(k < 2 ** bits for k in keys)
        assert _labels_by_first_occurrence(keys) == _labels_by_first_occurrence(z["ref_keys"])
    assert len(set(int(k) for k in z["ref_keys"])) < len(z["ref_keys"])        # the fixture does revisit states


def test_bulk_json_writer_is_character_identical_to_the_per_event_path():
    """KmcSim.format_records_json (what the CLI writes) against json.dumps(KmcSim._format(rec), sort_keys=True)
    on synthetic records covering every class, slot, lattice edge (periodic wrap) and integer-valued rates."""
    from demo1 import _native as nat
    from demo1.sim import KmcSim
    from demo1.state import State
    sim = KmcSim.__new__(KmcSim)                       # no device: only the formatting state
    sim.state = State((7, 9))
    sim._rates13 = [10, 20.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 1e-23, 300.5, 100.0, 0.1, 55, 30.0, 60, 2e-7, 40.5]
    sim._order = [8, 9, 10, 11, 0, 1, 13, 14, 15, 16, 2, 3, 4, 5, 6, 7, 12]
    sim._zobrist = None
    slot_class = {0: [0], 1: [1], 2: [8, 9], 3: [8, 9], 4: [10, 11], 5: [10, 11], 6: [12]}
    slot_class.update({s: [2, 3, 4, 5] for s in range(7, 13)})
    slot_class.update({13: [6], 14: [6], 15: [7], 16: [7]})
    slot_class.update({s: [13, 14, 15, 16] for s in range(17, 29)})       # MoveMonovacancy (direction, layer)
    rng = np.random.default_rng(5)
    recs = []
    for slot, classes in slot_class.items():
        for cls in classes:
            for site in (0, 8, 9 * 6, 62, int(rng.integers(0, 63))):
                r = np.zeros((), dtype=nat.EVENT_FULL)
                r["site"], r["cls"], r["slot"] = site, cls, slot
                r["counts"] = rng.integers(0, 2 ** 20, 17)
                recs.append(r)
    recs = np.array(recs, dtype=nat.EVENT_FULL)
    fast = sim.format_records_json(recs)
    slow = [json.dumps(sim._format(r), sort_keys=True) for r in recs]
    assert fast == slow
    assert sim.format_records_json(recs[:0]) == []
    # perform_random_move decodes a whole launch at once: same dictionaries as the record-by-record formatter
    sim._buffer, sim._cursor, sim._decoded_for = recs, 0, None
    assert [sim.perform_random_move() for _ in range(len(recs))] == [sim._format(r) for r in recs]


def test_product_log_format_replays_in_the_references_animate_consumer():
    """tests/golden/animate_replay.npz: the reference's animate.py `do_basic` (its own consumer of the JSON log,
    sliced from the source and run unmodified by oracle/animate_replay.py in the build container) was fed logs
    in this package's formats (State.to_dict, tables.move_info) and reproduced the reference simulation's final
    lattices.  Here: the log this package writes for the same events still has the digest that was replayed,
    and the replayed arrays agree with the fixtures' final lattices."""
    import hashlib
    z = np.load(os.path.join(GOLDEN, "animate_replay.npz"))
    for m in json.loads(str(z["meta"])):
        g = load(m["name"])
        dim = tuple(g["meta"]["dim"])
        events = [{"rule": T.CLASSES[int(e["cls"])][0], "kind": T.CLASSES[int(e["cls"])][1],
                   "move": T.move_info(int(e["site"]), int(e["slot"]), dim)} for e in g["events"]]
        full = {"grid": {"lattice_type": "hexagonal", "coord_format": "axial", "dim": list(dim)},
                "initial_state": State(dim, g["status0"]).to_dict(), "events": events}
        full = json.loads(json.dumps(full))
        assert hashlib.sha256(json.dumps(full, sort_keys=True).encode()).hexdigest() == m["log_sha256"], m["name"]
        final = z["final_" + m["name"]]
        want = g["status1"].reshape(dim)
        assert ((final < 65536) == (want < 4)).all()
        assert (final[want < 4] == want[want < 4]).all()
        assert len(set(final[want >= 4].tolist())) == m["n_trefoils"] == (want == 4).sum() + (want == 7).sum()
