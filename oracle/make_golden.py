"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Each fixture is one trajectory of the reference (``/root/reference/demo1``) driven by injected
Philox draws through ``oracle/ref_oracle.RefStepper`` -- i.e. the reference's own counts,
``weighted_choice``, ``perform_given_move`` and incremental move caches -- plus the reference's
view of the lattice at the end (status plane, full slot table from its move caches, class
counts) and a ``sim.validate()`` pass.  State-generator fixtures come from the reference's
``gen_random_state`` under ``random.Random(seed)``.

The fixtures pin (a) the CPU restatement oracle/kmc_oracle.c and (b) the CUDA engine, both of
which must reproduce every event (class, site, slot), the fsum total rate, and the final lattice
bit for bit.
"""
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import canon  # noqa: E402
import philox  # noqa: E402
import ref_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TEST_CFG_RULES = {          # /root/reference/test-cfg.yaml:11-29 with the stale 'Rule' prefix removed
    "CreateMonovacancy": {"rate": {"from-double": 10, "make-empty": 300}},
    "FillMonovacancy": {"rate": {"make-double": 100, "from-empty": 100}},
    "CreateDivacancy": {"rate": 10},
    "MoveDivacancy": {"rate": {"natural": 12.0, "missing-metal": 3.0, "missing-comb": 7.0,
                               "missing-both": 7.0}},
    "CreateTrefoil": {"rate": 5000},
    "DestroyTrefoil": {"rate": 400},
}
ALL_RULES = dict(TEST_CFG_RULES, FlipMonovacancy={"rate": 55.0}, FillDivacancy={"rate": 20.0})
WSE2_RULES = {              # /root/reference/config/WSe2.yaml:2-24 minus the stub MoveMonovacancy
    "CreateMonovacancy": {"barrier": {"from-double": 6.72584072, "make-empty": 6.72584072}},
    "FlipMonovacancy": {"barrier": 3.41217502},
    "MoveDivacancy": {"barrier": {"natural": 2.25, "missing-metal": 2.11199155,
                                  "missing-comb": 2.52009312, "missing-both": 2.25}},
    "CreateTrefoil": {"barrier": 3.12293677},
    "DestroyTrefoil": {"barrier": 4.43051273},
}
# /root/reference/config/WSe2.yaml:2-24 with EVERY rule it names: MoveMonovacancy is a stub in the reference
# (rules.py:295-296); oracle/move_monovacancy.py supplies it as a rule class of the reference's framework
WSE2_FULL_RULES = {
    "CreateMonovacancy": {"barrier": {"from-double": 6.72584072, "make-empty": 6.72584072}},
    "FlipMonovacancy": {"barrier": 3.41217502},
    "MoveMonovacancy": {"barrier": {"natural": 2.56175687, "merge-divacancy": 3.03662851,
                                    "split-divacancy": 2.33673358, "swap-divacancy": 2.56175687}},
    "MoveDivacancy": {"barrier": {"natural": 2.25, "missing-metal": 2.11199155,
                                  "missing-comb": 2.52009312, "missing-both": 2.25}},
    "CreateTrefoil": {"barrier": 3.12293677},
    "DestroyTrefoil": {"barrier": 4.43051273},
}
# all nine rules live at comparable rates (17 classes)
ALL_MM_RULES = dict(ALL_RULES, MoveMonovacancy={"rate": {"natural": 30.0, "merge-divacancy": 60.0,
                                                         "split-divacancy": 20.0, "swap-divacancy": 40.0}})
# equal weights on a pristine lattice: 10*N == 5*(2N) == 10*N -> exercises the stable tie-break
TIE_RULES = {
    "CreateMonovacancy": {"rate": {"from-double": 5, "make-empty": 40}},
    "CreateDivacancy": {"rate": 10},
    "FillDivacancy": {"rate": 40},
    "FillMonovacancy": {"rate": {"make-double": 40, "from-empty": 20}},
    "FlipMonovacancy": {"rate": 40},
    "MoveDivacancy": {"rate": {"natural": 10, "missing-metal": 10, "missing-comb": 10, "missing-both": 10}},
    "CreateTrefoil": {"rate": 40},
    "DestroyTrefoil": {"rate": 40},
}


def exact_state(dim, frac_div, frac_mono, seed):
    """Our canonical 'exact' initial state recipe (host-side): shuffle the node
    list with random.Random(seed), first round(fd*N) nodes -> divacancy, next -> monovacancy with
    layer rng.choice([1, 2])."""
    rng = random.Random(seed)
    n = dim[0] * dim[1]
    nodes = list(range(n))
    rng.shuffle(nodes)
    nd = round(frac_div * n)
    nm = round((frac_div + frac_mono) * n)
    st = np.zeros(n, dtype=np.uint8)
    for i in nodes[:nd]:
        st[i] = 3
    for i in nodes[nd:nm]:
        st[i] = rng.choice([1, 2])
    return st


def trajectory(name, dim, rules, status0, seed, nsteps, temperature=None, rule_order=None, replica=0):
    st = ref_oracle.RefStepper(list(map(int, status0)), dim, rules, temperature=temperature,
                               rule_order=rule_order)
    d = philox.draws(seed, replica, 0, nsteps)
    ev = np.zeros(nsteps, dtype=[("cls", "u1"), ("slot", "u1"), ("site", "<u4"), ("total_rate", "<f8")])
    infos = []
    counts0 = st.counts()
    for t in range(nsteps):
        r = st.step(float(d[t, 0]), float(d[t, 1]))
        ev[t] = (r["class"], r["slot"], r["site"], r["total_rate"])
        if t < 64:
            infos.append({"rule": r["rule"], "kind": r["kind"],
                          "move": json.loads(json.dumps(r["info"], default=list))})
    with ref_oracle.ref_modules():
        st.validate()
    meta = {
        "name": name, "dim": list(dim), "seed": seed, "replica": replica, "nsteps": nsteps,
        "temperature": temperature, "rules": rules,
        "rule_order": rule_order or list(rules), "class_order": st.class_order(),
        "n_ties": st.n_ties, "infos": infos,
        "source": "unmodified /root/reference/demo1 driven by oracle/ref_oracle.py",
    }
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(
        path, meta=json.dumps(meta), rates=np.array(st.rates(), dtype=np.float64),
        status0=np.asarray(status0, dtype=np.uint8), events=ev, counts0=np.array(counts0, dtype=np.int64),
        status1=np.array(st.status(), dtype=np.uint8), counts1=np.array(st.counts(), dtype=np.int64),
        slots1=np.array(st.slot_table(), dtype=np.uint8))
    hist = np.bincount(ev["cls"], minlength=canon.N_CLASSES)
    print("%-28s dim=%s steps=%d ties=%d hist=%s (%d bytes)" % (
        name, dim, nsteps, st.n_ties, hist.tolist(), os.path.getsize(path)))


def state_gen_fixtures():
    """gen_random_state of the reference under random.Random(seed) (state.py:329-378)."""
    stmod = ref_oracle.ref()["state"]
    out = {}
    cases = [
        ("exact", (12, 9), {"divacancy": 0.05, "monovacancy": 0.10}, 11),
        ("exact", (64, 64), {"divacancy": 0.07, "monovacancy": 0.03}, 20261019),
        ("exact", (7, 5), {"divacancy": 0.0, "monovacancy": 0.5}, 3),
        ("approx", (12, 9), {"divacancy": 0.05, "monovacancy": 0.10}, 11),
        ("approx", (16, 16), {"divacancy": 0.2, "monovacancy": 0.3}, 5),
        ("approx", (8, 8), {"divacancy": 0.0, "monovacancy": 0.0}, 1),
    ]
    metas = []
    for i, (mode, dim, params, seed) in enumerate(cases):
        # the reference iterates RANDOM_PARAMS (a list built from a set); only the *values* sorted by
        # rate matter in 'exact', and 'approx' sorts by weight inside weighted_choice -- with the
        # distinct fractions used here neither depends on hash order.
        state = stmod.gen_random_state(dim, mode, dict(params), rng=random.Random(seed))
        status = np.array(ref_oracle.status_from_state(state), dtype=np.uint8)
        out["status_%d" % i] = status
        metas.append({"mode": mode, "dim": list(dim), "params": params, "seed": seed})
    path = os.path.join(OUT, "state_gen.npz")
    np.savez_compressed(path, meta=json.dumps(metas), **out)
    print("state_gen.npz: %d cases (%d bytes)" % (len(cases), os.path.getsize(path)))


PHILOX_STATE_CASES = [
    ("exact", (12, 9), {"divacancy": 0.05, "monovacancy": 0.10}, 11, 0),
    ("exact", (64, 64), {"divacancy": 0.05, "monovacancy": 0.05 + 1e-9}, 20261019, 0),
    ("exact", (64, 64), {"divacancy": 0.05, "monovacancy": 0.05 + 1e-9}, 20261019, 16383),
    ("exact", (7, 5), {"divacancy": 0.0, "monovacancy": 0.5}, 3, 2),
    ("exact", (33, 47), {"divacancy": 0.30, "monovacancy": 0.20}, (7 << 32) | 9, 1),
    ("approx", (12, 9), {"divacancy": 0.05, "monovacancy": 0.10}, 11, 0),
    ("approx", (16, 16), {"divacancy": 0.2, "monovacancy": 0.3}, 5, 7),
    ("approx", (8, 8), {"divacancy": 0.0, "monovacancy": 0.0}, 1, 0),
    ("approx", (64, 64), {"divacancy": 0.6, "monovacancy": 0.4}, (1 << 40) + 3, 5),
    ("approx", (40, 70), {"divacancy": 0.01, "monovacancy": 0.0}, 99, 3),
]


def state_gen_philox_fixtures():
    """The reference's gen_random_state (state.py:329-378) handed philox.PhiloxRandom(seed, replica) as its
    rng: pins the device-side generator (kmc_generate_state) and oracle/state_gen.py."""
    stmod = ref_oracle.ref()["state"]
    out, metas = {}, []
    for i, (mode, dim, params, seed, replica) in enumerate(PHILOX_STATE_CASES):
        rng = philox.PhiloxRandom(seed, replica)
        state = stmod.gen_random_state(dim, mode, dict(params), rng=rng)
        out["status_%d" % i] = np.array(ref_oracle.status_from_state(state), dtype=np.uint8)
        metas.append({"mode": mode, "dim": list(dim), "params": params, "seed": seed, "replica": replica,
                      "draws": rng.q})
    path = os.path.join(OUT, "state_gen_philox.npz")
    np.savez_compressed(path, meta=json.dumps(metas), **out)
    print("state_gen_philox.npz: %d cases (%d bytes)" % (len(metas), os.path.getsize(path)))


def zobrist_reference_fixture():
    """A trajectory of the reference with its Zobrist key enabled (state.py:52, 64 bits): per step the
    reference's own key and the event.  The key VALUES are not reproducible (drawn on demand from the global
    random module); what the fixture pins is which steps share a key, i.e. share a lattice."""
    import random as pyrandom
    pyrandom.seed(12345)                     # only makes this file reproducible
    dim, nsteps = (5, 5), 1500
    rules = dict(ALL_RULES)
    status0 = exact_state(dim, 0.2, 0.2, 4)
    st = ref_oracle.RefStepper(list(map(int, status0)), dim, rules, zobrist=64)
    d = philox.draws(77, 0, 0, nsteps)
    ev = np.zeros(nsteps, dtype=[("cls", "u1"), ("slot", "u1"), ("site", "<u4")])
    keys = np.zeros(nsteps + 1, dtype=np.uint64)
    keys[0] = st.sim.state.zobrist_key()
    for t in range(nsteps):
        r = st.step(float(d[t, 0]), float(d[t, 1]))
        ev[t] = (r["class"], r["slot"], r["site"])
        keys[t + 1] = st.sim.state.zobrist_key()
    with ref_oracle.ref_modules():
        st.validate()
    path = os.path.join(OUT, "zobrist_ref.npz")
    np.savez_compressed(path, meta=json.dumps({"dim": list(dim), "seed": 77, "nsteps": nsteps, "rules": rules,
                                               "class_order": st.class_order(), "bits": 64}),
                        rates=np.array(st.rates()), status0=status0, events=ev, ref_keys=keys,
                        status1=np.array(st.status(), dtype=np.uint8))
    print("zobrist_ref.npz: %d steps, %d distinct keys (%d bytes)" % (nsteps, len(set(keys.tolist())),
                                                                     os.path.getsize(path)))


def animate_replay_fixture():
    """the product's JSON log (grid / initial_state / events in the formats of
    kmc-dichalcogen_b200/demo1: State.to_dict, tables.move_info) replayed by the reference's own consumer,
    animate.py do_basic.  For three golden trajectories the replayed final lattice must equal the reference
    simulation's; the fixture keeps the replayed array and a digest of the product-formatted log."""
    import hashlib
    import importlib
    sys.path.insert(0, os.path.dirname(HERE))
    T = importlib.import_module("kmc-dichalcogen_b200.demo1.tables")
    S = importlib.import_module("kmc-dichalcogen_b200.demo1.state")
    import animate_replay
    out, metas = {}, []
    for name in ("allrules_8x8", "wse2_hot_24x24", "allrules_40x70"):
        g = np.load(os.path.join(OUT, name + ".npz"))
        m = json.loads(str(g["meta"]))
        dim = tuple(m["dim"])
        events = [{"rule": T.CLASSES[int(e["cls"])][0], "kind": T.CLASSES[int(e["cls"])][1],
                   "move": T.move_info(int(e["site"]), int(e["slot"]), dim)} for e in g["events"]]
        full = {"grid": {"lattice_type": "hexagonal", "coord_format": "axial", "dim": list(dim)},
                "initial_state": S.State(dim, g["status0"]).to_dict(), "events": events}
        full = json.loads(json.dumps(full))               # exactly what a consumer reads back
        final = animate_replay.replay(full)
        want = g["status1"].reshape(dim).astype(np.int64)
        assert ((final < 65536) == (want < 4)).all(), name
        assert (final[want < 4] == want[want < 4]).all(), name
        # nodes of one trefoil share an id; different trefoils differ
        ids = {}
        for a, b in zip(*np.nonzero(want >= 4)):
            ids.setdefault(int(final[a, b]), []).append(int(want[a, b]))
        assert all(sorted(v) in ([4, 5, 6], [7, 8, 9]) for v in ids.values()), name
        out["final_" + name] = final.astype(np.int32)
        metas.append({"name": name, "n_events": len(events), "n_trefoils": len(ids),
                      "log_sha256": hashlib.sha256(json.dumps(full, sort_keys=True).encode()).hexdigest()})
        print("animate replay %-16s %5d events, %d trefoils at the end: ok" % (name, len(events), len(ids)))
    path = os.path.join(OUT, "animate_replay.npz")
    np.savez_compressed(path, meta=json.dumps(metas), **out)


def _tie_draw(stepper, want_index):
    """A class draw u1 in [0, 1) for which ``random.uniform(0, total)`` (kmc.py:49: 0 + (total - 0) * u1) lands
    EXACTLY on the ``want_index``-th cumulative weight of weighted_choice's sorted list (kmc.py:35-42), or None
    when no double does.  The candidates are the doubles next to cum / total."""
    counts = stepper.counts()
    rates = stepper.rates()
    ws = sorted(w for w in (counts[c] * rates[c] for c in stepper.class_order()) if w != 0.0)
    cum, acc = [], 0
    for w in ws:
        acc = acc + w
        cum.append(acc)
    if want_index >= len(cum) - 1:          # the last boundary needs u1 == 1: not a draw
        return None
    total, target = cum[-1], cum[want_index]
    u = target / total
    for _ in range(4):
        u = math.nextafter(u, 0.0)
    for _ in range(9):
        if 0.0 <= u < 1.0 and 0.0 + (total - 0.0) * u == target:
            return u
        u = math.nextafter(u, 2.0)
    return None


def boundary_tie_trajectory(name, dim, rules, status0, seed, nsteps, every=3):
    """north star: "trajectories that diverge only at a cumulative-rate boundary tie are reported and counted".
    Philox draws, except that every ``every``-th step the class draw is replaced by one that puts x exactly on a
    cumulative-weight boundary (cycling over the boundaries), and a few steps use the extreme draws u1 = 0 and
    u1 = 1 - 2**-53.  The reference's own bisect_left (kmc.py:50) decides each of them; n_ties > 0."""
    st = ref_oracle.RefStepper(list(map(int, status0)), dim, rules)
    d = philox.draws(seed, 0, 0, nsteps).copy()
    ev = np.zeros(nsteps, dtype=[("cls", "u1"), ("slot", "u1"), ("site", "<u4"), ("total_rate", "<f8")])
    tie_steps = []
    counts0 = st.counts()
    for t in range(nsteps):
        if t % every == 0:
            u = _tie_draw(st, (t // every) % 5)
            if u is None:
                u = _tie_draw(st, 0)
            if u is not None:
                d[t, 0] = u
        elif t % 50 == 1:
            d[t, 0] = 0.0
        elif t % 50 == 2:
            d[t, 0] = 1.0 - 2.0 ** -53
        before = st.n_ties
        r = st.step(float(d[t, 0]), float(d[t, 1]))
        if st.n_ties > before:
            tie_steps.append(t)
        ev[t] = (r["class"], r["slot"], r["site"], r["total_rate"])
    with ref_oracle.ref_modules():
        st.validate()
    assert st.n_ties > 0 and st.n_ties == len(tie_steps)
    meta = {"name": name, "dim": list(dim), "seed": seed, "replica": 0, "nsteps": nsteps, "temperature": None,
            "rules": rules, "rule_order": list(rules), "class_order": st.class_order(), "n_ties": st.n_ties,
            "tie_steps": tie_steps, "injected_draws": True,
            "source": "unmodified /root/reference/demo1 driven by oracle/ref_oracle.py with injected draws"}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(
        path, meta=json.dumps(meta), rates=np.array(st.rates(), dtype=np.float64), draws=d,
        status0=np.asarray(status0, dtype=np.uint8), events=ev, counts0=np.array(counts0, dtype=np.int64),
        status1=np.array(st.status(), dtype=np.uint8), counts1=np.array(st.counts(), dtype=np.int64),
        slots1=np.array(st.slot_table(), dtype=np.uint8))
    print("%-28s dim=%s steps=%d boundary ties=%d (%d bytes)" % (name, dim, nsteps, st.n_ties, os.path.getsize(path)))


def move_monovacancy_fixtures():
    """the reference engine running oracle/move_monovacancy.py's rule class."""
    # config/WSe2.yaml as shipped, every rule: at 1500 K divacancies split and the monovacancies wander
    trajectory("wse2_full_16x16", (16, 16), WSE2_FULL_RULES, exact_state((16, 16), 0.05, 0.05, 20261019),
               seed=20261019, nsteps=3000, temperature=1500.0)
    trajectory("wse2_full_64x64", (64, 64), WSE2_FULL_RULES, exact_state((64, 64), 0.05, 0.05, 20261019),
               seed=20261019, nsteps=2000, temperature=1500.0)
    # hot and dense: all four MoveMonovacancy kinds, all four MoveDivacancy kinds and trefoils are live
    trajectory("wse2_full_hot_24x24", (24, 24), WSE2_FULL_RULES, exact_state((24, 24), 0.30, 0.10, 9),
               seed=98, nsteps=3000, temperature=6000.0)
    # all nine rules (17 classes) at comparable rates: pristine start, odd ragged dims, rows wider than 64
    trajectory("allrules_mm_8x8", (8, 8), ALL_MM_RULES, np.zeros(64, np.uint8), seed=17, nsteps=3000)
    trajectory("allrules_mm_5x7", (5, 7), ALL_MM_RULES, exact_state((5, 7), 0.2, 0.2, 6), seed=18, nsteps=2000)
    trajectory("allrules_mm_64x64", (64, 64), ALL_MM_RULES, exact_state((64, 64), 0.05, 0.05, 1), seed=19,
               nsteps=2000)
    trajectory("allrules_mm_12x130", (12, 130), ALL_MM_RULES, exact_state((12, 130), 0.2, 0.1, 8), seed=20,
               nsteps=2500, replica=1)


def boundary_tie_fixtures():
    # equal class weights on a pristine start (two weights of 770 -> u1 = 0.5 is a boundary), then mixed
    boundary_tie_trajectory("boundary_ties_7x11", (7, 11), TIE_RULES, np.zeros(77, np.uint8), seed=31, nsteps=600)
    # rows wider than 64 sites: also runs on the wide bit-plane kernel
    boundary_tie_trajectory("boundary_ties_5x70", (5, 70), ALL_RULES, exact_state((5, 70), 0.2, 0.1, 4), seed=32,
                            nsteps=600)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:                     # regenerate only the named groups, e.g. `make_golden.py boundary_ties`
        for g in sys.argv[1:]:
            {"boundary_ties": boundary_tie_fixtures, "move_monovacancy": move_monovacancy_fixtures, "state_gen": state_gen_fixtures,
             "state_gen_philox": state_gen_philox_fixtures, "zobrist": zobrist_reference_fixture,
             "animate": animate_replay_fixture, "native_stats": native_statistics}[g]()
        return
    # 1. all eight rules, pristine start: trefoils are created and destroyed along the way
    trajectory("allrules_8x8", (8, 8), ALL_RULES, np.zeros(64, np.uint8), seed=7, nsteps=3000)
    # 2. the reference's own test-cfg.yaml rule set, ragged odd dims
    trajectory("testcfg_5x7", (5, 7), TEST_CFG_RULES, np.zeros(35, np.uint8), seed=11, nsteps=2000)
    trajectory("testcfg_9x5", (9, 5), TEST_CFG_RULES, exact_state((9, 5), 0.2, 0.2, 5), seed=12, nsteps=2000)
    # 3. WSe2 barriers at 1500 K on a seeded-defect lattice (BASELINE config 2 recipe, smaller)
    trajectory("wse2_16x16", (16, 16), WSE2_RULES, exact_state((16, 16), 0.05, 0.05, 20261019),
               seed=20261019, nsteps=3000, temperature=1500.0)
    trajectory("wse2_64x64", (64, 64), WSE2_RULES, exact_state((64, 64), 0.05, 0.05, 20261019),
               seed=20261019, nsteps=2000, temperature=1500.0)
    # hotter, denser: trefoil create/destroy and all four MoveDivacancy kinds are live
    trajectory("wse2_hot_24x24", (24, 24), WSE2_RULES, exact_state((24, 24), 0.30, 0.05, 9),
               seed=99, nsteps=3000, temperature=4000.0)
    # 4. equal class weights -> stable tie-break by canonical class order; two rule orders
    trajectory("ties_7x11", (7, 11), TIE_RULES, np.zeros(77, np.uint8), seed=3, nsteps=2500)
    trajectory("ties_reordered_7x11", (7, 11), TIE_RULES, np.zeros(77, np.uint8), seed=3, nsteps=2500,
               rule_order=["DestroyTrefoil", "FlipMonovacancy", "CreateDivacancy", "MoveDivacancy",
                           "FillMonovacancy", "CreateMonovacancy", "CreateTrefoil", "FillDivacancy"])
    # 5. mixed rules at the headline lattice size
    trajectory("allrules_64x64", (64, 64), ALL_RULES, exact_state((64, 64), 0.05, 0.05, 1), seed=5,
               nsteps=2000)
    # 6. non-square, larger than a warp's row span
    trajectory("allrules_40x70", (40, 70), ALL_RULES, exact_state((40, 70), 0.1, 0.1, 2), seed=6,
               nsteps=1500, replica=3)
    # 7. the reference's own default run: `python3 -m demo1 test-cfg.yaml` = 200x200, pristine start (BASELINE config 1)
    trajectory("testcfg_200x200", (200, 200), TEST_CFG_RULES, np.zeros(40000, np.uint8), seed=20261019, nsteps=3000)
    # 8. rows wider than 64 sites with a ragged last word (the wide bit-plane kernel's layout), trefoils live
    trajectory("allrules_12x130", (12, 130), ALL_RULES, exact_state((12, 130), 0.2, 0.1, 8), seed=13, nsteps=2500, replica=1)
    boundary_tie_fixtures()
    move_monovacancy_fixtures()
    state_gen_fixtures()
    state_gen_philox_fixtures()
    zobrist_reference_fixture()
    animate_replay_fixture()
    native_statistics()


def native_statistics():
    """The reference run with ITS OWN random draws (random / numpy.random, seeded only to make this file
    reproducible): per-step class frequencies over many short trajectories.  Pins the probability law
    -- in particular that our canonical tie-break and in-class order do not change it."""
    import random as pyrandom
    mods = ref_oracle.ref()
    dim, nsteps, ntraj = (12, 12), 40, 400
    status0 = exact_state(dim, 0.15, 0.10, 77)
    counts = np.zeros((nsteps, canon.N_CLASSES), dtype=np.int64)
    for tr in range(ntraj):
        pyrandom.seed(1000 + tr)
        np.random.seed(1000 + tr)
        specs = ref_oracle.make_rule_specs(ALL_RULES)
        init = ref_oracle.state_from_status(list(map(int, status0)), dim)
        with ref_oracle.ref_modules():
            sim = mods["sim"].KmcSim(init, specs, incremental=True)
            for t in range(nsteps):
                d = sim.perform_random_move()
                counts[t, canon.CLASS_ID[(d["rule"], d["kind"])]] += 1
    st = ref_oracle.RefStepper(list(map(int, status0)), dim, ALL_RULES)
    path = os.path.join(OUT, "native_stats.npz")
    np.savez_compressed(path, meta=json.dumps({"dim": list(dim), "nsteps": nsteps, "ntraj": ntraj,
                                               "class_order": st.class_order(),
                                               "source": "reference KmcSim.perform_random_move with its own RNGs"}),
                        status0=status0, rates=np.array(st.rates()), counts=counts)
    print("native_stats.npz: %d trajectories x %d steps (%d bytes)" % (ntraj, nsteps, os.path.getsize(path)))


if __name__ == "__main__":
    main()
