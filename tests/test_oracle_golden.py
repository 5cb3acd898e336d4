"""CPU oracle (oracle/kmc_oracle.c) pinned against fixtures produced by the unmodified reference.

Bar: bit-exact for class / site / slot / lattice / counts; total_rate is the exactly rounded
math.fsum (sim.py:138), also compared bit for bit."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import kmc_oracle as ko  # noqa: E402
import philox  # noqa: E402
from util_golden import GOLDEN, load, trajectory_names, boundary_tie_names, enabled_classes  # noqa: E402


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[[c] for c in ctr], [key[0]], [key[1]])
        assert tuple(int(x[0]) for x in got) == want


def test_c_draws_match_numpy_philox():
    for seed, rep, step0 in [(0, 0, 0), (20261019, 5, 0), ((1 << 40) + 17, 123456, (1 << 33) + 5)]:
        d = philox.draws(seed, rep, step0, 64)
        for t in range(64):
            assert ko.draw(seed, rep, step0 + t) == (d[t, 0], d[t, 1])
    assert ((d >= 0) & (d < 1)).all()


@pytest.mark.parametrize("name", trajectory_names())
def test_trajectory_matches_reference(name):
    g = load(name)
    m = g["meta"]
    o = ko.Oracle(m["dim"], g["rates"], m["class_order"], g["status0"])
    en = enabled_classes(m)
    assert (o.counts()[en] == g["counts0"][en]).all()
    done, ev = o.run_philox(m["seed"], m["replica"], 0, m["nsteps"])
    assert done == m["nsteps"]
    want = g["events"]
    for f in ("cls", "site", "slot"):
        bad = np.nonzero(ev[f] != want[f])[0]
        assert bad.size == 0, "%s: first %s mismatch at step %d" % (name, f, bad[0])
    assert (ev["total_rate"].view(np.uint64) == want["total_rate"].view(np.uint64)).all()
    assert (o.status() == g["status1"]).all()
    assert (o.counts()[en] == g["counts1"][en]).all()
    # slot table: the reference only holds slots of enabled rules
    slots = o.slots()
    mask = np.isin(slots, en)
    assert (np.where(mask, slots, 0xFF) == g["slots1"]).all()
    assert o.validate() == 0
    assert o.ties() == m["n_ties"]
    # naive prefix-sum total stays within the north-star tolerance of fsum
    rel = np.abs(ev["total_naive"] - ev["total_rate"]) / ev["total_rate"]
    assert rel.max() < 1e-12


@pytest.mark.parametrize("name", boundary_tie_names())
def test_boundary_ties_resolve_like_the_reference_bisect(name):
    """x == cumulative weight exactly (kmc.py:47-50: uniform(0, total) then bisect_left): the reference picks the
    class whose cumulative weight EQUALS x.  Every such draw is counted (north star: ties are reported)."""
    g = load(name)
    m = g["meta"]
    assert m["n_ties"] > 100
    o = ko.Oracle(m["dim"], g["rates"], m["class_order"], g["status0"])
    done, ev = o.run_draws(g["draws"])
    assert done == m["nsteps"]
    for f in ("cls", "site", "slot"):
        assert (ev[f] == g["events"][f]).all(), f
    assert (ev["total_rate"].view(np.uint64) == g["events"]["total_rate"].view(np.uint64)).all()
    assert (o.status() == g["status1"]).all()
    assert o.ties() == m["n_ties"]
    assert np.nonzero(ev["flags"] & 1)[0].tolist() == m["tie_steps"]
    assert o.validate() == 0


@pytest.mark.parametrize("name", ["allrules_8x8", "wse2_hot_24x24", "allrules_40x70"])
def test_incremental_equals_full_sweep(name):
    """kmco_validate restates KmcSim.validate (sim.py:159-168)."""
    g = load(name)
    m = g["meta"]
    o = ko.Oracle(m["dim"], g["rates"], m["class_order"], g["status0"])
    for chunk in range(10):
        o.run_philox(m["seed"] + 1, 0, chunk * 200, 200, log=False)
        assert o.validate() == 0
        slots, counts, rowcnt = ko.sweep(o.status(), m["dim"])
        assert (slots == o.slots()).all() and (counts == o.counts()).all()
        assert (rowcnt.sum(axis=0) == counts).all()


def replay_tracers(dim, status0, events):
    """Independent (pure Python) statement of the tracer rule, driven by an event log: a MoveDivacancy hop
    (slot 7 + k) carries the displacement to neighbour k of neighbors_and_mutuals (state.py:443-459); any other
    event that makes a site a divacancy starts a tracer at (0, 0).  Returns {site: (da, db)} of the divacancies
    on the final lattice, and that lattice."""
    import canon
    st = [int(x) for x in status0]
    disp = {i: (0, 0) for i, v in enumerate(st) if v == 3}
    for e in events:
        site, slot = int(e["site"]), int(e["slot"])
        p = canon.unlin(site, dim)
        if 7 <= slot <= 12:
            off = canon.NBR_NEAR_FAR[slot - 7][0]
            q = canon.lin(canon.add(p, off, dim), dim)
            da, db = disp.pop(site)
            disp[q] = (da + off[0], db + off[1])
            st[site], st[q] = 0, 3
            continue
        if slot == 0: new = {site: 3}
        elif slot == 1: new = {site: 0}
        elif slot in (2, 3): new = {site: st[site] | (slot - 1)}
        elif slot in (4, 5): new = {site: st[site] & ~(slot - 3)}
        elif slot == 6: new = {site: st[site] ^ 3}
        else:
            orient = (slot - 13) & 1
            nodes = [canon.lin(x, dim) for x in canon.trefoil_nodes(p, orient, dim)]
            new = {x: (4 + 3 * orient + r if slot < 15 else 3) for r, x in enumerate(nodes)}
        for x, v in new.items():
            if st[x] == 3 and v != 3:
                disp.pop(x)
            if v == 3 and st[x] != 3:
                disp[x] = (0, 0)
            st[x] = v
    return disp, st


@pytest.mark.parametrize("name", ["allrules_8x8", "wse2_hot_24x24", "wse2_16x16", "allrules_12x130"])
def test_clock_and_tracers_follow_the_reference_event_log(name):
    """SURVEY 8(f) rank 4 (extensions; the reference only logs total_rate for this, sim.py:138, README.md:36-37).
    The oracle's tracers against an independent replay of the REFERENCE-made event log, and its clocks against
    math.log / plain division over the logged totals (1e-12: kmc_log is within 2 ulp of libm)."""
    import math
    g = load(name)
    m = g["meta"]
    dim, n = tuple(m["dim"]), m["nsteps"]
    for mode in (1, 2):
        o = ko.Oracle(dim, g["rates"], m["class_order"], g["status0"])
        o.enable_tracers()
        o.set_clock(mode)
        _, ev = o.run_philox(m["seed"], m["replica"], 0, n)
        assert (ev["site"] == g["events"]["site"]).all()
        u3 = philox.draws3(m["seed"], m["replica"], 0, n)
        assert ((u3 > 0) & (u3 <= 1)).all()
        if mode == 1:
            want = math.fsum(1.0 / x for x in ev["total_naive"])
        else:
            want = math.fsum(-math.log(u) / x for u, x in zip(u3, ev["total_naive"]))
        assert abs(o.clock() - want) <= 1e-12 * want
        # the exactly rounded totals the reference logs (fsum) give the same time to 1e-12 as well
        want_ref = math.fsum((1.0 if mode == 1 else -math.log(u)) / x for u, x in zip(u3, g["events"]["total_rate"]))
        assert abs(o.clock() - want_ref) <= 1e-12 * want_ref
    disp, st = replay_tracers(dim, g["status0"], g["events"])
    assert st == [int(x) for x in g["status1"]]
    trc = o.tracers()
    assert sorted(disp) == np.nonzero(g["status1"] == 3)[0].tolist()
    for site, d in disp.items():
        assert tuple(trc[site]) == d
    sq, cnt = o.msd()
    assert cnt == len(disp) and sq == sum(a * a + a * b + b * b for a, b in disp.values())
    if name.startswith("wse2"):
        assert sq > 0


def test_deterministic_log_is_within_two_ulp_of_libm():
    import math
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.random(20000), 2.0 ** -rng.integers(0, 53, 200).astype(float),
                         [1.0, 2.0 ** -53, 0.5, 0.7071067811865476, 0.7071067811865475, 1.0 - 2.0 ** -53]])
    for x in xs:
        if x <= 0:
            continue
        a, b = ko.log(x), math.log(x)
        assert (a == 0.0 and b == 0.0) or abs(a - b) <= 2 * math.ulp(b)
    for step in (0, 1, 12345, (1 << 33) + 9):
        assert ko.draw3(20261019, 7, step) == philox.draws3(20261019, 7, step, 1)[0]


def test_rank_search_equals_linear_scan():
    g = load("allrules_64x64")
    m = g["meta"]
    o = ko.Oracle(m["dim"], g["rates"], m["class_order"], g["status1"])
    counts = o.counts()
    slots = o.slots()
    rng = np.random.default_rng(0)
    for c in range(13):
        if counts[c] == 0:
            continue
        flat = np.argwhere(slots == c)       # (site, slot) pairs -> canonical (row, slot, column) order
        d1 = m["dim"][1]
        flat = sorted(((int(s) // d1, int(j), int(s) % d1), int(s), int(j)) for s, j in flat)
        for r in rng.integers(0, counts[c], size=20):
            rc, site, slot = o.rank_to_move_scan(c, int(r))
            assert rc == 0 and (site, slot) == flat[r][1:]


def test_zero_total_weight_is_reported():
    # only FillDivacancy enabled on a pristine lattice: nothing can happen (kmc.py:31-32)
    rates = [0.0] * 13
    rates[1] = 5.0
    o = ko.Oracle((5, 5), rates, [1], np.zeros(25, np.uint8))
    rc, ev = o.step(0.3, 0.3)
    assert rc == 1 and ev["cls"] == 0xFF
    assert (o.status() == 0).all()


def test_known_answer_first_total_rate():
    # pristine 8x8 with test-cfg rates -> first total_rate = 64*(2*10 + 10) = 1920
    g = load("allrules_8x8")
    rates = np.array(g["rates"])
    rates[1] = 0.0
    rates[12] = 0.0
    o = ko.Oracle((8, 8), rates, g["meta"]["class_order"], np.zeros(64, np.uint8))
    rc, ev = o.step(0.5, 0.5)
    assert ev["total_rate"] == 1920.0


def test_batch_runner_matches_single():
    g = load("wse2_16x16")
    m = g["meta"]
    st = np.stack([g["status0"]] * 6)
    done, hist, st = ko.run_replicas(m["dim"], g["rates"], m["class_order"], st, seed=m["seed"],
                                     nsteps=500, threads=3)
    assert done == 6 * 500
    for r in range(6):
        o = ko.Oracle(m["dim"], g["rates"], m["class_order"], g["status0"])
        o.run_philox(m["seed"], r, 0, 500, log=False)
        assert (o.status() == st[r]).all()
    assert hist.sum() == done


def test_class_frequencies_match_the_reference_run_with_its_own_rng():
    """Statistical parity (north star): the reference stepped with ITS OWN random draws versus the
    oracle stepped with Philox draws -- per-step class frequencies over many trajectories must be
    samples of the same distribution (two-sample chi-square, p > 1e-4 at every checked step)."""
    import json
    from scipy.stats import chi2
    from util_golden import GOLDEN, chi2_two_sample
    z = np.load(os.path.join(GOLDEN, "native_stats.npz"))
    m = json.loads(str(z["meta"]))
    dim, nsteps = m["dim"], m["nsteps"]
    ours = np.zeros((nsteps, 13), dtype=np.int64)
    ntraj = 3000
    for r in range(ntraj):
        o = ko.Oracle(dim, z["rates"], m["class_order"], z["status0"])
        _, ev = o.run_philox(4242, r, 0, nsteps)
        ours[np.arange(nsteps), ev["cls"]] += 1
    worst = 1.0
    for t in (0, 1, 5, 10, 20, 39):
        stat, dof = chi2_two_sample(z["counts"][t], ours[t])
        worst = min(worst, chi2.sf(stat, dof))
    assert worst > 1e-4, worst
    stat, dof = chi2_two_sample(z["counts"].sum(axis=0), ours.sum(axis=0))
    assert chi2.sf(stat, dof) > 1e-4


def test_state_generator_restatement_matches_the_reference_under_philox():
    """oracle/state_gen.py against the reference's gen_random_state driven by philox.PhiloxRandom
    (fixtures made by oracle/make_golden.py:state_gen_philox_fixtures)."""
    import json
    import state_gen
    z = np.load(os.path.join(GOLDEN, "state_gen_philox.npz"))
    for i, m in enumerate(json.loads(str(z["meta"]))):
        got = state_gen.generate(tuple(m["dim"]), m["mode"], m["params"], m["seed"], m["replica"])
        assert (got == z["status_%d" % i]).all(), m
