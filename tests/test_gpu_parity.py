"""Parity of the CUDA engine (through the C ABI) with the CPU oracle and the golden fixtures.

Bar: bit-exact event sequences (class, site, slot), lattices, slot tables and counts; the logged
total rate (sequential fp64 sum) bit-equal to the oracle's, and the reference's math.fsum total
recomputed from the logged counts bit-equal to the fixture's (1e-12 relative is the north-star
tolerance for rates; we hold 0)."""
import math

import numpy as np
import pytest

from util_cases import (ko, philox, ALL_ORDER, TEST_RATES, WSE2_ORDER, wse2_rates, exact_state,
                        evolved_state, iid_state, MM_ORDER, MM_RATES, WSE2_FULL_ORDER, wse2_full_rates,
                        reference_rules_only)
from util_golden import load, trajectory_names, boundary_tie_names, enabled_classes
from demo1 import _native as nat
from demo1.engine import Engine

pytestmark = pytest.mark.gpu


def fsum_totals(ev, rates, order):
    rates = np.asarray(rates, dtype=np.float64)
    return np.array([math.fsum(float(int(c[k])) * float(rates[k]) for k in order) for c in ev["counts"]])


def assert_events_equal(dev, want, what):
    for f in ("cls", "site", "slot"):
        bad = np.nonzero(dev[f] != want[f])[0]
        assert bad.size == 0, "%s: first %s mismatch at step %d: got %r want %r" % (
            what, f, bad[0], dev[bad[0]], want[bad[0]])


@pytest.mark.parametrize("name", trajectory_names())
def test_golden_trajectory_philox(name):
    g = load(name)
    m = g["meta"]
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.upload_state(g["status0"])
    # replica0 = fixture's replica id: the Philox key is (seed, replica0 + r)
    ev = eng.run(m["nsteps"], m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL, validate=True)[0]
    assert_events_equal(ev, g["events"], name)
    got = fsum_totals(ev, g["rates"], m["class_order"])
    assert (got.view(np.uint64) == g["events"]["total_rate"].view(np.uint64)).all()
    rel = np.abs(ev["total_rate"] - g["events"]["total_rate"]) / g["events"]["total_rate"]
    assert rel.max() < 1e-12
    assert (eng.download_state()[0] == g["status1"]).all()
    en = enabled_classes(m)
    eng.sweep()
    assert (eng.counts()[0][en] == g["counts1"][en]).all()
    slots = eng.slots()[0]
    mask = np.isin(slots, en)
    assert (np.where(mask, slots, 0xFF) == g["slots1"]).all()
    assert (eng.histogram() == np.bincount(g["events"]["cls"], minlength=17)).all()
    assert eng.steps_done()[0] == m["nsteps"]
    assert eng.ties() == m["n_ties"]
    eng.close()


@pytest.mark.parametrize("variant", [1, 3, 5])
@pytest.mark.parametrize("name", ["allrules_40x70", "testcfg_200x200", "allrules_12x130"])
def test_golden_wide_rows_on_every_kernel(name, variant):
    """The reference-made fixtures with rows wider than 64 sites (among them the reference's own default run,
    200x200) on the byte kernel in shared memory, the byte kernel with the lattice in HBM and the wide bit-plane
    kernel."""
    g = load(name)
    m = g["meta"]
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.set_replica_kernel(variant)
    eng.upload_state(g["status0"])
    half = m["nsteps"] // 2
    ev = np.concatenate([eng.run(half, m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL, validate=True)[0],
                         eng.run(m["nsteps"] - half, m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL,
                                 validate=True)[0]])
    assert_events_equal(ev, g["events"], name)
    got = fsum_totals(ev, g["rates"], m["class_order"])
    assert (got.view(np.uint64) == g["events"]["total_rate"].view(np.uint64)).all()
    assert (eng.download_state()[0] == g["status1"]).all()
    eng.close()


@pytest.mark.parametrize("name", ["allrules_8x8", "ties_7x11", "wse2_hot_24x24"])
def test_golden_trajectory_injected_draws(name):
    """Same trajectories through the draw-injection seam (kmc_step_draws), in two chunks."""
    g = load(name)
    m = g["meta"]
    d = philox.draws(m["seed"], m["replica"], 0, m["nsteps"])
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.upload_state(g["status0"])
    half = m["nsteps"] // 2 + 7
    ev1 = eng.step_draws(d[:half], log_mode=nat.LOG_COMPACT, validate=True)[0]
    ev2 = eng.step_draws(d[half:], log_mode=nat.LOG_COMPACT, validate=True)[0]
    ev = np.concatenate([ev1, ev2])
    assert_events_equal(ev, g["events"], name)
    assert (eng.download_state()[0] == g["status1"]).all()
    eng.close()


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, "noinc"])
@pytest.mark.parametrize("name", boundary_tie_names())
def test_boundary_ties_on_every_kernel(name, variant, monkeypatch):
    """north star: "trajectories that diverge only at a cumulative-rate boundary tie are reported and counted".
    The fixtures (made from the unmodified reference, oracle/make_golden.py:boundary_tie_trajectory) place a
    third of the class draws EXACTLY on a cumulative weight, where the reference's bisect_left (kmc.py:50) takes
    the class whose cumulative weight equals x.  Every kernel must take the same class, flag the event and
    count the tie."""
    g = load(name)
    m = g["meta"]
    d1 = m["dim"][1]
    if (variant in (2, 4) and d1 > 64) or (variant == 5 and d1 <= 64):
        pytest.skip("kernel variant does not take this row length")
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.upload_state(g["status0"])
    if variant == "noinc":
        monkeypatch.setenv("KMC_NOINC_PLANES", "0")      # the byte-lattice path (the plane path has its own tie test)
        ev = eng.step_draws_noinc(g["draws"], log_mode=nat.LOG_FULL)[0]
    else:
        eng.set_replica_kernel(variant)
        half = m["nsteps"] // 2 + 3
        ev = np.concatenate([eng.step_draws(g["draws"][:half], log_mode=nat.LOG_FULL, validate=True)[0],
                             eng.step_draws(g["draws"][half:], log_mode=nat.LOG_FULL, validate=True)[0]])
    assert_events_equal(ev, g["events"], name)
    got = fsum_totals(ev, g["rates"], m["class_order"])
    assert (got.view(np.uint64) == g["events"]["total_rate"].view(np.uint64)).all()
    assert np.nonzero(ev["flags"] & 1)[0].tolist() == m["tie_steps"]
    assert m["n_ties"] > 100 and eng.ties() == m["n_ties"]
    assert (eng.download_state()[0] == g["status1"]).all()
    eng.close()


@pytest.mark.parametrize("dim,rates,order,nrep,nsteps", [
    ((64, 64), TEST_RATES, ALL_ORDER, 6, 20000),
    ((64, 64), wse2_rates(1500.0), WSE2_ORDER, 5, 20000),
    ((33, 47), TEST_RATES, ALL_ORDER, 3, 8000),
    ((5, 5), TEST_RATES, ALL_ORDER, 4, 5000),
    ((96, 80), TEST_RATES, ALL_ORDER, 2, 8000),
    ((7, 130), TEST_RATES, ALL_ORDER, 2, 5000),
])
def test_replicas_match_oracle(dim, rates, order, nrep, nsteps):
    """Several replicas, long runs, two launches (step counter continuity), vs the C oracle."""
    seed = 424242
    st0 = np.stack([exact_state(dim, 0.05 + 0.02 * r, 0.05, 100 + r) for r in range(nrep)])
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.upload_state(st0)
    n1 = nsteps // 3
    ev1 = eng.run(n1, seed, replica0=10, log_mode=nat.LOG_FULL, validate=True)
    ev2 = eng.run(nsteps - n1, seed, replica0=10, log_mode=nat.LOG_FULL, validate=True)
    ev = np.concatenate([ev1, ev2], axis=1)
    final = eng.download_state()
    hist = eng.histogram_per_replica()
    for r in range(nrep):
        o = ko.Oracle(dim, rates, order, st0[r])
        done, want = o.run_philox(seed, 10 + r, 0, nsteps)
        assert done == nsteps
        assert_events_equal(ev[r], want, "replica %d" % r)
        assert (ev[r]["total_rate"].view(np.uint64) == want["total_naive"].view(np.uint64)).all()
        assert (ev[r]["counts"][:, :13] == want["counts"][:, :13]).all() and (ev[r]["counts"][:, 13:] == 0).all()
        assert (ev[r]["flags"] == want["flags"]).all()
        assert (final[r] == o.status()).all()
        assert (hist[r] == o.hist()).all()
    eng.close()


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("dim", [(64, 64), (33, 47), (5, 5), (9, 64), (130, 7), (16, 20)])
def test_both_replica_kernels_match_oracle(dim, variant):
    """The byte-lattice kernel (1), the bit-sliced kernel (2) and the byte-lattice kernel with the
    lattice left in HBM (3) on the same lattices."""
    nrep, nsteps, seed = 3, 6000, 99
    st0 = np.stack([evolved_state(dim, 20 + r, nsteps=300) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.set_replica_kernel(variant)
    eng.upload_state(st0)
    ev = eng.run(nsteps, seed, log_mode=nat.LOG_FULL, validate=True)
    final = eng.download_state()
    for r in range(nrep):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[r])
        _, want = o.run_philox(seed, r, 0, nsteps)
        assert_events_equal(ev[r], want, "variant %d replica %d" % (variant, r))
        assert (ev[r]["counts"][:, :13] == want["counts"][:, :13]).all() and (ev[r]["counts"][:, 13:] == 0).all()
        assert (ev[r]["total_rate"].view(np.uint64) == want["total_naive"].view(np.uint64)).all()
        assert (final[r] == o.status()).all()
    eng.close()


def test_large_lattice_incremental_1024():
    """BASELINE config 5 shape: 1024x1024 lattices (1 MiB each) do not fit shared memory; the incremental
    kernel keeps them in HBM/L2 and only the row hierarchy on chip."""
    dim, nrep, nsteps = (1024, 1024), 3, 3000
    st0 = np.stack([exact_state(dim, 0.05, 0.05, 300 + r) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.upload_state(st0)
    # three launches: the second and third reuse the row hierarchy the previous one left in HBM
    ev = np.concatenate([eng.run(n, 77, log_mode=nat.LOG_FULL, validate=True) for n in (1000, 1500, 500)], axis=1)
    final = eng.download_state()
    for r in range(nrep):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[r])
        _, want = o.run_philox(77, r, 0, nsteps)
        assert_events_equal(ev[r], want, "1024^2 replica %d" % r)
        assert (ev[r]["counts"][:, :13] == want["counts"][:, :13]).all() and (ev[r]["counts"][:, 13:] == 0).all()
        assert (final[r] == o.status()).all()
    eng.close()


@pytest.mark.parametrize("cta", ["0", "1"])
@pytest.mark.parametrize("dim", [(12, 128), (40, 192), (7, 320), (130, 128), (64, 2048), (5, 128), (1100, 128),
                                 (12, 130), (40, 200), (7, 70), (9, 65), (200, 200), (20, 2044), (5, 67), (33, 127)])
def test_wide_bitplane_kernel_matches_oracle(dim, cta, monkeypatch):
    monkeypatch.setenv("KMC_WIDE_CTA", cta)        # 0: one warp per replica, 1: a CTA per replica (warps split by role)
    """Variant 5: bit planes of rows wider than one word kept in HBM, row table in shared memory."""
    nrep, nsteps, seed = 3, 5000, 4711
    st0 = np.stack([evolved_state(dim, 40 + r, nsteps=300) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.set_replica_kernel(5)
    eng.upload_state(st0)
    n1 = 1777
    ev = np.concatenate([eng.run(n1, seed, log_mode=nat.LOG_FULL, validate=True),
                         eng.run(nsteps - n1, seed, log_mode=nat.LOG_FULL, validate=True)], axis=1)
    final = eng.download_state()
    for r in range(nrep):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[r])
        _, want = o.run_philox(seed, r, 0, nsteps)
        assert_events_equal(ev[r], want, "wide replica %d" % r)
        assert (ev[r]["counts"][:, :13] == want["counts"][:, :13]).all() and (ev[r]["counts"][:, 13:] == 0).all()
        assert (ev[r]["total_rate"].view(np.uint64) == want["total_naive"].view(np.uint64)).all()
        assert (final[r] == o.status()).all()
    eng.close()


@pytest.mark.parametrize("cta", ["0", "1"])
def test_wide_kernel_interleaves_with_the_byte_paths(cta, monkeypatch):
    monkeypatch.setenv("KMC_WIDE_CTA", cta)
    """Planes are authoritative after a wide launch; every byte-side entry point must see the same lattice:
    wide run -> sweep -> perform_given -> byte-kernel run -> wide run -> injected draws."""
    dim, seed = (24, 128), 31
    st0 = evolved_state(dim, 5, nsteps=300)
    o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0)
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=1)
    eng.set_replica_kernel(5)
    eng.upload_state(st0)
    ev = eng.run(700, seed, log_mode=nat.LOG_COMPACT, validate=True)[0]
    _, want = o.run_philox(seed, 0, 0, 700)
    assert_events_equal(ev, want, "wide 1")
    eng.sweep()
    assert (eng.counts()[0] == reference_rules_only(counts=o.counts())).all()
    assert (eng.slots()[0] == reference_rules_only(o.slots())).all()
    # an explicit move: first CreateDivacancy-capable site
    site = int(np.flatnonzero(o.status() == 0)[0])
    o.perform_given(site, 0)
    eng.perform_given(0, site, 0)
    eng.set_replica_kernel(3)
    ev = eng.run(500, seed, log_mode=nat.LOG_COMPACT, validate=True)[0]
    _, want = o.run_philox(seed, 0, 700, 500)
    assert_events_equal(ev, want, "byte after wide")
    eng.set_replica_kernel(5)
    ev = eng.run(500, seed, log_mode=nat.LOG_COMPACT, validate=True)[0]
    _, want = o.run_philox(seed, 0, 1200, 500)
    assert_events_equal(ev, want, "wide 2")
    d = philox.draws(seed + 1, 0, 0, 300)
    ev = eng.step_draws(d, log_mode=nat.LOG_COMPACT, validate=True)[0]
    _, want = o.run_draws(d)
    assert_events_equal(ev, want, "wide draws")
    assert (eng.download_state()[0] == o.status()).all()
    assert (eng.histogram_per_replica()[0] == o.hist()).all()
    eng.close()


@pytest.mark.parametrize("cta", ["0", "1"])
def test_wide_kernel_stops_on_zero_total_rate(cta, monkeypatch):
    monkeypatch.setenv("KMC_WIDE_CTA", cta)
    dim = (8, 128)
    rates = [0.0] * 13
    rates[1] = 5.0                                   # FillDivacancy only: runs out after the divacancies
    st0 = np.zeros(dim[0] * dim[1], np.uint8)
    st0[[3, 200, 777]] = 3
    eng = Engine(dim, rates, [1], n_replicas=2)
    eng.set_replica_kernel(5)
    eng.upload_state(np.stack([st0, st0]))
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.run(10, 1, log_mode=nat.LOG_COMPACT)
    ev = eng._last_events
    assert (eng.steps_done() == 3).all()
    assert (ev[:, 3]["cls"] == nat.CLASS_NONE).all()
    assert (eng.download_state() == 0).all()
    eng.close()


@pytest.mark.parametrize("variant", [1, 2, 4])
def test_kmc_clock_is_the_sum_of_inverse_total_rates(variant):
    """Extension: t = sum 1/total_rate, accumulated in event order ACROSS launches, bit-equal
    to the same sum over the logged totals; and for the no-incremental path."""
    dim, nrep = (16, 20) if variant == 1 else (64, 64), 3
    st0 = np.stack([exact_state(dim, 0.1, 0.1, 7 + r) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.set_replica_kernel(variant)
    eng.enable_clock()
    eng.upload_state(st0)
    ev = np.concatenate([eng.run(n, 3, log_mode=nat.LOG_COMPACT) for n in (700, 300)], axis=1)
    got = eng.clock()
    for r in range(nrep):
        t = 0.0
        for x in ev[r]["total_rate"]:
            t += 1.0 / float(x)
        assert got[r] == t
    ev2 = eng.run_noinc(50, 3, log_mode=nat.LOG_COMPACT)
    got2 = eng.clock()
    for r in range(nrep):
        t = got[r]
        for x in ev2[r]["total_rate"]:
            t += 1.0 / float(x)
        assert got2[r] == t
    eng.close()


@pytest.mark.parametrize("variant,dim", [(1, (33, 47)), (2, (64, 64)), (3, (16, 80)), (4, (64, 64)), (4, (9, 24)),
                                         (5, (24, 128)), ("noinc", (20, 36))])
def test_stochastic_clock_and_tracers_match_the_oracle(variant, dim):
    """the BKL clock t += -ln(u3) / total_rate (third Philox draw per event, deterministic
    fp64 logarithm) is BIT-equal to the oracle's on every kernel, across launches; the divacancy tracers
    (unwrapped displacement carried through MoveDivacancy hops, reset when a divacancy appears) and the summed
    squared displacement are integer-exact.  Hot WSe2 rates: trefoils form and break, so tracers are born and die."""
    import math
    nrep, seed = 3, 77
    rates, order = wse2_rates(4000.0), WSE2_ORDER
    st0 = np.stack([exact_state(dim, 0.30, 0.05, 40 + r) for r in range(nrep)])
    eng = Engine(dim, rates, order, n_replicas=nrep)
    if variant != "noinc":
        eng.set_replica_kernel(variant)
    eng.enable_clock("stochastic")
    eng.enable_tracers()
    eng.upload_state(st0)
    chunks = (60, 40) if variant == "noinc" else (1500, 1, 999)
    for nsteps in chunks:
        if variant == "noinc":
            eng.run_noinc(nsteps, seed, replica0=5)
        else:
            eng.run(nsteps, seed, replica0=5, validate=True)
    t = eng.clock()
    sq, cnt = eng.msd()
    final = eng.download_state()
    for r in range(nrep):
        o = ko.Oracle(dim, rates, order, st0[r])
        o.enable_tracers()
        o.set_clock(2)
        o.run_philox(seed, 5 + r, 0, sum(chunks), log=False)
        assert (final[r] == o.status()).all()
        assert t[r] == o.clock() and t[r] > 0, (t[r], o.clock())
        div = final[r] == 3
        assert (eng.tracers(r)[div] == o.tracers()[div]).all()
        assert (int(sq[r]), int(cnt[r])) == o.msd()
        assert cnt[r] == div.sum()
    assert sq.sum() > 0
    # reset_stats restarts clock and tracers; a replaced lattice restarts the tracers
    eng.reset_stats()
    assert (eng.clock() == 0).all() and (eng.msd()[0] == 0).all()
    eng.close()


def test_mean_residence_clock_and_stochastic_clock_agree_on_average():
    """E[-ln u3] = 1: over many replicas the stochastic clock averages to the mean-residence clock."""
    dim, nrep, nsteps = (64, 64), 512, 400
    rates, order = wse2_rates(1500.0), WSE2_ORDER
    st0 = np.tile(exact_state(dim, 0.05, 0.05, 3), (nrep, 1))
    out = []
    for mode in (1, 2):
        eng = Engine(dim, rates, order, n_replicas=nrep)
        eng.enable_clock(mode)
        eng.upload_state(st0)
        eng.run(nsteps, 11)
        eng.check_status()
        out.append(eng.clock())
        eng.close()
    mean, stoch = out
    assert abs(stoch.mean() / mean.mean() - 1.0) < 4.0 / np.sqrt(nrep * nsteps)
    assert stoch.std() > 2 * mean.std()


def test_stats_only_mode_equals_logged_mode():
    dim, nrep, nsteps = (64, 64), 40, 3000
    st0 = np.stack([exact_state(dim, 0.05, 0.05, r) for r in range(nrep)])
    a = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    b = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    a.upload_state(st0)
    b.upload_state(st0)
    a.run(nsteps, 5, log_mode=nat.LOG_NONE)
    b.run(nsteps, 5, log_mode=nat.LOG_COMPACT, validate=True)
    assert (a.download_state() == b.download_state()).all()
    assert (a.histogram() == b.histogram()).all()
    assert a.histogram().sum() == nrep * nsteps
    a.close()
    b.close()


@pytest.mark.parametrize("dim", [(8, 8), (64, 64), (16, 20), (5, 7), (9, 11), (256, 256), (7, 130), (16, 16),
                                 (128, 32), (512, 4096), (8, 8192), (37, 4096), (5, 4096)])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_sweep_matches_oracle(dim, variant):
    sts = [evolved_state(dim, 3) if dim[0] * dim[1] <= 4096 else iid_state(dim, 3), iid_state(dim, 4)]
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=2)
    eng.set_sweep_kernel(variant)
    eng.upload_state(np.stack(sts))
    eng.sweep()
    slots, counts, rows = eng.slots(), eng.counts(), eng.row_counts()
    for r in range(2):
        ws, wc, wr = reference_rules_only(*ko.sweep(sts[r], dim))
        assert (slots[r] == ws).all()
        assert (counts[r] == wc).all()
        assert (rows[r] == wr).all()
    eng.close()


@pytest.mark.parametrize("dim", [(8, 8), (64, 64), (16, 20), (5, 7), (9, 11), (7, 130), (128, 32), (300, 260), (37, 1024),
                                 (64, 1024), (16, 4096), (37, 4096), (5, 4096), (512, 4096), (512, 512), (256, 16), (8, 8192)])
def test_sweep_with_move_monovacancy_matches_oracle(dim):
    """All nine rules: 29 slot planes, 17 classes, per-row counts."""
    sts = [evolved_state(dim, 3, rates=MM_RATES, order=MM_ORDER) if dim[0] * dim[1] <= 4096 else iid_state(dim, 3),
           iid_state(dim, 4, p=(0.6, 0.15, 0.15, 0.1))]
    eng = Engine(dim, MM_RATES, MM_ORDER, n_replicas=2)
    eng.upload_state(np.stack(sts))
    eng.sweep()
    slots, counts, rows = eng.slots(), eng.counts(), eng.row_counts()
    for r in range(2):
        ws, wc, wr = ko.sweep(sts[r], dim)
        assert (slots[r] == ws).all()
        assert (counts[r] == wc).all()
        assert (rows[r] == wr).all()
        assert wc[13:].sum() > 0
    eng.close()


def test_sweep_full_size_4096():
    """BASELINE config 4 (4096x4096): counts equal the oracle's; checksum of the slot planes agrees."""
    dim = (4096, 4096)
    st = iid_state(dim, 20261019)
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=1)
    eng.upload_state(st)
    eng.sweep()
    eng.sweep()             # twice: the running-total / ticket state must return to zero between launches
    ws, wc, wr = reference_rules_only(*ko.sweep(st, dim))
    assert (eng.counts()[0] == wc).all()
    assert (eng.row_counts()[0] == wr).all()
    slots = eng.slots()[0]
    assert (slots == ws).all()
    eng.close()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("dim", [(64, 64), (16, 4096), (128, 32), (256, 16), (8, 2048)])
def test_sweep_with_move_monovacancy_on_the_forced_aligned_kernels(dim, variant):
    """Sweep variants 1 and 2 with MoveMonovacancy enabled both run the MoveMonovacancy build of the 16-sites-per-thread
    nibble kernel (the byte-parallel variant has none); three replicas, two launches."""
    sts = [iid_state(dim, 11 + r, p=(0.6, 0.15, 0.15, 0.1)) for r in range(3)]
    eng = Engine(dim, MM_RATES, MM_ORDER, n_replicas=3)
    eng.set_sweep_kernel(variant)
    eng.upload_state(np.stack(sts))
    eng.sweep()
    eng.sweep()
    slots, counts, rows = eng.slots(), eng.counts(), eng.row_counts()
    for r in range(3):
        ws, wc, wr = ko.sweep(sts[r], dim)
        assert (slots[r] == ws).all() and (counts[r] == wc).all() and (rows[r] == wr).all()
    eng.close()


@pytest.mark.parametrize("rules", ["reference", "with_move_monovacancy"])
@pytest.mark.parametrize("dim,nrep", [((16, 64), 8), ((64, 16), 4), ((32, 32), 12), ((8, 16), 64), ((16, 16), 48), ((8, 128), 6),
                                      ((32, 64), 2), ((16, 64), 3)])
def test_sweep_of_small_lattices_several_replicas_per_cta(dim, nrep, rules):
    """Lattices of fewer than 4096 sites whose 16-site chunks divide a CTA: one CTA of the nibble kernel holds several
    whole replicas (the last case does not fill whole CTAs and takes the general kernels)."""
    RATES, ORDER = (MM_RATES, MM_ORDER) if rules == "with_move_monovacancy" else (TEST_RATES, ALL_ORDER)
    sts = [iid_state(dim, 70 + r, p=(0.6, 0.15, 0.15, 0.1)) for r in range(nrep)]
    eng = Engine(dim, RATES, ORDER, n_replicas=nrep)
    eng.upload_state(np.stack(sts))
    eng.sweep()
    eng.sweep()
    slots, counts, rows = eng.slots(), eng.counts(), eng.row_counts()
    for r in range(nrep):
        ws, wc, wr = ko.sweep(sts[r], dim)
        if rules == "reference":
            ws, wc, wr = reference_rules_only(ws, wc, wr)
        assert (slots[r] == ws).all() and (counts[r] == wc).all() and (rows[r] == wr).all(), r
    eng.close()


def test_sweep_full_size_4096_every_rule():
    """BASELINE config 4 with config/WSe2.yaml's full rule set: 29 planes, 17 classes in ONE pass -- the rolling-window
    kernel (automatic) and the one-row-per-CTA nibble kernel (sweep variant 1) must both give the oracle's bytes."""
    dim = (4096, 4096)
    st = iid_state(dim, 20261020)
    ws, wc, wr = ko.sweep(st, dim)
    for variant in (0, 1):
        eng = Engine(dim, MM_RATES, MM_ORDER, n_replicas=1)
        eng.set_sweep_kernel(variant)
        eng.upload_state(st)
        eng.sweep()
        eng.sweep()         # twice: the running totals / ticket return to zero between launches
        assert (eng.counts()[0] == wc).all()
        assert (eng.row_counts()[0] == wr).all()
        assert (eng.slots()[0] == ws).all()
        eng.close()


@pytest.mark.parametrize("dim,nsteps", [((64, 64), 300), ((50, 36), 200), ((9, 11), 200), ((300, 260), 60)])
def test_no_incremental_path_matches_oracle(dim, nsteps, monkeypatch):
    """The byte-lattice path (count-only sweep + one-CTA selection); the bit-plane path has its own tests below."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "0")
    st0 = exact_state(dim, 0.1, 0.1, 77)
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=2)
    eng.upload_state(np.stack([st0, st0]))
    ev = eng.run_noinc(nsteps, 9, replica0=0, log_mode=nat.LOG_FULL)
    final = eng.download_state()
    for r in range(2):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0)
        _, want = o.run_philox(9, r, 0, nsteps)
        assert_events_equal(ev[r], want, "noinc replica %d" % r)
        assert (ev[r]["counts"][:, :13] == want["counts"][:, :13]).all() and (ev[r]["counts"][:, 13:] == 0).all()
        assert (final[r] == o.status()).all()
    eng.close()


def test_incremental_and_no_incremental_agree():
    """The reference's A/B check (--no-incremental, sim.py:114-115): same draws, same trajectory."""
    dim = (64, 64)
    st0 = exact_state(dim, 0.05, 0.05, 1)
    a = Engine(dim, TEST_RATES, ALL_ORDER)
    b = Engine(dim, TEST_RATES, ALL_ORDER)
    a.upload_state(st0)
    b.upload_state(st0)
    ea = a.run(400, 31, log_mode=nat.LOG_COMPACT)
    eb = b.run_noinc(400, 31, log_mode=nat.LOG_COMPACT)
    assert (ea == eb).all()
    assert (a.download_state() == b.download_state()).all()
    a.close()
    b.close()


def test_zero_total_rate_raises_value_error():
    rates = [0.0] * 13
    rates[1] = 5.0        # only FillDivacancy, on a lattice with two divacancies
    st = np.zeros(25, np.uint8)
    st[3] = st[17] = 3
    eng = Engine((5, 5), rates, [1])
    eng.upload_state(st)
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.run(10, 1, log_mode=nat.LOG_COMPACT)
    ev = eng._last_events[0]
    assert list(ev["cls"][:3]) == [1, 1, 0xFF]
    assert eng.steps_done()[0] == 2
    assert (eng.download_state() == 0).all()
    eng.close()


@pytest.mark.parametrize("variant", [1, 2, 4, 5])
def test_statistics_only_launch_reports_a_stop_through_check_status(variant):
    """include/kmc_b200.h: a statistics-only kmc_run is asynchronous and returns KMC_OK; the zero-rate stop
    (reference: ValueError, kmc.py:31-32) is reported by the next kmc_check_status, stays raised until the
    lattice is replaced, and does not leak into a run on a new lattice."""
    dim = (8, 128) if variant == 5 else (8, 8)
    n = dim[0] * dim[1]
    rates = [0.0] * 13
    rates[1] = 5.0                                   # FillDivacancy only: the lattice drains
    a = np.zeros(n, np.uint8); a[[3, 17]] = 3
    b = np.zeros(n, np.uint8); b[::2] = 3
    eng = Engine(dim, rates, [1], n_replicas=2)
    eng.set_replica_kernel(variant)
    eng.upload_state(np.stack([a, b]))
    eng.run(20, 7)                                   # KMC_LOG_NONE, no validation: returns once queued
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.check_status()
    with pytest.raises(ValueError, match="replica 0"):
        eng.check_status()                           # still raised
    assert list(eng.steps_done()) == [2, 20]
    # a new lattice starts with a clean status: the old stop must not fail (or short-circuit) this run
    eng.upload_state(np.stack([b, b]))
    ev = eng.run(10, 7, log_mode=nat.LOG_COMPACT)
    eng.check_status()
    assert (ev["cls"] == 1).all()
    assert ((eng.download_state() == 3).sum(axis=1) == n // 2 - 10).all()
    eng.close()


def test_no_incremental_path_after_a_stop_on_an_earlier_lattice(monkeypatch):
    """kmc_noinc_select_apply_kernel returns at entry when the replica's zero-rate flag is set; the
    flag must be cleared with the lattice, and records of steps that never ran must read CLASS_NONE."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "0")          # the byte-lattice path (the plane path: see below)
    dim = (9, 11)
    rates = [0.0] * 13
    rates[1] = 5.0
    a = np.zeros(99, np.uint8); a[[3, 17]] = 3
    b = np.zeros(99, np.uint8); b[::3] = 3
    eng = Engine(dim, rates, [1])
    eng.upload_state(a)
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.run_noinc(6, 5, log_mode=nat.LOG_COMPACT)
    ev = eng._last_events[0]
    assert list(ev["cls"]) == [1, 1, 0xFF, 0xFF, 0xFF, 0xFF]            # no stale records after the stop
    assert (ev["site"][2:] == 0xFFFFFFFF).all()
    eng.upload_state(b)
    ev = eng.run_noinc(6, 5, log_mode=nat.LOG_COMPACT)[0]
    assert (ev["cls"] == 1).all() and (eng.download_state()[0] == 3).sum() == 33 - 6
    o = ko.Oracle(dim, rates, [1], b)
    _, want = o.run_philox(5, 0, 2, 6)               # the step counter kept running: 2 events were done before
    assert_events_equal(ev, want, "noinc after stop")
    eng.close()


def test_event_records_after_a_stop_are_class_none():
    """records past a replica's stop must not carry an earlier launch's data."""
    rates = [0.0] * 13
    rates[1] = 5.0
    full = np.zeros(64, np.uint8); full[::2] = 3
    few = np.zeros(64, np.uint8); few[[1, 9]] = 3
    for variant in (1, 2, 4):
        eng = Engine((8, 8), rates, [1], n_replicas=2)
        eng.set_replica_kernel(variant)
        eng.upload_state(np.stack([full, full]))
        eng.run(16, 3, log_mode=nat.LOG_FULL)                 # fills the device event buffer with real records
        eng.upload_state(np.stack([few, full]))
        with pytest.raises(ValueError):
            eng.run(16, 3, log_mode=nat.LOG_FULL)
        ev = eng._last_events
        assert (ev[0]["cls"][:2] == 1).all() and (ev[0]["cls"][2:] == 0xFF).all(), variant
        assert (ev[0]["site"][2:] == 0xFFFFFFFF).all(), variant
        assert (ev[1]["cls"] == 1).all()
        eng.close()


@pytest.mark.parametrize("variant", [2, 4])
def test_one_replica_of_a_pair_stops_the_other_continues(variant):
    """Two replicas share a warp in the half-warp kernel: when one reaches total rate zero the lockstep
    phase ends and the partner finishes on its own, still bit-equal to the oracle."""
    dim = (8, 8)
    rates = [0.0] * 13
    rates[1] = 5.0                                   # FillDivacancy only: the lattice drains
    a = np.zeros(64, np.uint8); a[[3, 17]] = 3       # two events, then nothing can happen
    b = np.zeros(64, np.uint8); b[::2] = 3           # 32 events available
    c = np.zeros(64, np.uint8); c[[5]] = 3
    st0 = np.stack([a, b, c])
    eng = Engine(dim, rates, [1], n_replicas=3)
    eng.set_replica_kernel(variant)
    eng.upload_state(st0)
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.run(20, 7, log_mode=nat.LOG_FULL, validate=True)
    ev = eng._last_events
    assert list(eng.steps_done()) == [2, 20, 1]
    for r in range(3):
        o = ko.Oracle(dim, rates, [1], st0[r])
        done, want = o.run_philox(7, r, 0, 20)
        assert done == eng.steps_done()[r]
        assert_events_equal(ev[r][:done], want[:done], "replica %d" % r)
        assert ev[r]["cls"][done] == 0xFF if done < 20 else True
    final = eng.download_state()
    assert (final[0] == 0).all() and (final[2] == 0).all() and (final[1] == 3).sum() == 12
    eng.close()


def test_class_frequencies_match_the_reference_with_its_own_rng():
    """Statistical parity: 4096 replicas on the GPU (Philox) against 400 trajectories of the reference run
    with its own random draws (tests/golden/native_stats.npz): per-step class frequencies."""
    import json
    import os
    from scipy.stats import chi2
    from util_golden import GOLDEN, chi2_two_sample
    z = np.load(os.path.join(GOLDEN, "native_stats.npz"))
    m = json.loads(str(z["meta"]))
    nrep, nsteps = 4096, m["nsteps"]
    eng = Engine(m["dim"], z["rates"], m["class_order"], n_replicas=nrep)
    eng.upload_state(np.tile(z["status0"], (nrep, 1)))
    ev = eng.run(nsteps, 987654321, log_mode=nat.LOG_COMPACT, validate=True)
    ours = np.zeros((nsteps, 17), dtype=np.int64)
    for t in range(nsteps):
        ours[t] = np.bincount(ev[:, t]["cls"], minlength=17)
    for t in (0, 1, 5, 10, 20, 39):
        stat, dof = chi2_two_sample(z["counts"][t], ours[t][:13])
        assert chi2.sf(stat, dof) > 1e-4, (t, stat, dof)
    assert (eng.histogram() == ours.sum(axis=0)).all()
    eng.close()


def test_perform_given_move():
    dim = (8, 8)
    st = evolved_state(dim, 5)
    o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st)
    eng = Engine(dim, TEST_RATES, ALL_ORDER)
    eng.upload_state(st)
    slots = o.slots()
    done = set()
    for site, slot in np.argwhere(slots != 0xFF):
        if slot in done:
            continue
        if o.slots()[site, slot] == 0xFF:
            continue
        done.add(slot)
        assert o.perform_given(site, slot) == 0
        eng.perform_given(0, site, slot)
        assert (eng.download_state()[0] == o.status()).all()
    assert len(done) >= 10
    with pytest.raises(RuntimeError, match="does not exist"):
        site = int(np.argwhere(o.status() == 0)[0][0])
        eng.perform_given(0, site, 1)
    eng.close()


def test_invalid_lattice_is_rejected():
    eng = Engine((8, 8), TEST_RATES, ALL_ORDER)
    st = np.zeros(64, np.uint8)
    st[10] = 5                    # trefoil member without its partners
    with pytest.raises(AssertionError, match="trefoil"):
        eng.upload_state(st)
    st[10] = 11
    with pytest.raises(AssertionError, match="status"):
        eng.upload_state(st)
    eng.close()


def test_full_size_replica_batch_properties():
    """BASELINE config 3 shape (64x64, many replicas): size-independent properties at scale --
    histogram total = replicas x steps, incremental tables == fresh enumeration on every replica
    (validate), and a sample of replicas bit-equal to the oracle."""
    dim, nrep, nsteps = (64, 64), 2048, 2000
    rates, order = wse2_rates(1500.0), WSE2_ORDER
    base = [exact_state(dim, 0.05, 0.05, 20261019 + r) for r in range(16)]
    st0 = np.stack([base[r % 16] for r in range(nrep)])
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.upload_state(st0)
    eng.run(nsteps, 20261019, log_mode=nat.LOG_NONE, validate=True)
    assert eng.histogram().sum() == nrep * nsteps
    final = eng.download_state()
    for r in [0, 1, 17, 1000, 2047]:
        o = ko.Oracle(dim, rates, order, st0[r])
        o.run_philox(20261019, r, 0, nsteps, log=False)
        assert (final[r] == o.status()).all()
    eng.close()


def test_headline_batch_at_full_size():
    """BASELINE config 3 exactly as benchmarked: 64x64 WSe2, 16384 replicas (two per warp), 10^4 events each
    in one launch, lattices generated on the device.  Every replica validates its incremental tables against a
    fresh enumeration; totals are conserved; a sample of replicas -- both halves of warps, first and last
    blocks -- is bit-equal to the oracle in lattice, histogram and hop counters."""
    dim, nrep, nsteps, seed = (64, 64), 16384, 10000, 20261019
    rates, order = wse2_rates(1500.0), WSE2_ORDER
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, seed)
    st0 = eng.download_state()
    eng.run(nsteps, seed, log_mode=nat.LOG_NONE, validate=True)
    assert (eng.steps_done() == nsteps).all()
    hist = eng.histogram_per_replica()
    assert (hist.sum(axis=1) == nsteps).all()
    hops = eng.hops()
    assert (hops.sum(axis=1) == hist[:, 2:6].sum(axis=1)).all()
    final = eng.download_state()
    assert ((final == 3).sum(axis=1) + 3 * (final == 4).sum(axis=1) + 3 * (final == 7).sum(axis=1)
            == (st0 == 3).sum(axis=1)).all()              # WSe2 rules conserve divacancies (trefoil = 3 of them)
    for r in [0, 1, 14, 15, 16, 4097, 8190, 8191, 16382, 16383]:
        o = ko.Oracle(dim, rates, order, st0[r])
        _, ev = o.run_philox(seed, r, 0, nsteps)
        assert (final[r] == o.status()).all(), r
        assert (hist[r] == o.hist()).all(), r
        slots = ev["slot"].astype(int)
        assert (hops[r] == np.bincount(slots[(slots >= 7) & (slots < 13)] - 7, minlength=6)).all(), r
    eng.close()


def test_config5_batch_at_per_gpu_size():
    """BASELINE config 5, one GPU's share: 1024x1024 x 128 replicas on the wide bit-plane kernel, lattices
    generated on the device ('approx': 5 % divacancies, 5 % monovacancies), two launches; the row table is
    validated against a fresh one on every replica and a sample of replicas is bit-equal to the oracle."""
    dim, nrep, seed = (1024, 1024), 128, 20261019
    rates, order = wse2_rates(1500.0), WSE2_ORDER
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.generate_state("approx", {"divacancy": 0.05, "monovacancy": 0.05}, seed)
    picks = [0, 77, 127]
    st0 = eng.download_state()[picks].copy()
    eng.run(3000, seed, validate=True)
    eng.run(2000, seed, validate=True)
    assert (eng.steps_done() == 5000).all()
    final = eng.download_state()
    hist = eng.histogram_per_replica()
    for i, r in enumerate(picks):
        o = ko.Oracle(dim, rates, order, st0[i])
        o.run_philox(seed, r, 0, 5000, log=False)
        assert (final[r] == o.status()).all(), r
        assert (hist[r] == o.hist()).all(), r
    eng.close()


# ---- device-side initial-state generators -------------------------------------------
def test_device_state_generator_matches_reference_fixtures():
    """kmc_generate_state against the reference's gen_random_state driven by the same Philox stream
    (tests/golden/state_gen_philox.npz)."""
    import json
    import os
    from util_golden import GOLDEN
    z = np.load(os.path.join(GOLDEN, "state_gen_philox.npz"))
    for i, m in enumerate(json.loads(str(z["meta"]))):
        eng = Engine(m["dim"], TEST_RATES, ALL_ORDER, n_replicas=1)
        eng.generate_state(m["mode"], m["params"], m["seed"], replica0=m["replica"])
        assert (eng.download_state()[0] == z["status_%d" % i]).all(), m
        eng.close()


@pytest.mark.parametrize("mode,dim,params", [
    ("exact", (64, 64), {"divacancy": 0.05, "monovacancy": 0.05}),
    ("exact", (9, 130), {"divacancy": 0.4, "monovacancy": 0.1}),
    ("approx", (64, 64), {"divacancy": 0.05, "monovacancy": 0.05}),
    ("approx", (31, 17), {"divacancy": 0.5, "monovacancy": 0.5}),
    ("exact", (16, 16), {}),
])
def test_device_state_generator_matches_oracle_per_replica(mode, dim, params):
    """Several replicas in one call (replica r = stream (seed, replica0 + r)) against oracle/state_gen.py;
    the generated lattices then run like uploaded ones."""
    import state_gen
    nrep, seed, replica0 = 5, (3 << 32) | 20261019, 1000
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.generate_state(mode, params, seed, replica0=replica0)
    got = eng.download_state()
    for r in range(nrep):
        want = state_gen.generate(dim, mode, params, seed, replica0 + r)
        assert (got[r] == want).all(), (mode, dim, r)
    ev = eng.run(500, 7, log_mode=nat.LOG_COMPACT, validate=True)
    o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, got[2])
    _, want = o.run_philox(7, 2, 0, 500)
    assert_events_equal(ev[2], want, "replica 2 after device generation")
    eng.close()


def test_device_state_generator_exact_counts_at_the_headline_batch():
    """Size-independent property at 64x64 x 4096 replicas: 'exact' puts round(f*N) defects of each type in
    every replica, replicas differ, layers are balanced."""
    nrep, n = 4096, 64 * 64
    eng = Engine((64, 64), wse2_rates(1500.0), WSE2_ORDER, n_replicas=nrep)
    eng.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, 20261019)
    st = eng.download_state()
    assert ((st == 3).sum(axis=1) == round(0.05 * n)).all()
    assert (((st == 1) | (st == 2)).sum(axis=1) == round(0.10 * n) - round(0.05 * n)).all()
    assert st.max() == 3
    assert len({row.tobytes() for row in st[:64]}) == 64
    frac_layer1 = (st == 1).sum() / ((st == 1) | (st == 2)).sum()
    assert abs(frac_layer1 - 0.5) < 0.01
    eng.close()


def test_generate_state_argument_errors():
    eng = Engine((8, 8), TEST_RATES, ALL_ORDER)
    with pytest.raises((RuntimeError, ValueError)):
        nat.check(eng._lib.kmc_generate_state(eng._h, 2, 1, np.zeros(1, np.int32).ctypes.data, None, None, 1, 0))
    with pytest.raises((RuntimeError, ValueError)):
        nat.check(eng._lib.kmc_generate_state(eng._h, 0, 1, np.array([3], np.int32).ctypes.data,
                                              np.array([0, 65], np.int64).ctypes.data, None, 1, 0))
    eng.close()


@pytest.mark.parametrize("variant,dim", [(1, (33, 47)), (2, (64, 64)), (3, (16, 80)), (4, (64, 64)), (5, (24, 128))])
def test_divacancy_hop_counters_match_the_event_log(variant, dim):
    """kmc_get_hops: MoveDivacancy events per direction, accumulated on the device, against the slots of the
    oracle's event log; two launches accumulate, reset_stats clears."""
    nrep, nsteps = 3, 3000
    st0 = np.stack([evolved_state(dim, 60 + r, nsteps=200) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.set_replica_kernel(variant)
    eng.upload_state(st0)
    eng.run(1000, 21)
    eng.run(nsteps - 1000, 21)
    hops = eng.hops()
    disp = eng.displacement()
    nbr = np.array([(-1, 1), (0, -1), (1, 0), (0, 1), (-1, 0), (1, -1)])
    for r in range(nrep):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[r])
        _, want = o.run_philox(21, r, 0, nsteps)
        slots = want["slot"].astype(int)
        w = np.bincount(slots[(slots >= 7) & (slots < 13)] - 7, minlength=6)
        assert (hops[r] == w).all(), (variant, r)
        assert (disp[r] == w @ nbr).all()
        assert w.sum() == o.hist()[2:6].sum()
    eng.reset_stats()
    assert (eng.hops() == 0).all()
    eng.close()


def test_hop_counters_in_the_no_incremental_path():
    dim = (16, 16)
    st0 = evolved_state(dim, 8, nsteps=100)
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=1)
    eng.upload_state(st0)
    ev = eng.run_noinc(200, 4, log_mode=nat.LOG_COMPACT)[0]
    slots = ev["slot"].astype(int)
    assert (eng.hops()[0] == np.bincount(slots[(slots >= 7) & (slots < 13)] - 7, minlength=6)).all()
    eng.close()


# ---- Zobrist key -------------------------------------------------------------------
@pytest.mark.parametrize("dim", [(64, 64), (33, 47), (24, 128)])
def test_device_zobrist_key_matches_oracle(dim):
    import zobrist as oz
    nrep = 4
    st0 = np.stack([evolved_state(dim, 70 + r, nsteps=400) for r in range(nrep)])
    assert (st0 >= 4).any()                                  # trefoils present
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.upload_state(st0)
    for bits, seed in ((64, 0), (32, 123456789012345), (17, 9)):
        got = eng.zobrist(bits, seed)
        for r in range(nrep):
            assert int(got[r]) == oz.key(st0[r], bits, seed), (bits, r)
    # after a run (wide kernel: planes are authoritative) the key follows the lattice
    eng.run(500, 3)
    final = eng.download_state()
    got = eng.zobrist(64, 1)
    for r in range(nrep):
        assert int(got[r]) == oz.key(final[r], 64, 1)
    eng.close()


def test_million_step_single_trajectory_is_bit_exact():
    """BASELINE config 2 at its full length: 64x64 WSe2 at 1500 K, one trajectory, 10^6 events in launches of
    2^17, every event (class, site, slot, sequential total rate) and the final lattice against the C oracle."""
    dim, nsteps, seed = (64, 64), 1_000_000, 20261019
    rates = wse2_rates(1500.0)
    st0 = exact_state(dim, 0.05, 0.05, seed)
    o = ko.Oracle(dim, rates, WSE2_ORDER, st0)
    done, want = o.run_philox(seed, 0, 0, nsteps)
    assert done == nsteps
    eng = Engine(dim, rates, WSE2_ORDER, n_replicas=1)
    eng.upload_state(st0)
    pos = 0
    while pos < nsteps:
        n = min(1 << 17, nsteps - pos)
        ev = eng.run(n, seed, log_mode=nat.LOG_COMPACT, validate=True)[0]
        w = want[pos:pos + n]
        assert (ev["site"] == w["site"]).all() and (ev["cls"] == w["cls"]).all() and (ev["slot"] == w["slot"]).all()
        assert (ev["total_rate"].view(np.uint64) == w["total_naive"].view(np.uint64)).all()
        pos += n
    assert (eng.download_state()[0] == o.status()).all()
    assert (eng.histogram_per_replica()[0] == o.hist()).all()
    eng.close()


@pytest.mark.parametrize("variant,dim", [(1, (16, 20)), (2, (33, 47)), (3, (16, 80)), (4, (64, 64)), (4, (9, 64)),
                                         (5, (12, 128))])
def test_injected_draws_on_every_replica_kernel(variant, dim):
    """kmc_step_draws (caller-supplied uniforms, the parity seam of SURVEY 8(c)) on every kernel variant, three
    replicas with different draw streams, two calls."""
    nrep, nsteps = 3, 1500
    st0 = np.stack([evolved_state(dim, 90 + r, nsteps=200) for r in range(nrep)])
    d = np.stack([philox.draws(1234, 50 + r, 0, nsteps) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.set_replica_kernel(variant)
    eng.upload_state(st0)
    ev = np.concatenate([eng.step_draws(d[:, :700], log_mode=nat.LOG_COMPACT, validate=True),
                         eng.step_draws(d[:, 700:], log_mode=nat.LOG_COMPACT, validate=True)], axis=1)
    final = eng.download_state()
    for r in range(nrep):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[r])
        _, want = o.run_draws(d[r])
        assert_events_equal(ev[r], want, "variant %d replica %d" % (variant, r))
        assert (final[r] == o.status()).all()
    eng.close()


# ---- MoveMonovacancy (a stub in the reference, semantics in DESIGN.md section 2) ------------------
MM_FIXTURES = ["wse2_full_16x16", "wse2_full_64x64", "wse2_full_hot_24x24", "allrules_mm_8x8", "allrules_mm_5x7",
               "allrules_mm_64x64", "allrules_mm_12x130"]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, "noinc", "draws"])
@pytest.mark.parametrize("name", MM_FIXTURES)
def test_move_monovacancy_fixtures_on_the_general_kernels(name, variant):
    """Fixtures made by the REFERENCE engine running oracle/move_monovacancy.py's rule class (all of
    config/WSe2.yaml's rules; all nine rules at comparable rates): byte-lattice kernel with the lattice in shared
    memory (1) and in HBM (3), the no-incremental path, and the draw-injection seam."""
    g = load(name)
    m = g["meta"]
    if variant == 4 and (m["dim"][1] > 64 or len(m["class_order"]) == 17):
        pytest.skip("half-warp kernel: rows of <= 64 sites and at most 16 enabled classes")
    if variant == 5 and m["dim"][1] <= 64:
        pytest.skip("wide bit-plane kernel: rows wider than 64 sites")
    if variant == 2 and m["dim"][1] > 64:
        pytest.skip("warp-per-replica bit-plane kernel: rows of <= 64 sites")
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.upload_state(g["status0"])
    n = m["nsteps"] if variant != "noinc" else min(m["nsteps"], 400)
    if variant == "noinc":
        ev = eng.run_noinc(n, m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL)[0]
    elif variant == "draws":
        d = philox.draws(m["seed"], m["replica"], 0, n)
        ev = np.concatenate([eng.step_draws(d[:n // 3], log_mode=nat.LOG_FULL, validate=True)[0],
                             eng.step_draws(d[n // 3:], log_mode=nat.LOG_FULL, validate=True)[0]])
    else:
        eng.set_replica_kernel(variant)
        ev = np.concatenate([eng.run(n // 2, m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL, validate=True)[0],
                             eng.run(n - n // 2, m["seed"], replica0=m["replica"], log_mode=nat.LOG_FULL, validate=True)[0]])
    assert_events_equal(ev, g["events"][:n], name)
    got = fsum_totals(ev, g["rates"], m["class_order"])
    assert (got.view(np.uint64) == g["events"]["total_rate"][:n].view(np.uint64)).all()
    if n == m["nsteps"]:
        assert (eng.download_state()[0] == g["status1"]).all()
        eng.sweep()
        en = enabled_classes(m)
        assert (eng.counts()[0][en] == g["counts1"][en]).all()
        slots = eng.slots()[0]
        assert (np.where(np.isin(slots, en), slots, 0xFF) == g["slots1"]).all()
    assert (ev["cls"] >= 13).sum() > 40
    eng.close()


# sixteen live classes: everything but FillDivacancy (the half-warp kernel has 16 lanes for 17 classes)
MM16_ORDER = [c for c in MM_ORDER if c != 1]
MM16B_ORDER = [16, 15, 14, 13] + [c for c in range(13) if c != 9]          # another order, another lane for class 16


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dim,rates,order", [((64, 64), MM_RATES, MM_ORDER), ((33, 47), MM_RATES, MM_ORDER),
                                             ((5, 5), MM_RATES, MM_ORDER), ((7, 130), MM_RATES, MM_ORDER),
                                             ((24, 200), MM_RATES, MM_ORDER), ((130, 128), wse2_full_rates(5000.0), WSE2_FULL_ORDER),
                                             ((9, 65), MM_RATES, MM16_ORDER), ((40, 2044), MM_RATES, MM_ORDER),
                                             ((64, 64), wse2_full_rates(1500.0), WSE2_FULL_ORDER),
                                             ((24, 40), wse2_full_rates(5000.0), WSE2_FULL_ORDER),
                                             ((64, 64), MM_RATES, MM16_ORDER), ((33, 47), MM_RATES, MM16B_ORDER),
                                             ((5, 5), MM_RATES, MM16_ORDER), ((9, 64), MM_RATES, MM16B_ORDER)])
def test_move_monovacancy_replicas_match_oracle(dim, rates, order, variant):
    """Several replicas, long runs, two launches, with the clock and the divacancy tracers on (a swap carries a
    tracer along, a merge starts one, a split ends one), against the C oracle."""
    nrep, nsteps, seed = 3, 6000, 515
    if variant == 4 and (dim[1] > 64 or len(order) == 17):
        pytest.skip("half-warp kernel: rows of <= 64 sites and at most 16 enabled classes")
    if variant == 5 and dim[1] <= 64:
        pytest.skip("wide bit-plane kernel: rows wider than 64 sites")
    if variant == 2 and dim[1] > 64:
        pytest.skip("warp-per-replica bit-plane kernel: rows of <= 64 sites")
    st0 = np.stack([exact_state(dim, 0.12 + 0.03 * r, 0.10, 300 + r) for r in range(nrep)])
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.set_replica_kernel(variant)
    eng.enable_clock("stochastic")
    eng.enable_tracers()
    eng.upload_state(st0)
    ev = np.concatenate([eng.run(2500, seed, replica0=4, log_mode=nat.LOG_FULL, validate=True),
                         eng.run(nsteps - 2500, seed, replica0=4, log_mode=nat.LOG_FULL, validate=True)], axis=1)
    final, hist, t = eng.download_state(), eng.histogram_per_replica(), eng.clock()
    sq, cnt = eng.msd()
    for r in range(nrep):
        o = ko.Oracle(dim, rates, order, st0[r])
        o.enable_tracers()
        o.set_clock(2)
        done, want = o.run_philox(seed, 4 + r, 0, nsteps)
        assert done == nsteps
        assert_events_equal(ev[r], want, "replica %d" % r)
        assert (ev[r]["total_rate"].view(np.uint64) == want["total_naive"].view(np.uint64)).all()
        assert (ev[r]["counts"][:, order] == want["counts"][:, order]).all()      # (the reference counts enabled rules)
        assert (final[r] == o.status()).all()
        assert (hist[r] == o.hist()).all()
        assert t[r] == o.clock()
        div = final[r] == 3
        assert (eng.tracers(r)[div] == o.tracers()[div]).all()
        assert (int(sq[r]), int(cnt[r])) == o.msd()
    assert hist[:, 13:].sum() > 1000
    eng.close()


def test_kernels_that_cannot_take_move_monovacancy_decline():
    eng = Engine((64, 64), MM_RATES, MM_ORDER)             # all 17 classes: no free lane in the half-warp kernel
    eng.set_replica_kernel(4)
    with pytest.raises(RuntimeError, match="MoveMonovacancy is enabled"):
        eng.run(10, 1)
    eng.set_replica_kernel(2)                              # (the warp-per-replica kernel has 32 weight lanes)
    eng.run(10, 1, validate=True)
    eng.close()
    eng = Engine((8, 128), MM_RATES, MM16_ORDER)           # rows wider than 64 sites: not the half-warp kernel
    eng.set_replica_kernel(4)
    with pytest.raises(RuntimeError, match="MoveMonovacancy is enabled"):
        eng.run(10, 1)
    for variant in (0, 1, 5):
        eng.set_replica_kernel(variant)
        eng.run(10, 1, validate=True)
    eng.close()


def test_full_wse2_batch_on_the_half_warp_kernel():
    """config/WSe2.yaml with every rule (MoveMonovacancy live) at the benchmarked shape, statistics-only build,
    both halves of warps and several blocks: validation on every replica, a sample bit-equal to the oracle."""
    dim, nrep, nsteps, seed = (64, 64), 2048, 4000, 20261019
    rates, order = wse2_full_rates(1500.0), WSE2_FULL_ORDER
    eng = Engine(dim, rates, order, n_replicas=nrep)
    eng.generate_state("exact", {"divacancy": 0.05, "monovacancy": 0.05}, seed)
    st0 = eng.download_state()
    eng.run(nsteps, seed)                                  # PLAIN build
    eng.check_status()
    eng.run(100, seed, validate=True)                      # general build, validates what the PLAIN build left
    hist = eng.histogram_per_replica()
    assert (hist.sum(axis=1) == nsteps + 100).all()
    final = eng.download_state()
    for r in [0, 1, 15, 16, 1000, 2046, 2047]:
        o = ko.Oracle(dim, rates, order, st0[r])
        o.run_philox(seed, r, 0, nsteps + 100, log=False)
        assert (final[r] == o.status()).all(), r
        assert (hist[r] == o.hist()).all(), r
    assert hist[:, 13].sum() > 0.8 * hist.sum()            # the monovacancies do the moving
    eng.close()


def test_move_monovacancy_incremental_equals_no_incremental_and_perform_given():
    dim = (20, 36)
    st0 = exact_state(dim, 0.15, 0.15, 8)
    a = Engine(dim, MM_RATES, MM_ORDER)
    b = Engine(dim, MM_RATES, MM_ORDER)
    a.upload_state(st0)
    b.upload_state(st0)
    ea = a.run(300, 21, log_mode=nat.LOG_COMPACT, validate=True)[0]
    eb = b.run_noinc(300, 21, log_mode=nat.LOG_COMPACT)[0]
    assert (ea == eb).all()
    assert (a.download_state() == b.download_state()).all()
    # explicit moves: one of every MoveMonovacancy slot that exists on the lattice
    o = ko.Oracle(dim, MM_RATES, MM_ORDER, a.download_state()[0])
    done = set()
    for site, slot in np.argwhere(o.slots() != 0xFF):
        if slot < 17 or slot in done or o.slots()[site, slot] == 0xFF:
            continue
        done.add(slot)
        assert o.perform_given(site, slot) == 0
        a.perform_given(0, site, slot)
        assert (a.download_state()[0] == o.status()).all()
    assert len(done) == 12
    with pytest.raises(RuntimeError, match="does not exist"):
        a.perform_given(0, int(np.flatnonzero(o.status() == 0)[0]), 17)
    a.close()
    b.close()


# ---- the no-incremental event on the bit-sliced lattice (noinc_planes_kernel.cuh) ---------------------------------
@pytest.mark.parametrize("rules", ["reference", "with_move_monovacancy"])
@pytest.mark.parametrize("dim,nsteps", [((64, 64), 300), ((50, 36), 200), ((9, 11), 200), ((300, 260), 120), ((7, 130), 150),
                                        ((33, 2100), 60), ((5, 64), 100), ((130, 65), 100), ((512, 4096), 40),
                                        ((7, 8192), 40), ((5, 16384), 40)])
def test_no_incremental_on_planes_matches_oracle(dim, nsteps, rules, monkeypatch):
    """Large lattices take the plane path by themselves (>= 65536 sites); KMC_NOINC_PLANES=1 forces it for the
    small and ragged shapes too.  Trefoils present (evolved lattice), two replicas, two launches, hop counters,
    stochastic clock and tracers on -- everything against the oracle; then the byte-lattice paths take over from
    the planes (lazy unpack) and the two no-incremental paths agree with each other."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    RATES, ORDER = (MM_RATES, MM_ORDER) if rules == "with_move_monovacancy" else (TEST_RATES, ALL_ORDER)
    ncmp = 17 if rules == "with_move_monovacancy" else 13
    big = dim[0] * dim[1] > 100000
    st0 = iid_state(dim, 77, p=(0.80, 0.05, 0.05, 0.10)) if big else evolved_state(dim, 31, nsteps=300)
    st1 = iid_state(dim, 78, p=(0.70, 0.05, 0.05, 0.20)) if big else exact_state(dim, 0.2, 0.1, 5)
    eng = Engine(dim, RATES, ORDER, n_replicas=2)
    eng.enable_clock("stochastic")
    eng.enable_tracers()
    eng.upload_state(np.stack([st0, st1]))
    n1 = nsteps // 3
    ev = np.concatenate([eng.run_noinc(n1, 9, replica0=3, log_mode=nat.LOG_FULL),
                         eng.run_noinc(nsteps - n1, 9, replica0=3, log_mode=nat.LOG_FULL)], axis=1)
    final, t, hops = eng.download_state(), eng.clock(), eng.hops()
    sq, cnt = eng.msd()
    for r, st in enumerate((st0, st1)):
        o = ko.Oracle(dim, RATES, ORDER, st)
        o.enable_tracers()
        o.set_clock(2)
        done, want = o.run_philox(9, 3 + r, 0, nsteps)
        assert done == nsteps
        assert_events_equal(ev[r], want, "planes noinc replica %d" % r)
        assert (ev[r]["total_rate"].view(np.uint64) == want["total_naive"].view(np.uint64)).all()
        assert (ev[r]["counts"][:, :ncmp] == want["counts"][:, :ncmp]).all()
        assert (final[r] == o.status()).all()
        assert t[r] == o.clock()
        slots = want["slot"].astype(int)
        assert (hops[r] == np.bincount(slots[(slots >= 7) & (slots < 13)] - 7, minlength=6)).all()
        assert (int(sq[r]), int(cnt[r])) == o.msd()
    # hand over to the incremental kernels and back
    ev2 = eng.run(50, 9, replica0=3, log_mode=nat.LOG_COMPACT, validate=True)
    monkeypatch.setenv("KMC_NOINC_PLANES", "0")
    ev3 = eng.run_noinc(20, 9, replica0=3, log_mode=nat.LOG_COMPACT)
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    ev4 = eng.run_noinc(20, 9, replica0=3, log_mode=nat.LOG_COMPACT)
    for r, st in enumerate((st0, st1)):
        o = ko.Oracle(dim, RATES, ORDER, st)
        _, want = o.run_philox(9, 3 + r, 0, nsteps + 90)
        assert_events_equal(np.concatenate([ev2[r], ev3[r], ev4[r]]), want[nsteps:], "hand-over %d" % r)
        assert (eng.download_state()[r] == o.status()).all()
    eng.close()


@pytest.mark.parametrize("launch", ["cooperative", "per_event_pdl", "per_event_plain"])
@pytest.mark.parametrize("dim", [(300, 260), (512, 4096)])
def test_plane_no_incremental_launch_forms_agree(dim, launch, monkeypatch):
    """The same event kernel runs as one cooperative launch per call (epoch hand-over between events) or as one launch
    per event, with or without programmatic dependent launch: same records, same lattices, three replicas."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    if launch != "cooperative":
        monkeypatch.setenv("KMC_NOINC_NO_PERSIST", "1")
    if launch == "per_event_plain":
        monkeypatch.setenv("KMC_NOINC_NO_PDL", "1")
    nsteps = 60
    sts = np.stack([iid_state(dim, 40 + r, p=(0.80, 0.05, 0.05, 0.10)) for r in range(3)])
    eng = Engine(dim, MM_RATES, MM_ORDER, n_replicas=3)
    eng.upload_state(sts)
    ev = np.concatenate([eng.run_noinc(1, 13, log_mode=nat.LOG_COMPACT),          # (a one-event call is never cooperative)
                         eng.run_noinc(nsteps - 1, 13, log_mode=nat.LOG_COMPACT)], axis=1)
    final = eng.download_state()
    for r in range(3):
        o = ko.Oracle(dim, MM_RATES, MM_ORDER, sts[r])
        _, want = o.run_philox(13, r, 0, nsteps)
        assert_events_equal(ev[r], want, "%s replica %d" % (launch, r))
        assert (final[r] == o.status()).all()
    assert (eng.steps_done() == nsteps).all()
    eng.close()


def test_plane_no_incremental_batch_too_large_for_one_cooperative_launch(monkeypatch):
    """48 replicas of 300 x 260 are 1824 CTAs -- more than fit on the device at once -- so the call falls back to one
    launch per event by itself; a sample of replicas against the oracle."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    dim, nrep, nsteps = (300, 260), 48, 40
    sts = np.stack([iid_state(dim, 900 + r, p=(0.80, 0.05, 0.05, 0.10)) for r in range(nrep)])
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=nrep)
    eng.upload_state(sts)
    l0 = eng.launch_count()
    ev = eng.run_noinc(nsteps, 17, log_mode=nat.LOG_COMPACT)
    assert eng.launch_count() - l0 >= nsteps            # (one kernel per event, plus the pack of the planes)
    final = eng.download_state()
    for r in (0, 1, 23, 47):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, sts[r])
        _, want = o.run_philox(17, r, 0, nsteps)
        assert_events_equal(ev[r], want, "replica %d" % r)
        assert (final[r] == o.status()).all()
    eng.close()


@pytest.mark.parametrize("name", boundary_tie_names())
def test_boundary_ties_on_the_plane_no_incremental_path(name, monkeypatch):
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    g = load(name)
    m = g["meta"]
    eng = Engine(m["dim"], g["rates"], m["class_order"], n_replicas=1)
    eng.upload_state(g["status0"])
    ev = eng.step_draws_noinc(g["draws"], log_mode=nat.LOG_FULL)[0]
    assert_events_equal(ev, g["events"], name)
    assert np.nonzero(ev["flags"] & 1)[0].tolist() == m["tie_steps"] and eng.ties() == m["n_ties"]
    assert (eng.download_state()[0] == g["status1"]).all()
    eng.close()


def test_plane_no_incremental_path_stops_on_zero_rate(monkeypatch):
    monkeypatch.setenv("KMC_NOINC_PLANES", "1")
    dim = (9, 70)
    rates = [0.0] * 13
    rates[1] = 5.0
    a = np.zeros(630, np.uint8); a[[3, 417]] = 3
    eng = Engine(dim, rates, [1])
    eng.upload_state(a)
    with pytest.raises(ValueError, match="total weight of zero"):
        eng.run_noinc(6, 5, log_mode=nat.LOG_COMPACT)
    ev = eng._last_events[0]
    assert list(ev["cls"]) == [1, 1, 0xFF, 0xFF, 0xFF, 0xFF]
    assert (eng.download_state() == 0).all() and eng.steps_done()[0] == 2
    eng.close()


@pytest.mark.parametrize("dim", [(64, 64), (16, 4096), (128, 32)])
def test_no_incremental_with_move_monovacancy_on_aligned_shapes(dim, monkeypatch):
    """Byte-lattice path: aligned shapes enumerate with the MoveMonovacancy builds of the nibble sweep kernels, whose
    MoveMonovacancy row counts live in their own table: the no-incremental selection must read both."""
    monkeypatch.setenv("KMC_NOINC_PLANES", "0")
    st0 = exact_state(dim, 0.15, 0.15, 8)
    eng = Engine(dim, MM_RATES, MM_ORDER)
    eng.upload_state(st0)
    ev = eng.run_noinc(150, 21, log_mode=nat.LOG_FULL)[0]
    o = ko.Oracle(dim, MM_RATES, MM_ORDER, st0)
    _, want = o.run_philox(21, 0, 0, 150)
    assert_events_equal(ev, want, "noinc mm")
    assert (ev["counts"] == want["counts"]).all()
    assert (eng.download_state()[0] == o.status()).all()
    assert (ev["cls"] >= 13).sum() > 20
    eng.close()
