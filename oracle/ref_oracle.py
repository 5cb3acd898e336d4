"""Reference-driven, draw-injected KMC stepper (TEST INFRASTRUCTURE ONLY).

Runs the UNMODIFIED reference (``/root/reference/demo1``) one event at a time while injecting
the two uniform draws each event consumes, so that its trajectory becomes reproducible and can
pin the CPU restatement (``oracle/kmc_oracle.c``) and the CUDA engine.  It exists only in the
build container (``/root/reference`` does not travel to the GPU box); its outputs are committed
as fixtures under ``tests/golden/`` by ``oracle/make_golden.py``.

What is the reference's and what is ours:

* counts per (rule, kind)     reference: ``IncrementalMoveCache.undecided_counts`` (incremental.py:103)
* weights, class choice       reference: ``weighted_choice`` itself (kmc.py:21-54) with an injected
                              ``rng`` whose ``uniform(a, b)`` returns ``a + (b - a) * u1`` (the formula
                              of CPython's ``random.uniform``)
* class order fed to it       OURS: canonical = rules in config order x ``rule.kinds()`` order
                              (the reference's order is Counter-insertion history, sim.py:101)
* in-class pick               OURS: ``sorted(moves, key=(row, slot, column))[min(int(u2 * n), n - 1)]``
                              (the reference's list order is a swap-remove artefact, incremental.py:306)
* mutation + refresh          reference: ``KmcSim.perform_given_move`` (sim.py:145-157)
* total_rate                  reference formula: ``math.fsum`` of the weights (sim.py:138)
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("KMC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "demo1"))


def _import_reference():
    """Import the reference's modules (needs the compat shim first on sys.path)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True          # the reference tree is read-only
    shim = os.path.join(HERE, "ref_shim")
    for p in (REFERENCE_ROOT, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sitecustomize  # noqa: F401  (the shim; idempotent)
    import importlib
    importlib.reload(sitecustomize)
    # our product package is also called ``demo1``: make sure we get the reference's
    mods = {}
    saved = {k: v for k, v in sys.modules.items() if k == "demo1" or k.startswith("demo1.")}
    for k in saved:
        del sys.modules[k]
    try:
        for name in ("sim", "kmc", "rules", "state", "incremental", "config"):
            mods[name] = importlib.import_module("demo1." + name)
        assert os.path.abspath(mods["sim"].__file__).startswith(os.path.abspath(REFERENCE_ROOT))
        mods["_sys_modules"] = {k: v for k, v in sys.modules.items()
                                if k == "demo1" or k.startswith("demo1.")}
        # MoveMonovacancy is a stub in the reference (rules.py:295-296): our stated semantics, as a rule class of
        # the reference's own framework, take its place on the module object (oracle/move_monovacancy.py)
        import move_monovacancy
        move_monovacancy.install(mods)
    finally:
        # leave the reference's modules reachable only through ``mods`` / ``ref_modules()``
        for k in [k for k in sys.modules if k == "demo1" or k.startswith("demo1.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mods


class ref_modules:
    """Context manager: while active, ``sys.modules['demo1*']`` are the reference's modules.

    The reference clones its State through pickle (state.py:68-70), which resolves classes by
    module name; the product's host package is *also* called ``demo1``, so the reference's
    modules are swapped in only for the duration of a call into it."""

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items() if k == "demo1" or k.startswith("demo1.")}
        for k in self.saved:
            del sys.modules[k]
        sys.modules.update(ref()["_sys_modules"])

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == "demo1" or k.startswith("demo1.")]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


_REF = None


def ref():
    global _REF
    if _REF is None:
        _REF = _import_reference()
    return _REF


sys.path.insert(0, HERE)
import canon  # noqa: E402


class _FixedUniform:
    """``rng`` seam of weighted_choice (kmc.py:21): hands back one injected uniform."""

    def __init__(self, u):
        self.u = u
        self.x = None

    def uniform(self, a, b):
        self.x = a + (b - a) * self.u      # CPython's random.uniform formula
        return self.x


def state_from_status(status, dim, zobrist=None):
    """Build a reference ``State`` from a flat uint8 status array (canon status codes)."""
    st = ref()["state"]
    s = st.State(tuple(dim), zobrist=zobrist)
    for i, code in enumerate(status):
        code = int(code)
        node = canon.unlin(i, dim)
        if 1 <= code <= 3:
            s.new_vacancy(node, code)
        elif code >= canon.TREFOIL_BASE:
            orient, role = divmod(code - canon.TREFOIL_BASE, 3)
            if role == 0:
                s.new_trefoil(canon.trefoil_nodes(node, orient, dim))
    s.validate()
    return s


def status_from_state(state):
    dim = tuple(state.dim())
    out = [0] * (dim[0] * dim[1])
    for v in state.vacancies():
        out[canon.lin(v.node, dim)] = v.layers
    for t in state.trefoils():
        anchor, orient = canon.trefoil_anchor(t.nodes, dim)
        for role, n in enumerate(canon.trefoil_nodes(anchor, orient, dim)):
            out[canon.lin(n, dim)] = canon.TREFOIL_BASE + 3 * orient + role
    return out


def make_rule_specs(rules_cfg, order=None):
    """``rules_cfg``: {RuleName: {'rate': x | {kind: x}}} or {'barrier': ...}; returns RuleSpecs
    built by the reference's own config code (config.py:23-74), in ``order`` (default: dict order)."""
    import copy
    cfg = ref()["config"]
    d = {"rules": copy.deepcopy({k: rules_cfg[k] for k in (order or list(rules_cfg))})}
    return cfg.consume__rule_specs(d)


class RefStepper:
    """One reference ``KmcSim`` driven by injected draws."""

    def __init__(self, status, dim, rules_cfg, temperature=None, rule_order=None, zobrist=None):
        self.dim = tuple(dim)
        self.mods = ref()
        specs = make_rule_specs(rules_cfg, rule_order)
        init = state_from_status(status, self.dim, zobrist=zobrist)
        with ref_modules():
            self.sim = self.mods["sim"].KmcSim(init, specs, temperature=temperature,
                                               incremental=True)
        # canonical class list: rules in config order x kinds() order
        self.pairs = []
        for rule in self.sim.rules_and_rates:       # dict preserves spec order (sim.py:57-68)
            for kind in rule.kinds():
                self.pairs.append((rule, kind))
        self.n_ties = 0

    # -- views ---------------------------------------------------------------------------
    def class_order(self):
        return [canon.CLASS_ID[(r.format(), k)] for (r, k) in self.pairs]

    def rates(self):
        out = [0.0] * canon.N_CLASSES
        for (r, k) in self.pairs:
            out[canon.CLASS_ID[(r.format(), k)]] = self.sim.rate(r, k)
        return out

    def counts(self):
        out = [0] * canon.N_CLASSES
        for rule in self.sim.rules_and_rates:
            for kindset, n in rule.move_cache.undecided_counts().items():
                assert len(kindset) == 1, "multi-kind move: not expected in shipped rules"
                (kind,) = kindset
                out[canon.CLASS_ID[(rule.format(), kind)]] += n
        return out

    def status(self):
        return status_from_state(self.sim.state)

    def slot_table(self):
        """[site][slot] -> class id or SLOT_NONE, from the reference's incremental move caches."""
        n = self.dim[0] * self.dim[1]
        tab = [[canon.SLOT_NONE] * canon.N_SLOTS for _ in range(n)]
        for rule in self.sim.rules_and_rates:
            for kindset, moves in rule.move_cache.undecided_sets().items():
                (kind,) = kindset
                cid = canon.CLASS_ID[(rule.format(), kind)]
                for mv in moves:
                    site, slot = canon.slot_from_move(rule.format(), mv, self.dim)
                    assert tab[site][slot] == canon.SLOT_NONE
                    tab[site][slot] = cid
        return tab

    # -- one event -----------------------------------------------------------------------
    def step(self, u1, u2):
        sim = self.sim
        counts = {}
        for rule in sim.rules_and_rates:
            for kindset, n in rule.move_cache.undecided_counts().items():
                (kind,) = kindset
                counts[(rule, kind)] = n
        k_w_pairs = [((r, k), counts.get((r, k), 0) * sim.rate(r, k)) for (r, k) in self.pairs]
        rng = _FixedUniform(u1)
        (rule, kind) = self.mods["kmc"].weighted_choice(k_w_pairs, rng=rng)

        moves = rule.move_cache.undecided_sets()[frozenset([kind])]
        d1 = self.dim[1]
        keyed = []
        for m in moves:
            site, slot = canon.slot_from_move(rule.format(), m, self.dim)
            keyed.append(((site // d1, slot, site % d1), site, slot, m))     # (row, slot, column) order
        keyed.sort(key=lambda t: t[0])
        n = len(keyed)
        assert n == counts[(rule, kind)]
        idx = min(int(u2 * n), n - 1)
        _, site, slot, move = keyed[idx]

        # boundary-tie bookkeeping: x landed exactly on a cumulative weight
        ws = sorted(w for (_, w) in k_w_pairs if w != 0.0)
        acc = 0
        for w in ws:
            acc = acc + w
            if acc == rng.x:
                self.n_ties += 1
        info = rule.info(move)
        sim.perform_given_move(rule, move)
        return {
            "class": canon.CLASS_ID[(rule.format(), kind)],
            "site": site,
            "slot": slot,
            "total_rate": math.fsum(w for (_, w) in k_w_pairs),
            "rule": rule.format(),
            "kind": kind,
            "info": info,
        }

    def validate(self):
        return self.sim.validate()
