"""Integer tables of the hot path: event classes, per-site move slots, stencils, status codes.

These restate, as constants, what the reference derives at run time (paths relative to the
reference root): rule/kind names from demo1/rules.py:60-293, stencils from demo1/state.py:392-401,
443-459.  The same numbering is used by the CUDA kernels (csrc/kmc_common.cuh) and documented in
include/kmc_b200.h.
"""

DEFAULT_KIND = "natural"            # reference demo1/sim.py:22

# status byte of a site: 0..3 = vacant layer set (reference state.py:22-23,171-177), >= 4 trefoil
PRISTINE, MONO1, MONO2, DIVACANCY = 0, 1, 2, 3
TREFOIL_BASE = 4                    # 4 + 3*orientation + role

# (rule, kind) -> class id
CLASSES = [
    ("CreateDivacancy", "natural"), ("FillDivacancy", "natural"),
    ("MoveDivacancy", "natural"), ("MoveDivacancy", "missing-metal"),
    ("MoveDivacancy", "missing-comb"), ("MoveDivacancy", "missing-both"),
    ("CreateTrefoil", "natural"), ("DestroyTrefoil", "natural"),
    ("CreateMonovacancy", "from-double"), ("CreateMonovacancy", "make-empty"),
    ("FillMonovacancy", "make-double"), ("FillMonovacancy", "from-empty"),
    ("FlipMonovacancy", "natural"),
    # MoveMonovacancy: named with four kinds and barriers in the reference's config/WSe2.yaml:8-13 but only a
    # stub there (rules.py:295-296).  Extension with stated semantics: the chalcogen
    # vacancy in one layer of a node hops to a neighbouring node whose layer set lacks that layer;
    # kind = [natural, merge-divacancy, split-divacancy, swap-divacancy][2 * (source is a divacancy)
    #                                                                    + (target is not pristine)]
    ("MoveMonovacancy", "natural"), ("MoveMonovacancy", "merge-divacancy"),
    ("MoveMonovacancy", "split-divacancy"), ("MoveMonovacancy", "swap-divacancy"),
]
CLASS_ID = {rk: i for i, rk in enumerate(CLASSES)}
N_CLASSES = len(CLASSES)

# rule name -> kinds in Rule.kinds() order (rules.py:108-109,241,260; default sim.py:262-265)
RULE_KINDS = {}
for _r, _k in CLASSES:
    RULE_KINDS.setdefault(_r, []).append(_k)
IMPLEMENTED_RULES = list(RULE_KINDS)
STUB_RULES = {}                     # (round 1: MoveMonovacancy was skipped with a warning)
# classes the reference implements itself; 13.. are the MoveMonovacancy extension
N_REFERENCE_CLASSES = 13

N_SLOTS = 29
MM_SLOT0 = 17                       # MoveMonovacancy (node, nbr_k, layer): slot 17 + 2*k + (layer - 1)
SLOT_RULE = (["CreateDivacancy", "FillDivacancy"] + ["CreateMonovacancy"] * 2 + ["FillMonovacancy"] * 2 +
             ["FlipMonovacancy"] + ["MoveDivacancy"] * 6 + ["CreateTrefoil"] * 2 + ["DestroyTrefoil"] * 2 +
             ["MoveMonovacancy"] * 12)

# neighbors_and_mutuals order (state.py:443-459): offset of the move target for slot 7 + k
MOVE_TARGET = [(-1, 1), (0, -1), (1, 0), (0, 1), (-1, 0), (1, -1)]
# trefoil member offsets by orientation (Up, Down), roles 0, 1, 2
TREFOIL_MEMBERS = [[(0, 0), (2, 0), (0, 2)], [(0, 0), (2, 0), (2, -2)]]


def node_of(site, dim):
    return (site // dim[1], site % dim[1])


def add(node, off, dim):
    return ((node[0] + off[0]) % dim[0], (node[1] + off[1]) % dim[1])


def move_info(site, slot, dim):
    """The reference's Rule.info(move) payload (rules.py:68,85,104-106,151,186,216-218,286) for the
    move in (site, slot).  Trefoil node lists are given in role order (the reference's order is that
    of a frozenset, i.e. arbitrary)."""
    p = node_of(site, dim)
    if slot in (0, 1, 6):
        return {"node": list(p)}
    if 2 <= slot <= 5:
        return {"node": list(p), "layer": slot - 1 if slot <= 3 else slot - 3}
    if 7 <= slot <= 12:
        return {"was": list(p), "now": list(add(p, MOVE_TARGET[slot - 7], dim))}
    if slot >= MM_SLOT0:
        k, layer = divmod(slot - MM_SLOT0, 2)
        return {"was": list(p), "now": list(add(p, MOVE_TARGET[k], dim)), "layer": layer + 1}
    orient = (slot - 13) & 1
    return {"nodes": [list(add(p, off, dim)) for off in TREFOIL_MEMBERS[orient]]}


def slot_of_move(rule, move, dim):
    """(site, slot) of a move given in the reference's own key form (reference demo1/rules.py):
    a node for Create/FillDivacancy and FlipMonovacancy, (node, layer, new_layerset) for
    Create/FillMonovacancy, (old, new) for MoveDivacancy, an iterable of three nodes for the trefoil
    rules.  A pair (site, slot) of ints is passed through."""
    if (isinstance(move, tuple) and len(move) == 2 and all(isinstance(x, int) for x in move)
            and rule not in ("CreateDivacancy", "FillDivacancy", "FlipMonovacancy")):
        return move
    lin = lambda node: (node[0] % dim[0]) * dim[1] + (node[1] % dim[1])
    if rule in ("CreateDivacancy", "FillDivacancy", "FlipMonovacancy"):
        return lin(move), {"CreateDivacancy": 0, "FillDivacancy": 1, "FlipMonovacancy": 6}[rule]
    if rule == "CreateMonovacancy":
        return lin(move[0]), 1 + move[1]
    if rule == "FillMonovacancy":
        return lin(move[0]), 3 + move[1]
    if rule == "MoveDivacancy":
        old, new = move
        ks = [k for k in range(6) if add(old, MOVE_TARGET[k], dim) == add(new, (0, 0), dim)]
        if len(ks) != 1:
            raise ValueError("not a nearest-neighbour move: %r" % (move,))
        return lin(old), 7 + ks[0]
    if rule == "MoveMonovacancy":
        old, new, layer = move
        ks = [k for k in range(6) if add(old, MOVE_TARGET[k], dim) == add(new, (0, 0), dim)]
        if len(ks) != 1 or layer not in (1, 2):
            raise ValueError("not a nearest-neighbour single-layer move: %r" % (move,))
        return lin(old), MM_SLOT0 + 2 * ks[0] + (layer - 1)
    if rule in ("CreateTrefoil", "DestroyTrefoil"):
        want = set(add(tuple(n), (0, 0), dim) for n in move)
        for p in want:
            for orient in (0, 1):
                if set(add(p, off, dim) for off in TREFOIL_MEMBERS[orient]) == want:
                    return lin(p), (13 if rule == "CreateTrefoil" else 15) + orient
        raise ValueError("nodes do not form a trefoil: %r" % (move,))
    raise KeyError(rule)


def event_sites(event, dim):
    """Linear indices of the sites a logged event touches (the reference's nodes_affected_by, rules.py)."""
    mv = event["move"]
    lin = lambda node: (node[0] % dim[0]) * dim[1] + (node[1] % dim[1])
    if "nodes" in mv:
        return [lin(n) for n in mv["nodes"]]
    if "was" in mv:
        return [lin(mv["was"]), lin(mv["now"])]
    return [lin(mv["node"])]


def replay_event(status, event, dim):
    """Apply one logged event (the dict `perform_random_move` returns / the CLI writes) to a flat
    status array: an independent restatement of the transition function, in the spirit of the
    reference's animate.py replay (animate.py:82-141)."""
    rule, mv = event["rule"], event["move"]
    lin = lambda node: (node[0] % dim[0]) * dim[1] + (node[1] % dim[1])
    if rule == "CreateDivacancy":
        status[lin(mv["node"])] = 3
    elif rule == "FillDivacancy":
        status[lin(mv["node"])] = 0
    elif rule == "CreateMonovacancy":
        status[lin(mv["node"])] |= mv["layer"]
    elif rule == "FillMonovacancy":
        status[lin(mv["node"])] &= ~mv["layer"] & 0xFF
    elif rule == "FlipMonovacancy":
        status[lin(mv["node"])] ^= 3
    elif rule == "MoveDivacancy":
        status[lin(mv["was"])] = 0
        status[lin(mv["now"])] = 3
    elif rule == "MoveMonovacancy":
        status[lin(mv["was"])] &= ~mv["layer"] & 0xFF
        status[lin(mv["now"])] |= mv["layer"]
    elif rule in ("CreateTrefoil", "DestroyTrefoil"):
        site, slot = slot_of_move(rule, [tuple(n) for n in mv["nodes"]], dim)
        orient = (slot - 13) & 1
        p = node_of(site, dim)
        for role, off in enumerate(TREFOIL_MEMBERS[orient]):
            status[lin(add(p, off, dim))] = (TREFOIL_BASE + 3 * orient + role) if rule == "CreateTrefoil" else 3
    else:
        raise KeyError(rule)
