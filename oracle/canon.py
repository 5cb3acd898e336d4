"""Canonical tables shared by the oracle tools (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it, and only as the checker.

The tables restate, as plain integers, what the reference derives at run time:

* stencils          <- /root/reference/demo1/state.py:392-401,443-459 and hexagonal.py:56-62
* class ids         <- the (rule, kind) pairs of /root/reference/demo1/rules.py:60-293
* per-site slots    <- the move keys those rules produce (slot table)
* status codes      <- /root/reference/demo1/state.py:22-23,171-177 (0..3 = vacant layer set);
                       4.. = trefoil membership (orientation, role), our own encoding.
"""

# --- status codes -----------------------------------------------------------------------
PRISTINE, MONO1, MONO2, DIVAC = 0, 1, 2, 3
TREFOIL_BASE = 4          # code = 4 + 3*orientation + role ; orientation 0 = Up, 1 = Down
N_STATUS = 10

# --- stencils (axial offsets) -----------------------------------------------------------
# Grid.neighbors: rotations of cubic (-1, 0, 1)  (state.py:392-394)
NEIGHBORS = [(-1, 0), (0, -1), (1, -1), (1, 0), (0, 1), (-1, 1)]
# Grid.trefoil_neighbors: rotations of cubic (2, -2, 0)  (state.py:396-401)
TREFOIL_NEIGHBORS = [(2, -2), (2, 0), (0, 2), (-2, 2), (-2, 0), (0, -2)]
# Grid.neighbors_and_mutuals: (nbr, near, far), in the order the reference yields them
# (state.py:443-459)
NBR_NEAR_FAR = [
    ((-1, 1), (0, 1), (-1, 0)),
    ((0, -1), (-1, 0), (1, -1)),
    ((1, 0), (1, -1), (0, 1)),
    ((0, 1), (-1, 1), (1, 0)),
    ((-1, 0), (0, -1), (-1, 1)),
    ((1, -1), (1, 0), (0, -1)),
]
# A trefoil is an equilateral triangle of side 2.  Anchored forms:
#   Up(p)   = {p, p+(2,0), p+(0,2)}      Down(p) = {p, p+(2,0), p+(2,-2)}
TREFOIL_MEMBERS = [
    [(0, 0), (2, 0), (0, 2)],    # orientation 0 (Up),   roles 0,1,2
    [(0, 0), (2, 0), (2, -2)],   # orientation 1 (Down), roles 0,1,2
]

# --- event classes ------------------------------------------------------------------------
CLASSES = [
    ("CreateDivacancy", "natural"),          # 0
    ("FillDivacancy", "natural"),            # 1
    ("MoveDivacancy", "natural"),            # 2
    ("MoveDivacancy", "missing-metal"),      # 3
    ("MoveDivacancy", "missing-comb"),       # 4
    ("MoveDivacancy", "missing-both"),       # 5
    ("CreateTrefoil", "natural"),            # 6
    ("DestroyTrefoil", "natural"),           # 7
    ("CreateMonovacancy", "from-double"),    # 8
    ("CreateMonovacancy", "make-empty"),     # 9
    ("FillMonovacancy", "make-double"),      # 10
    ("FillMonovacancy", "from-empty"),       # 11
    ("FlipMonovacancy", "natural"),          # 12
    # MoveMonovacancy is a stub in the reference (rules.py:295-296); its kinds are named in config/WSe2.yaml:8-13.
    # Semantics: oracle/move_monovacancy.py (extension, stated in DESIGN.md).
    ("MoveMonovacancy", "natural"),          # 13  monovacancy -> pristine neighbour
    ("MoveMonovacancy", "merge-divacancy"),  # 14  monovacancy -> neighbour with the opposite-layer monovacancy
    ("MoveMonovacancy", "split-divacancy"),  # 15  one layer of a divacancy -> pristine neighbour
    ("MoveMonovacancy", "swap-divacancy"),   # 16  one layer of a divacancy -> opposite-layer monovacancy
]
N_CLASSES = len(CLASSES)
CLASS_ID = {rk: i for i, rk in enumerate(CLASSES)}
RULES = ["CreateDivacancy", "FillDivacancy", "MoveDivacancy", "CreateTrefoil",
         "DestroyTrefoil", "CreateMonovacancy", "FillMonovacancy", "FlipMonovacancy", "MoveMonovacancy"]
RULE_KINDS = {r: [k for (rr, k) in CLASSES if rr == r] for r in RULES}

# --- per-site move slots ------------------------------------------------------------------
N_SLOTS = 29
SLOT_NONE = 0xFF
# slots 17..28: MoveMonovacancy (node, nbr_k, layer) at 17 + 2*k + (layer - 1), k in neighbors_and_mutuals order
MM_SLOT0 = 17
SLOT_RULE = (["CreateDivacancy", "FillDivacancy"] + ["CreateMonovacancy"] * 2 +
             ["FillMonovacancy"] * 2 + ["FlipMonovacancy"] + ["MoveDivacancy"] * 6 +
             ["CreateTrefoil"] * 2 + ["DestroyTrefoil"] * 2 + ["MoveMonovacancy"] * 12)
assert len(SLOT_RULE) == N_SLOTS


def wrap(node, dim):
    return (node[0] % dim[0], node[1] % dim[1])


def add(node, off, dim):
    return ((node[0] + off[0]) % dim[0], (node[1] + off[1]) % dim[1])


def lin(node, dim):
    return node[0] * dim[1] + node[1]


def unlin(i, dim):
    return (i // dim[1], i % dim[1])


def trefoil_nodes(anchor, orient, dim):
    return [add(anchor, off, dim) for off in TREFOIL_MEMBERS[orient]]


def trefoil_anchor(nodes, dim):
    """(anchor, orientation) of a 3-node trefoil; raises if it is not one."""
    nodes = [tuple(n) for n in nodes]
    found = []
    for p in nodes:
        for orient in (0, 1):
            if set(trefoil_nodes(p, orient, dim)) == set(nodes):
                found.append((p, orient))
    if len(found) != 1:
        raise ValueError("not a uniquely anchored trefoil: %r -> %r" % (nodes, found))
    return found[0]


def move_from_slot(site, slot, status, dim):
    """Reference move key for (site, slot) on a lattice whose status at ``site`` is ``status``.

    Returns (rule_name, move) with ``move`` in the exact form the reference's rules use as
    dictionary keys (rules.py: node tuple / (node, layer, new_layerset) / (old, new) /
    frozenset of 3 nodes).
    """
    p = unlin(site, dim)
    rule = SLOT_RULE[slot]
    if slot in (0, 1, 6):
        return rule, p
    if slot in (2, 3):
        layer = slot - 1
        return rule, (p, layer, status | layer)
    if slot in (4, 5):
        layer = slot - 3
        return rule, (p, layer, status & ~layer)
    if 7 <= slot <= 12:
        return rule, (p, add(p, NBR_NEAR_FAR[slot - 7][0], dim))
    if slot >= MM_SLOT0:
        k, layer = divmod(slot - MM_SLOT0, 2)
        return rule, (p, add(p, NBR_NEAR_FAR[k][0], dim), layer + 1)
    orient = (slot - 13) & 1
    return rule, frozenset(trefoil_nodes(p, orient, dim))


def slot_from_move(rule, move, dim):
    """Inverse of :func:`move_from_slot`: canonical (site, slot) of a reference move key."""
    if rule in ("CreateDivacancy", "FillDivacancy", "FlipMonovacancy"):
        return lin(move, dim), {"CreateDivacancy": 0, "FillDivacancy": 1, "FlipMonovacancy": 6}[rule]
    if rule == "CreateMonovacancy":
        node, layer, _new = move
        return lin(node, dim), 1 + layer
    if rule == "FillMonovacancy":
        node, layer, _new = move
        return lin(node, dim), 3 + layer
    if rule == "MoveDivacancy":
        old, new = move
        ks = [k for k in range(6) if add(old, NBR_NEAR_FAR[k][0], dim) == tuple(new)]
        if len(ks) != 1:
            raise ValueError("aliased MoveDivacancy geometry (lattice too small): %r" % (move,))
        return lin(old, dim), 7 + ks[0]
    if rule == "MoveMonovacancy":
        old, new, layer = move
        ks = [k for k in range(6) if add(old, NBR_NEAR_FAR[k][0], dim) == tuple(new)]
        if len(ks) != 1:
            raise ValueError("aliased MoveMonovacancy geometry (lattice too small): %r" % (move,))
        return lin(old, dim), MM_SLOT0 + 2 * ks[0] + (layer - 1)
    if rule in ("CreateTrefoil", "DestroyTrefoil"):
        anchor, orient = trefoil_anchor(move, dim)
        return lin(anchor, dim), (13 if rule == "CreateTrefoil" else 15) + orient
    raise KeyError(rule)
