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
]
CLASS_ID = {rk: i for i, rk in enumerate(CLASSES)}
N_CLASSES = len(CLASSES)

# rule name -> kinds in Rule.kinds() order (rules.py:108-109,241,260; default sim.py:262-265)
RULE_KINDS = {}
for _r, _k in CLASSES:
    RULE_KINDS.setdefault(_r, []).append(_k)
IMPLEMENTED_RULES = list(RULE_KINDS)
# named in config/WSe2.yaml:8-13 but a stub in the reference (rules.py:295-296)
STUB_RULES = {"MoveMonovacancy": ["natural", "merge-divacancy", "split-divacancy", "swap-divacancy"]}

N_SLOTS = 17
SLOT_RULE = (["CreateDivacancy", "FillDivacancy"] + ["CreateMonovacancy"] * 2 + ["FillMonovacancy"] * 2 +
             ["FlipMonovacancy"] + ["MoveDivacancy"] * 6 + ["CreateTrefoil"] * 2 + ["DestroyTrefoil"] * 2)

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
    orient = (slot - 13) & 1
    return {"nodes": [list(add(p, off, dim)) for off in TREFOIL_MEMBERS[orient]]}
