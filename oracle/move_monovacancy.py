"""MoveMonovacancy as a rule class of the REFERENCE's own framework (TEST INFRASTRUCTURE ONLY).

The reference names this rule in ``config/WSe2.yaml:8-13`` (kinds natural / merge-divacancy / split-divacancy /
swap-divacancy, each with a DFT barrier) but ships only a stub (``demo1/rules.py:295-296``:
``class RuleMoveMonovacancy(MultiKindRule): pass``), so ``config/WSe2.yaml`` cannot run.  There are NO reference
semantics to be faithful to; this file states ours, against the reference's ``MultiKindRule`` interface
(``rules.py:31-51``), so that the unmodified reference engine -- its move caches, ``weighted_choice``,
``perform_given_move`` and ``validate`` -- runs the rule and produces the golden fixtures that pin the C oracle
and the CUDA kernels.  "Parity pinned to our stated semantics" (DESIGN.md section 2).

Semantics (modelled on ``RuleMoveDivacancy``, ``rules.py:94-136``, and ``__Rule__Monovacancy``, ``:203-230``):

* a move is ``(old, new, layer)``: the chalcogen vacancy in ``layer`` (1 or 2) at node ``old`` hops to the
  neighbouring node ``new`` (one of ``grid.neighbors_and_mutuals(old)``), staying in its layer;
* it exists when ``old`` is a vacancy whose layer set contains ``layer`` and ``new`` has a defined layer set
  (not a trefoil member) that does NOT contain ``layer``;
* performing it gives ``old`` the layer set ``src & ~layer`` and ``new`` the layer set ``dst | layer``;
* kind = KINDS[2 * (src is a divacancy) + (dst is not pristine)]:
    natural          monovacancy           -> pristine neighbour
    merge-divacancy  monovacancy           -> neighbour holding the opposite-layer monovacancy (forms a divacancy)
    split-divacancy  one layer of a divacancy -> pristine neighbour (leaves the other layer behind)
    swap-divacancy   one layer of a divacancy -> opposite-layer monovacancy (the two defects trade places)
* affected nodes: ``[old, new]``; a move's existence and kind depend only on those two nodes, so
  ``moves_dependent_on`` re-derives the moves leaving every node within distance 1 of a changed node
  (as ``RuleMoveDivacancy.moves_dependent_on`` does, ``rules.py:115-117``);
* log payload: ``{'was': old, 'now': new, 'layer': layer}``.
"""

KINDS = ["natural", "merge-divacancy", "split-divacancy", "swap-divacancy"]


def make_rule_class(rules_module):
    """The class, built against the reference's ``demo1.rules`` module object."""

    class RuleMoveMonovacancy(rules_module.MultiKindRule):
        """Permit a single chalcogen vacancy to hop to a neighbouring site within its layer."""
        KINDS = list(KINDS)

        def perform(self, move, state):
            (old, new, layer) = move
            src = state.vacant_layerset_at(old)
            dst = state.vacant_layerset_at(new)
            assert (src & layer) and not (dst & layer), "no-op in move-list!"
            state.pop_vacancy(old)
            if src & ~layer:
                state.new_vacancy(old, src & ~layer)
            if dst:
                state.pop_vacancy(new)
            state.new_vacancy(new, dst | layer)

        def nodes_affected_by(self, move):
            (old, new, layer) = move
            return [old, new]

        def info(self, move):
            (old, new, layer) = move
            return {"was": old, "now": new, "layer": layer}

        def kinds(self):
            return list(self.KINDS)

        def all_moves(self, state):
            for x in state.vacancies():
                yield from self.moves_from_node(state, x.node)

        def moves_dependent_on(self, state, nodes):
            for node in state.grid.nodes_in_distance_range(nodes, 0, 1):
                yield from self.moves_from_node(state, node)

        def moves_from_node(self, state, node):
            if state.is_vacancy(node):
                src = state.vacant_layerset_at(node)
                for (nbr, near, far) in state.grid.neighbors_and_mutuals(node):
                    if state.has_defined_layerset(nbr):
                        dst = state.vacant_layerset_at(nbr)
                        for layer in [1, 2]:
                            if (src & layer) and not (dst & layer):
                                yield ((node, nbr, layer), self.move_kind(src, dst))

        def move_kind(self, src, dst):
            return self.KINDS[2 * int(src == 3) + int(dst != 0)]

    return RuleMoveMonovacancy


def install(mods):
    """Replace the reference's stub with the class above (on the module object only; nothing is written)."""
    rules = mods["rules"]
    stub = getattr(rules, "RuleMoveMonovacancy", None)
    if stub is None or not getattr(stub, "_kmc_b200_extension", False):
        cls = make_rule_class(rules)
        cls._kmc_b200_extension = True
        cls.__module__ = rules.__name__
        cls.__qualname__ = "RuleMoveMonovacancy"
        rules.RuleMoveMonovacancy = cls
    return rules.RuleMoveMonovacancy
