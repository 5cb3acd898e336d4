"""YAML configuration -> rule specs + state generator.

Keeps the reference's schema and error behaviour (reference demo1/config.py:10-190): top-level keys
`rules` (required) and `state` (optional), anything else is an error; per rule an optional `enabled`
flag and exactly one of `barrier` / `rate`, given as a number (shorthand for {natural: x}) or a
{kind: number} mapping with non-negative values; `state.random` with `mode` in {exact, approx} and
`divacancy` / `monovacancy` fractions (float or "N%").  Several files are deep-merged, later wins.

Two deliberate differences so that the reference's own shipped files run as they are:
  * a leading "Rule" in a rule name is accepted (test-cfg.yaml still uses the class names);
  * MoveMonovacancy, which config/WSe2.yaml:8-13 names with four kinds but the reference only stubs
    (rules.py:295-296 -> NotImplementedError), is a real rule here (semantics: tables.py / DESIGN.md section 2).
"""
import logging
import warnings
from functools import partial

import yaml

from . import state as state_mod
from . import tables as T
from .sim import RuleSpec


def error(msg="", where=""):
    raise RuntimeError("config: " + where + (": " if where else "") + msg)


def from_dict(d):
    """Consumes ``d`` (every recognised key is popped, as the reference does)."""
    if not isinstance(d, dict):
        error("document is not a YAML mapping")
    out = {"rule_specs": consume_rule_specs(d), "state_gen_func": consume_state(d)}
    if d:
        error("unrecognized section: %r" % d.popitem()[0])
    return out


def consume_rule_specs(config):
    if "rules" not in config:
        error("missing required key: 'rules'")
    rules_map = config.pop("rules")
    specs = []
    for key in list(rules_map):
        name = key[4:] if key.startswith("Rule") else key
        rule_map = rules_map.pop(key)
        if name not in T.RULE_KINDS:
            error("unknown rule: %s" % name)
        enabled = rule_map.pop("enabled", True)
        if not isinstance(enabled, bool):
            error("enabled must be a boolean", where="rules:%s" % name)
        if enabled:
            specs.append(build_rule_spec(name, rule_map))
    return specs


def build_rule_spec(name, rule_map):
    where = "rules:" + name
    found = [k for k in ("barrier", "rate") if k in rule_map]
    if len(found) != 1:
        error("need exactly one of: 'barrier', 'rate'", name)
    key = found[0]
    the_map = rule_map.pop(key)
    if isinstance(the_map, (int, float)) and not isinstance(the_map, bool):
        the_map = {T.DEFAULT_KIND: the_map}
    if not isinstance(the_map, dict):
        error("barrier/rate must be a real number or a mapping", where)
    if not all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in the_map.values()):
        error("non-numeric energy barrier or rate", where)
    if any(x < 0 for x in the_map.values()):
        error("negative energy barrier or rate", where)
    if rule_map:
        error("invalid attribute: %r" % next(iter(rule_map)), where)
    return RuleSpec(name, {k: float(v) for k, v in the_map.items()}, rate_is_barrier=(key == "barrier"))


def consume_state(config):
    d = config.pop("state", {})
    found = [k for k in ("random",) if k in d]
    sub = d.pop("random") if found else {"mode": "exact"}
    if d:
        error("invalid attribute: {!r}".format(next(iter(d))), where="state")
    return _consume_state_random(sub)


def _consume_state_random(d):
    d = dict(d)
    if "mode" not in d:
        error("missing required key: 'mode'")
    mode = d.pop("mode")
    if mode not in state_mod.RANDOM_MODES:
        error("invalid choice for 'mode'. Choices: %s" % state_mod.RANDOM_MODES)

    def read_rate(val):
        if isinstance(val, str) and val.endswith("%"):
            return float(val[:-1]) / 100
        return float(val)

    params = {key: read_rate(d.pop(key, 0)) for key in state_mod.RANDOM_PARAMS}
    if any(x < 0.0 for x in params.values()):
        error("negative rate", where="state.random")
    if sum(params.values()) > 1.0:
        error("sum of rates > 100%", where="state.random")
    if d:
        error("unrecognized parameter: %r\nvalid parameters are: %r" % (d.popitem()[0], state_mod.RANDOM_PARAMS),
              where="state:random")
    return partial(state_mod.gen_random_state, mode=mode, params=params)


def merge(older, newer, where=()):
    """Deep merge of two config trees: mappings are merged key by key, anything else is replaced by the newer
    value with a logged warning.  Keys keep their order of first appearance (the reference walks a set here,
    which makes the rule order -- and with it tie-breaks -- hash-dependent: SURVEY Q3)."""
    both_maps = type(older) is dict and type(newer) is dict
    if not both_maps:
        if older != newer:
            logging.warning("config key overriden at %s\n  old: %r\n  new: %r",
                            ":".join(repr(k) for k in where) or "root", older, newer)
        return newer
    merged = dict(older)
    for key, value in newer.items():
        merged[key] = merge(older[key], value, where + (key,)) if key in older else value
    return merged


def load_all(files):
    """Parse every config file and fold them left to right with merge(); empty files are skipped."""
    result = {}
    for handle in files:
        tree = yaml.safe_load(handle)
        if tree is None:
            logging.debug("empty config file")
            continue
        if not isinstance(tree, dict):
            error("config must be a YAML mapping")
        result = merge(result, tree)
    return result
