"""The reference's second statement of the transition function -- animate.py `do_basic` (animate.py:82-141),
the replay a consumer of the JSON log runs -- executed on logs in the PRODUCT's format (TEST INFRASTRUCTURE
ONLY; run in the build container by oracle/make_golden.py, needs /root/reference).

animate.py imports pylab (absent here) at module level, so the functions it needs are taken from its source
text by slicing ("secondary oracle") and executed unmodified.  One compatibility shim, in the
spirit of oracle/ref_shim: `do_basic` indexes with a list of two lists (`state[[xs, ys]] = id`,
animate.py:126-131), which NumPy < 1.23 read as a tuple of index arrays and NumPy 2 reads as one 2x3 index
array; the namespace's `np` hands out an ndarray subclass that restores the old meaning.
"""
import os
import re

import numpy as np

REFERENCE = "/root/reference"


class _OldIndexArray(np.ndarray):
    @staticmethod
    def _fix(k):
        return tuple(k) if isinstance(k, list) and k and isinstance(k[0], list) else k

    def __getitem__(self, k):
        return super().__getitem__(self._fix(k))

    def __setitem__(self, k, v):
        super().__setitem__(self._fix(k), v)


class _Np:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def full(*a, **k):
        return np.full(*a, **k).view(_OldIndexArray)

    @staticmethod
    def array(x, copy=True, **k):
        return np.array(x, copy=copy, **k).view(_OldIndexArray)


def _load():
    src = open(os.path.join(REFERENCE, "animate.py")).read()
    start = src.index("FLAG_TREFOIL = 2**16")
    end = src.index("def do_animation(")
    helpers = src[src.index("def pop_entire("):src.index("def say(")]
    ns = {"np": _Np(), "warnings": __import__("warnings"), "die": lambda *a, **k: (_ for _ in ()).throw(RuntimeError(a))}
    exec(compile(src[start:end] + "\n" + helpers, "animate.py[sliced]", "exec"), ns)
    assert re.search(r"def do_basic\(fullinfo\)", src[start:end])
    return ns


def replay(fullinfo):
    """Final int32 site array of the reference's replay: 0 pristine, 1/2 monovacancy layer, 3 divacancy,
    >= 65536 trefoil id (nodes of one trefoil share an id)."""
    ns = _load()
    state = None
    for state in ns["do_basic"](fullinfo):
        pass
    return np.asarray(state)
