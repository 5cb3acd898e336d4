"""Helpers to load the committed golden fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def trajectory_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("state_gen", "native_stats", "zobrist_ref", "animate_replay",
                                                         "boundary_ties")))


def boundary_tie_names():
    """Fixtures whose class draws were placed exactly on cumulative-weight boundaries (injected draws)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "boundary_ties*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(str(z["meta"]))
    return d


def enabled_classes(meta):
    return sorted(meta["class_order"])


def chi2_two_sample(a, b):
    """Two-sample chi-square statistic and degrees of freedom for two count vectors over the same
    categories (categories empty in both samples are dropped)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    keep = (a + b) > 0
    a, b = a[keep], b[keep]
    ka, kb = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    stat = float((((ka * a - kb * b) ** 2) / (a + b)).sum())
    return stat, int(keep.sum()) - 1
