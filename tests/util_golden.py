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


N_CLASSES, N_SLOTS = 17, 29


def load(name):
    """Fixtures made before MoveMonovacancy existed carry 13-class / 17-slot tables: pad them to the current
    widths (absent classes have rate 0 / count 0, absent slots are empty)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(str(z["meta"]))
    for k in ("rates", "counts0", "counts1"):
        if k in d and d[k].shape[-1] < N_CLASSES:
            d[k] = np.concatenate([d[k], np.zeros(N_CLASSES - d[k].shape[-1], dtype=d[k].dtype)])
    if "slots1" in d and d["slots1"].shape[-1] < N_SLOTS:
        pad = np.full((d["slots1"].shape[0], N_SLOTS - d["slots1"].shape[1]), 0xFF, dtype=np.uint8)
        d["slots1"] = np.concatenate([d["slots1"], pad], axis=1)
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
