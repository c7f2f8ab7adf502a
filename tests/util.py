"""Shared helpers of the test-suite: dataset fixtures (tests/golden/data/*.gz are the SBA-format
problems shipped with the reference, gzip-compressed) and comparison utilities."""
import gzip
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "data")
_TMP = None

DATASETS = {
    "7": ("7camsvarK.txt", "7pts.txt", 11),
    "9": ("9camsvarK.txt", "9pts.txt", 11),
    "54": ("54camsvarK.txt", "54pts.txt", 11),
    "54KD": ("54camsvarKD.txt", "54pts.txt", 16),
    "T21": ("Trafalgar-21-11315-cams.txt", "Trafalgar-21-11315-pts.txt", 11),
}


def data_file(name):
    """Path of the decompressed copy of tests/golden/data/<name>.gz."""
    global _TMP
    if _TMP is None:
        _TMP = tempfile.mkdtemp(prefix="psba_data_")
    dst = os.path.join(_TMP, name)
    if not os.path.exists(dst):
        with gzip.open(os.path.join(DATA, name + ".gz"), "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
    return dst


def dataset_paths(key):
    c, p, cnp = DATASETS[key]
    return data_file(c), data_file(p), cnp


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def pattern(trace):
    """accept / reject string of a run log: A accepted try, x rejected, C modified-Cholesky event"""
    return "".join("C" if r["phase"] == 2 else ("A" if r["accepted"] else "x") for r in trace)
