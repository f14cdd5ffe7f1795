"""Helpers shared by the CPU and GPU golden tests."""
import glob
import gzip
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    names = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.json*"))):
        names.append(os.path.basename(path).split(".json")[0])
    return names


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".json")
    if os.path.exists(path):
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)
    with gzip.open(path + ".gz", "rt", encoding="utf-8") as f:
        return json.load(f)


def to_value(doc):
    if "tuple" in doc:
        return tuple(to_value(t) for t in doc["tuple"])
    a = np.array(doc["data"], dtype=np.int64).reshape(doc["shape"])
    return int(a) if doc["shape"] == [] else a


def same(a, b):
    if isinstance(a, tuple) or isinstance(b, tuple):
        return isinstance(a, tuple) and isinstance(b, tuple) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    return np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))
