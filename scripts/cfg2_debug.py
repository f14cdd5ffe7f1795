"""Development probe: run the golden cfg2 circuit and list which outputs differ from the reference evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CNP_B200_KEY_SEED", "20261019")
import numpy as np, torch
from concrete_numpy_b200 import api
from golden_util import load_golden, to_value

doc = load_golden("cfg2_univariate8_64x64")
cfg = api.Configuration(jit=True, global_p_error=doc["global_p_error"])
circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], cfg)
for case in doc["cases"]:
    args = [to_value(a) for a in case["args"]]
    got = np.asarray(circuit.encrypt_run_decrypt(*args))
    exp = np.asarray(to_value(case["expected"]))
    bad = np.argwhere(got != exp)
    print("PIPE", os.environ.get("FHE_PBS_PIPE", "1"), "mismatches", len(bad), "of", exp.size, "first", bad[:12].tolist(),
          "got", got[tuple(bad[:6].T)].tolist() if len(bad) else [], "exp", exp[tuple(bad[:6].T)].tolist() if len(bad) else [], flush=True)
