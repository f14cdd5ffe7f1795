"""Circuit latency of the five BASELINE configs on one GPU (development / reporting aid).

For every golden circuit cfg*: compile with the backend's own parameter selection, keygen, encrypt the
first fixture case, time `Server.run` (CUDA events, inputs resident, median of 3 after one warm-up),
check the decrypted output against the reference's clear evaluation, and report PBS/s.
Writes gpurun_out/config_latency.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from concrete_numpy_b200 import api
from golden_util import load_golden, same, to_value

NAMES = ["cfg1_lut4_1024", "cfg2_univariate8_64x64", "cfg3_matmul6_relu", "cfg3b_matmul6_dense_relu", "cfg4_conv2d_relu",
         "cfg5_kvdb128_query", "cfg5_kvdb128_replace", "cfg5_kvdb128_insert"]
only = sys.argv[1:]
res = {}
for name in NAMES:
    if only and not any(name.startswith(o) for o in only):
        continue
    doc = load_golden(name)
    cfg = api.Configuration(jit=True, global_p_error=doc["global_p_error"])
    t0 = time.time()
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], cfg)
    case = doc["cases"][0]
    args = [to_value(a) for a in case["args"]]
    circuit.keygen()
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    enc = circuit.encrypt(*args)
    ev = circuit.client.evaluation_keys
    out = circuit.server.run(enc, ev)
    ok = same(circuit.decrypt(out), to_value(case["expected"]))
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        circuit.server.run(enc, ev)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    fb = circuit.server._feedback
    p = circuit.server.client_specs.client_parameters.crypto
    res[name] = {"ok": bool(ok), "latency_ms": round(ms, 3), "n_pbs": int(fb.n_pbs), "pbs_per_s": round(fb.n_pbs / (ms * 1e-3), 1),
                 "params": {"p": p.p, "n": p.n, "k": p.k, "N": p.N, "pbs_level": p.pbs_level, "pbs_base_log": p.pbs_base_log,
                            "ks_level": p.ks_level, "ks_base_log": p.ks_base_log, "moduli": list(p.moduli)},
                 "setup_s": round(t_setup, 2)}
    print(name, json.dumps(res[name]), flush=True)
    circuit.cleanup()
    del circuit, enc, ev, out
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/config_latency.json", "w") as f:
    json.dump(res, f, indent=1)
