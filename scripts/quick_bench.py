"""Quick device-timed throughput probe (development aid, not the contract bench)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from concrete_numpy_b200 import engine as E

dev = torch.device("cuda:0")
res = {"fp64_peak_tflops": E.measure_fp64_peak(dev)}
print(res, flush=True)

SETS = {
    "p4": E.CryptoParams(p=4, n=742, k=1, N=2048, pbs_level=1, pbs_base_log=23, ks_level=5, ks_base_log=3,
                         lwe_std=7.0698e-6, glwe_std=2.9404e-16),
    "p5": E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                         lwe_std=3.2e-6, glwe_std=2.168e-19),
    "p6": E.CryptoParams(p=6, n=864, k=1, N=8192, pbs_level=2, pbs_base_log=15, ks_level=6, ks_base_log=3,
                         lwe_std=7.58e-7, glwe_std=2.168e-19),
    "p8": E.CryptoParams(p=8, n=996, k=1, N=32768, pbs_level=2, pbs_base_log=15, ks_level=7, ks_base_log=3,
                         lwe_std=6.7677e-8, glwe_std=2.168e-19),
}
BATCH = {"p4": 4096, "p5": 2048, "p6": 1024, "p8": int(os.environ.get("QB_P8_BATCH", "240"))}   # p8: 15 clusters x 16 waves


def flops_pbs(p):
    M = p.N // 2
    return p.n * ((p.k + 1) * (p.pbs_level + 1) * 5 * M * np.log2(M) + 8 * (p.k + 1) ** 2 * p.pbs_level * M)


which = sys.argv[1:] or list(SETS)
for name in which:
    p = SETS[name]
    t0 = time.time()
    ks = E.KeySet(p, dev, seed=42)
    torch.cuda.synchronize()
    t_keygen = time.time() - t0
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(0)
    B = BATCH[name]
    m = rng.integers(0, 1 << p.p, B).astype(np.uint64)
    table = rng.integers(0, 1 << p.p, (1, 1 << p.p))
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    ct = E.encrypt(ks, m)
    for _ in range(2):
        small = E.keyswitch(ev, ct)
        out = E.bootstrap(ev, small, luts, None)
    torch.cuda.synchronize()
    ok = bool(np.array_equal(E.decrypt(ks, out), table[0, m.astype(np.int64)].astype(np.uint64)))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 3
    tk = tp = 0.0
    for _ in range(reps):
        e[0].record()
        small = E.keyswitch(ev, ct)
        e[1].record()
        out = E.bootstrap(ev, small, luts, None)
        e[2].record()
        torch.cuda.synchronize()
        tk += e[0].elapsed_time(e[1]) / reps
        tp += e[1].elapsed_time(e[2]) / reps
    fl = flops_pbs(p)
    r = dict(set=name, B=B, correct=ok, keygen_s=round(t_keygen, 2), ks_ms=round(tk, 3), pbs_ms=round(tp, 3),
             pbs_per_s=round(B / (tp * 1e-3), 1), ks_pbs_per_s=round(B / ((tp + tk) * 1e-3), 1),
             pbs_tflops=round(B * fl / (tp * 1e-3) / 1e12, 3),
             frac_fp64=round(B * fl / (tp * 1e-3) / 1e12 / res["fp64_peak_tflops"], 4))
    print(json.dumps(r), flush=True)
    res[name] = r
    del ks, ev
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/quick_bench.json", "w") as f:
    json.dump(res, f, indent=1)
