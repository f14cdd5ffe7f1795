"""Development aid: bootstrap latency per plan variant and batch size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E
from concrete_numpy_b200.params import REFERENCE_SETS
dev = torch.device("cuda:0")
sets = {"p5": E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                             lwe_std=2.0 ** -20, glwe_std=2.0 ** -50),
        "p4": REFERENCE_SETS[4], "p6": REFERENCE_SETS[6],
        "p7": E.CryptoParams(p=7, n=920, k=1, N=16384, pbs_level=2, pbs_base_log=15, ks_level=6, ks_base_log=3,
                             lwe_std=2.0 ** -22, glwe_std=2.0 ** -62)}
only = sys.argv[1:]
for name, p in sets.items():
    if only and name not in only:
        continue
    ks = E.KeySet(p, dev, seed=1)
    ev = E.EvalKeys.of(ks)
    table = np.arange(1 << p.p)[None, :]
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    for B in (1, 8, 17, 30, 60):
        m = np.arange(B) % (1 << p.p)
        small = E.keyswitch(ev, E.encrypt(ks, m.astype(np.uint64)))
        row = []
        for v, c in enumerate(ev.variant_clusters):
            E.bootstrap(ev, small, luts, None, variant=v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                out = E.bootstrap(ev, small, luts, None, variant=v)
            e1.record(); torch.cuda.synchronize()
            row.append(f"C={c}: {e0.elapsed_time(e1) / 3:.2f} ms")
        print(name, f"N={p.N} n={p.n} B={B:3d}", " | ".join(row), flush=True)
    del ks, ev
    torch.cuda.empty_cache()
