"""Development aid: bootstrap throughput of alternative GLWE shapes for the same precision (parameter-search weights)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E
dev = torch.device("cuda:0")
SETS = {
    "p2 k3 N512": E.CryptoParams(p=2, n=696, k=3, N=512, pbs_level=1, pbs_base_log=17, ks_level=2, ks_base_log=5, lwe_std=2.0 ** -15.9, glwe_std=2.0 ** -38.1),
    "p2 k1 N2048": E.CryptoParams(p=2, n=656, k=1, N=2048, pbs_level=1, pbs_base_log=18, ks_level=3, ks_base_log=4, lwe_std=2.0 ** -14.8, glwe_std=2.0 ** -51.6),
    "p4 k2 N1024": E.CryptoParams(p=4, n=812, k=2, N=1024, pbs_level=1, pbs_base_log=22, ks_level=3, ks_base_log=5, lwe_std=2.0 ** -18.9, glwe_std=2.0 ** -51.6),
    "p4 k1 N2048": E.CryptoParams(p=4, n=762, k=1, N=2048, pbs_level=1, pbs_base_log=21, ks_level=3, ks_base_log=4, lwe_std=2.0 ** -17.6, glwe_std=2.0 ** -51.6),
}
B = 4096
for name, p in SETS.items():
    ks = E.KeySet(p, dev, seed=3)
    ev = E.EvalKeys.of(ks)
    m = np.arange(B) % (1 << p.p)
    luts = E.from_np_u64(E.build_lut_accumulators(p, np.arange(1 << p.p)[None, :]), dev)
    ct = E.encrypt(ks, m.astype(np.uint64))
    out = E.table_lookup(ev, ct, luts)
    ok = bool(np.array_equal(E.decrypt(ks, out), m.astype(np.uint64)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        E.table_lookup(ev, ct, luts)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:14s} ok={ok} {3 * B / (e0.elapsed_time(e1) * 1e-3):10.0f} lookups/s (keyswitch + bootstrap)", flush=True)
    del ks, ev
    torch.cuda.empty_cache()
