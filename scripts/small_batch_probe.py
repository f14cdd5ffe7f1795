"""Development probe: keyswitch / bootstrap / element-wise latency at the small batches of the key-value insert chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E
dev = torch.device("cuda:0")
p = E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                   lwe_std=2.0 ** -20, glwe_std=2.0 ** -50)
ks = E.KeySet(p, dev, seed=1)
ev = E.EvalKeys.of(ks)
luts = E.from_np_u64(E.build_lut_accumulators(p, np.arange(32)[None, :]), dev)


def timed(fn, reps=50):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in (1, 17, 64, 148):
    ct = E.encrypt(ks, (np.arange(B) % 32).astype(np.uint64))
    small = E.keyswitch(ev, ct)
    v = ev.variant_for_batch(B)
    print(f"B={B:4d}: keyswitch {timed(lambda: E.keyswitch(ev, ct)):.3f} ms | bootstrap (variant {v}, C={ev.variant_clusters[v]}) "
          f"{timed(lambda: E.bootstrap(ev, small, luts, None, variant=v), 10):.3f} ms | table_lookup {timed(lambda: E.table_lookup(ev, ct, luts), 10):.3f} ms | "
          f"add {timed(lambda: E.lin_add(ct, ct, p.ct_words)):.4f} ms", flush=True)
