"""Development aid: N = 2048 bootstrap throughput under the three kernels of its plan -- k_pbs (two CTAs per SM), k_pbs_group
without staging (one CTA per SM, two slots, tensor-memory twiddles) and k_pbs_group with the staged GGSW ring -- at 1..3
decomposition levels (level >= 2 does not fit the ring)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, _lib

dev = torch.device("cuda:0")
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
for level, base_log, n in ((1, 23, 742), (2, 15, 668), (3, 11, 644)):
    p = E.CryptoParams(p=4, n=n, k=1, N=2048, pbs_level=level, pbs_base_log=base_log, ks_level=5, ks_base_log=3,
                       lwe_std=2.0 ** -17, glwe_std=2.0 ** -51.6)
    ks = E.KeySet(p, dev, seed=3)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(0)
    m = rng.integers(0, 16, B).astype(np.uint64)
    table = rng.integers(0, 16, (1, 16))
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    row = []
    for mode in (0, 2, 1):
        prev = lib.fhe_pbs_set_group(mode)
        out = E.bootstrap(ev, small, luts, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = E.bootstrap(ev, small, luts, None)
        e1.record()
        torch.cuda.synchronize()
        lib.fhe_pbs_set_group(prev)
        ok = bool(np.array_equal(E.decrypt(ks, out), table[0, m.astype(np.int64)].astype(np.uint64)))
        row.append(f"mode {mode}: {3 * B / (e0.elapsed_time(e1) * 1e-3):9.0f} PBS/s ok={ok}")
    print(f"N=2048 level={level} n={n} B={B}  " + " | ".join(row), flush=True)
    del ks, ev
    torch.cuda.empty_cache()
