"""Development aid: per-phase cycle breakdown of k_pbs (one CTA) for a parameter set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, _lib
from concrete_numpy_b200.params import REFERENCE_SETS

dev = torch.device("cuda:0")
NAMES = ["P1 decomp+stage1", "sync A", "mid fwd", "fused MAC", "mid inv", "sync B", "P5 gather+acc", "sync C", "extract+init"]
# "w" = the bootstrap of the 16-bit wide lookup (N = 2048, 3 levels, 256-thread plan), run as a 4-bit bootstrap
WIDE_PBS = E.CryptoParams(p=4, n=644, k=1, N=2048, pbs_level=3, pbs_base_log=11, ks_level=4, ks_base_log=3,
                          lwe_std=2.0 ** -14.5, glwe_std=2.0 ** -51.6)
P5 = E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                    lwe_std=2.0 ** -20, glwe_std=2.0 ** -50)
VARIANT = int(os.environ.get("PHASE_VARIANT", "0"))      # plan variant (latency plans: larger clusters)
for pbits, B in ((8, 36), (6, 148), (4, 888), ("w", 444), (5, int(os.environ.get("PHASE_B", "148")))):
    if len(sys.argv) > 1 and str(pbits) not in sys.argv[1:]:
        continue
    p = WIDE_PBS if pbits == "w" else P5 if pbits == 5 else REFERENCE_SETS[pbits]
    ks = E.KeySet(p, dev, seed=1)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(0)
    m = rng.integers(0, 1 << p.p, B).astype(np.uint64)
    table = rng.integers(0, 1 << p.p, (1, 1 << p.p))
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    E.bootstrap(ev, small, luts, None, variant=VARIANT)
    prof = torch.zeros(256, dtype=torch.int64, device=dev)
    _lib.call("fhe_pbs_set_profile", prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = E.bootstrap(ev, small, luts, None, variant=VARIANT)
    e1.record()
    torch.cuda.synchronize()
    _lib.call("fhe_pbs_set_profile", None)
    ok = bool(np.array_equal(E.decrypt(ks, out), table[0, m.astype(np.int64)].astype(np.uint64)))
    c = prof.cpu().numpy()[:9].astype(np.float64)
    ms = e0.elapsed_time(e1)
    ncmux = c.sum() / (ms * 1e-3 * 1.9e9) and None
    print(f"p={pbits} N={p.N} n={p.n} B={B} ok={ok} kernel {ms:.2f} ms; CTA0 total {c.sum()/1e6:.2f} Mcycles "
          f"=> {c.sum() / (ms * 1e-3) / 1e6:.0f} MHz")
    for nm, v in zip(NAMES, c):
        print(f"   {nm:18s} {v / max(c.sum(), 1) * 100:5.1f}%   {v / 1e3:10.0f} kcycles")
    allp = prof.cpu().numpy().reshape(16, 16).astype(np.float64)
    C = max(1, ev.variant_clusters[VARIANT])
    if C > 1:
        print("   per-rank kcycles of cluster 0 (P1 | syncA | midF | fused | midI | syncB | P5 | syncC):")
        for r in range(min(C, 16)):
            print("     rank %2d: " % r + " ".join("%7.0f" % (allp[r, i] / 1e3) for i in range(8)))
    extra = prof.cpu().numpy()[9:16].astype(np.float64)
    if extra[2:].sum() > 0:
        # pipelined kernel (pbs_pipe.cuh): slices of P1 in order; slot 0 above is the tail (stage + send of the last emit),
        # "sync C" the wait for rotated polynomial 0, "sync A" the wait for the forward data of polynomial 0
        names = ["read poly 0 + sync", "emit (0, lv1) + send", "emit (0, lv0) + send", "wait rotated poly 1",
                 "read poly 1, emit (1, lv1) + send", "wait buffer 2 landed + arrive", "emit (1, lv0) compute + cluster wait"]
        print("   P1 slices of the pipelined kernel (kcycles): " + "; ".join(f"{n} {v / 1e3:.0f}" for n, v in zip(names, extra)))
        print(f"   P1 total {(extra.sum() + c[0]) / 1e3:.0f} kcycles; CMUX total {(c.sum() + extra.sum()) / 1e3:.0f} kcycles")
    else:
        print(f"   (thread 0 inside P1: loads+diff {extra[0] / 1e3:.0f} kcycles, level loop {extra[1] / 1e3:.0f} kcycles)")
    del ks, ev
    torch.cuda.empty_cache()
