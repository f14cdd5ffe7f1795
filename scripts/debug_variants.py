"""Development aid: per plan variant, compare (a) native BSK conversion vs relayout, (b) PBS correctness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, _lib
dev = torch.device("cuda:0")
for (p_bits, n, k, N, level, base_log) in [(3, 16, 1, 1024, 2, 10), (5, 16, 1, 4096, 2, 13), (4, 20, 1, 2048, 1, 23)]:
    p = E.CryptoParams(p=p_bits, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=8, ks_base_log=3,
                       lwe_std=2.0 ** -40, glwe_std=2.0 ** -55)
    ks = E.KeySet(p, dev, seed=9, keep_std_bsk=True)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(N + 1)
    B = 9
    m = rng.integers(0, 1 << p_bits, B).astype(np.uint64)
    table = rng.integers(0, 1 << p_bits, (1, 1 << p_bits))
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    exp = table[0, m.astype(np.int64)].astype(np.uint64)
    for v, c in enumerate(ev.variant_clusters):
        native = torch.empty_like(ev.bsk_f)
        _lib.call("fhe_bsk_to_fourier_variant", native.data_ptr(), ks.bsk_std.data_ptr(), p.n, p.k, p.N, p.pbs_level,
                  p.pbs_base_log, v, torch.cuda.current_stream().cuda_stream)
        rel = ev.bsk_for_variant(v)
        torch.cuda.synchronize()
        d = (native - rel).abs().max().item() / native.abs().max().item()
        out = torch.empty((B, p.ct_words), dtype=torch.int64, device=dev)
        _lib.call("fhe_pbs_batch_variant", out.data_ptr(), small.data_ptr(), B, native.data_ptr(), luts.data_ptr(), None,
                  p.n, p.k, p.N, p.pbs_level, p.pbs_base_log, v, torch.cuda.current_stream().cuda_stream)
        got_native = E.decrypt(ks, out)
        out2 = E.bootstrap(ev, small, luts, None, variant=v)
        got_rel = E.decrypt(ks, out2)
        print(f"N={N} variant {v} C={c}: relayout-vs-native rel diff {d:.2e}; native ok {np.array_equal(got_native, exp)}; "
              f"relayout ok {np.array_equal(got_rel, exp)}", flush=True)
