"""Development aid: tensor-core keyswitch vs the CUDA-core kernel (bit-exact) + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream
cases = [(128, 256, 31, 1, 3), (128, 256, 31, 2, 3), (200, 1024, 97, 5, 3), (1000, 2048, 742, 5, 3), (4096, 32768, 996, 7, 3)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (B, kN, n, level, base_log) in cases:
    assert lib.fhe_ks_tc_supported(kN, n, level, base_log)
    g = torch.Generator(device=dev); g.manual_seed(B + kN)
    ksk = torch.randint(-2**63, 2**63 - 1, (kN, level, n + 1), dtype=torch.int64, device=dev, generator=g)
    ct = torch.randint(-2**63, 2**63 - 1, (B, kN + 1), dtype=torch.int64, device=dev, generator=g)
    ref = torch.empty((B, n + 1), dtype=torch.int64, device=dev)
    _lib.call("fhe_keyswitch_batch", ref.data_ptr(), ct.data_ptr(), B, ksk.data_ptr(), kN, n, level, base_log, st())
    ksk_tc = torch.empty(lib.fhe_ksk_tc_bytes(kN, n, level), dtype=torch.uint8, device=dev)
    _lib.call("fhe_ksk_to_tc", ksk_tc.data_ptr(), ksk.data_ptr(), kN, n, level, st())
    scratch = torch.empty(lib.fhe_keyswitch_tc_scratch_bytes(B, kN, level), dtype=torch.uint8, device=dev)
    out = torch.full((B, n + 1), 7, dtype=torch.int64, device=dev)
    _lib.call("fhe_keyswitch_batch_tc", out.data_ptr(), ct.data_ptr(), B, ksk_tc.data_ptr(), scratch.data_ptr(), kN, n, level, base_log, st())
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    nbad = int((out != ref).sum().item())
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    _lib.call("fhe_keyswitch_batch", ref.data_ptr(), ct.data_ptr(), B, ksk.data_ptr(), kN, n, level, base_log, st())
    e[1].record()
    _lib.call("fhe_keyswitch_batch_tc", out.data_ptr(), ct.data_ptr(), B, ksk_tc.data_ptr(), scratch.data_ptr(), kN, n, level, base_log, st())
    e[2].record(); torch.cuda.synchronize()
    print(f"B={B} kN={kN} n={n} level={level}: bit-exact {same} (mismatches {nbad}/{out.numel()}); cuda-core {e[0].elapsed_time(e[1]):.2f} ms, tensor-core {e[1].elapsed_time(e[2]):.2f} ms", flush=True)
    if not same:
        d = (out != ref).nonzero()[:5].tolist()
        print("   first mismatches", d, [(hex(out[i, j].item() & (2**64 - 1)), hex(ref[i, j].item() & (2**64 - 1))) for i, j in d])
    del ksk, ct, ksk_tc, scratch
    torch.cuda.empty_cache()
