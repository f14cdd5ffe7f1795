"""HBM roofline of the torus linear kernels at the BASELINE shapes (SURVEY 8d algorithmic bytes / CUDA-event time /
measured HBM copy bandwidth of MEASURED_PEAKS.json).  Writes gpurun_out/op_profile.json.
usage: op_profile.py [reps]            (ncu target: op_profile.py 1)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from concrete_numpy_b200 import engine as E

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:                                                       # noqa: BLE001
    HBM = 6650.0
rng = np.random.default_rng(0)
res = {"hbm_gbs_peak": HBM}


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()      # > L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    res[name] = {"ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "gbs": round(gbs, 1), "frac_of_hbm": round(gbs / HBM, 3), "note": note}
    print(name, json.dumps(res[name]), flush=True)


def rand_ct(rows, L):
    return torch.randint(-(1 << 62), 1 << 62, (rows, L), dtype=torch.int64, device=dev)


# cfg3: x [64,128] @ W [128,64] ternary (10 non-zeros per column), N = 8192
L = 8193
x = rand_ct(64 * 128, L)
W = np.zeros((128, 64), dtype=np.int64)
for n in range(64):
    W[rng.choice(128, 10, replace=False), n] = rng.choice([-1, 1], 10)
Wd = torch.from_numpy(W).to(dev)
report("cfg3_matmul_64x128x64_sparse_ternary", timed(lambda: E.matmul_ct_clear(x, Wd, 64, 128, 64, L)),
       (64 * 128 + 64 * 64) * L * 8 + 128 * 64 * 8, "tap-list kernel (x read once, bulk-copied tiles), 10 non-zero weights per column")
Wfull = torch.from_numpy(rng.integers(-3, 4, (128, 64)).astype(np.int64)).to(dev)
report("matmul_64x128x64_dense_weights", timed(lambda: E.matmul_ct_clear(x, Wfull, 64, 128, 64, L)),
       (64 * 128 + 64 * 64) * L * 8 + 128 * 64 * 8, "same shape, dense weights: 4.3e9 u64 multiply-adds -> bound by the integer pipe (register-blocked kernel)")
del x
# cfg4: conv2d 1x8x28x28, 3x3, pad 1 (72 taps per output) through gather-MAC
xin = rand_ct(8 * 28 * 28, L)
idx = np.full((8, 28, 28, 72), -1, dtype=np.int32)
w = np.zeros((8, 28, 28, 72), dtype=np.int64)
Wc = rng.integers(-1, 2, (8, 8, 3, 3))
for f in range(8):
    for c in range(8):
        for kh in range(3):
            for kw in range(3):
                kk = (c * 3 + kh) * 3 + kw
                for oh in range(28):
                    ih = oh - 1 + kh
                    if not 0 <= ih < 28:
                        continue
                    for ow in range(28):
                        iw = ow - 1 + kw
                        if 0 <= iw < 28:
                            idx[f, oh, ow, kk] = (c * 28 + ih) * 28 + iw
                            w[f, oh, ow, kk] = Wc[f, c, kh, kw]
di, dw = torch.from_numpy(idx.reshape(-1, 72)).to(dev), torch.from_numpy(w.reshape(-1, 72)).to(dev)
report("cfg4_conv2d_8x28x28_3x3_gather_mac", timed(lambda: E.lin_gather_mac(xin, di, dw, L)), 2 * 6272 * L * 8,
       f"{int((w != 0).sum())} u64 MACs x {L} words: INT-ALU bound")
# elementwise at the cfg2 shape (4096 ciphertexts of N = 32768): add, add-plain, mul-clear, neg
L2 = 32769
a, b = rand_ct(4096, L2), rand_ct(4096, L2)
report("add_eint_4096x32769_flat", timed(lambda: E.lin_add(a, b, L2)), 3 * 4096 * L2 * 8)
ident = torch.arange(4096, dtype=torch.int32, device=dev)
report("add_eint_4096x32769_row_mapped", timed(lambda: E.lin_add(a, b, L2, ident, ident, 4096)), 3 * 4096 * L2 * 8)
pt = torch.randint(0, 1 << 62, (4096,), dtype=torch.int64, device=dev)
report("add_eint_int_4096x32769", timed(lambda: E.lin_add_plain(a, pt, L2)), 2 * 4096 * L2 * 8)
report("mul_eint_int_4096x32769", timed(lambda: E.lin_mul_clear(a, pt, L2)), 2 * 4096 * L2 * 8)
report("neg_eint_4096x32769", timed(lambda: E.lin_neg(a, L2)), 2 * 4096 * L2 * 8)
rev = torch.arange(4095, -1, -1, dtype=torch.int32, device=dev)
report("add_eint_4096x32769_reversed_rows", timed(lambda: E.lin_add(a, b, L2, rev, ident, 4096)), 3 * 4096 * L2 * 8,
       "operand rows on the other 16-byte phase: 8-byte accesses")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/op_profile.json", "w"), indent=1)
