"""world_size-2 gloo test of the N>1 host logic: table lookups sharded by ciphertext over the
ranks + all-gather, linear ops replicated; every rank must obtain the reference's clear result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, names, failures):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clear_engine
    from concrete_numpy_b200 import executor as X
    from concrete_numpy_b200.engine import CryptoParams
    from concrete_numpy_b200.mlir_text.parser import parse_module
    from golden_util import load_golden, same, to_value

    X.E = clear_engine
    try:
        for name in names:
            doc = load_golden(name)
            program = parse_module(doc["mlir"])
            p = X.analyze(program).precision
            params = CryptoParams(p=p, n=1, k=0, N=2048, pbs_level=1, pbs_base_log=1, ks_level=1, ks_base_log=1,
                                  lwe_std=0.0, glwe_std=0.0)
            ex = X.Executor(program, params, "cpu", group=dist.group.WORLD)
            for case in doc["cases"]:
                vals = []
                for (_, t), a, signed in zip(program.args, [to_value(a) for a in case["args"]], doc["input_signs"]):
                    a = np.asarray(a, dtype=np.int64)
                    off = (1 << (t.width - 1)) if signed else 0
                    m = ((a + off).astype(np.uint64) << np.uint64(63 - p)).reshape(-1, 1)
                    vals.append(X.Ct(torch.from_numpy(m.view(np.int64)), t.dims))
                outs = ex.run(clear_engine.ClearEval(params), vals)
                res = []
                for o, t, signed in zip(outs, program.result_types, doc["output_signs"]):
                    m = clear_engine.decode(o.data.numpy().view(np.uint64)[:, -1], p).astype(np.int64)
                    if signed:
                        m %= 1 << t.width
                        m = np.where(m < (1 << (t.width - 1)), m, m - (1 << t.width))
                    res.append(int(m[0]) if t.dims == () else m.reshape(t.dims))
                got = res[0] if len(res) == 1 else tuple(res)
                if not same(got, to_value(case["expected"])):
                    failures.put((rank, name))
        # wide integers: every rank looks up its slice of the elements of all residues, then all-gathers per residue
        clear_engine.install_wide()
        for name in ("wide_matmul_relu_14bit", "wide_multi_table_9bit", "wide_linear_tlu_16bit"):
            doc = load_golden(name)
            for case in doc["cases"]:
                got = clear_engine.run_crt(X, doc, [to_value(a) for a in case["args"]], group=dist.group.WORLD)
                if not same(got, to_value(case["expected"])):
                    failures.put((rank, name))
    finally:
        dist.destroy_process_group()


def test_sharded_table_lookups_world2():
    names = ["cfg1_lut4_1024", "ops_tlu_multi_table", "ops_uni_chain", "cfg5_kvdb8_query", "cfg5_kvdb8_insert",
             "ops_less_enc_chunked", "ops_tlu_scalar"]
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, names, failures)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert failures.empty(), failures.get()
