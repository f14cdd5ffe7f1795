"""world_size-2 (and 3) gloo tests of the N>1 host logic: every encrypted tensor is sharded by ciphertext over the
ranks, table lookups and element-wise ops run on the local block, ops that read other elements all-gather their
operand, sums over the sharded axis all-reduce partial sums; every rank must obtain the reference's clear result
for every golden circuit (all operator families), and the collectives counted must match the design."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _run_clear(X, clear_engine, doc, group, to_value):
    from concrete_numpy_b200.engine import CryptoParams
    from concrete_numpy_b200.mlir_text.parser import parse_module

    program = parse_module(doc["mlir"])
    p = X.analyze(program).precision
    params = CryptoParams(p=p, n=1, k=0, N=2048, pbs_level=1, pbs_base_log=1, ks_level=1, ks_base_log=1,
                          lwe_std=0.0, glwe_std=0.0)
    ex = X.Executor(program, params, "cpu", group=group)
    results = []
    for case in doc["cases"]:
        vals = []
        for (_, t), a, signed in zip(program.args, [to_value(a) for a in case["args"]], doc["input_signs"]):
            a = np.asarray(a, dtype=np.int64)
            if not t.encrypted:
                vals.append(a)
                continue
            off = (1 << (t.width - 1)) if signed else 0
            m = ((a + off).astype(np.uint64) << np.uint64(63 - p)).reshape(-1, 1)
            vals.append(X.Ct(torch.from_numpy(m.view(np.int64)), t.dims))
        outs = ex.run(clear_engine.ClearEval(params), vals)
        res = []
        for o, t, signed in zip(outs, program.result_types, doc["output_signs"]):
            assert not o.sharded and o.data.shape[0] == max(t.size, 1)
            m = clear_engine.decode(o.data.numpy().view(np.uint64)[:, -1], p).astype(np.int64)
            if signed:
                m %= 1 << t.width
                m = np.where(m < (1 << (t.width - 1)), m, m - (1 << t.width))
            res.append(int(m[0]) if t.dims == () else m.reshape(t.dims))
        results.append((res[0] if len(res) == 1 else tuple(res), to_value(case["expected"]), dict(ex.last_stats)))
    return results


def _worker(rank, world, port, names, shard_min, failures, stats_out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CNP_B200_SHARD_MIN"] = str(shard_min)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clear_engine
    from concrete_numpy_b200 import executor as X
    from golden_util import load_golden, same, to_value

    X.E = clear_engine
    try:
        for name in names:
            doc = load_golden(name)
            if doc.get("crypto", {}).get("moduli") or name.startswith(("wide_", "lin_")):
                continue
            for got, exp, st in _run_clear(X, clear_engine, doc, dist.group.WORLD, to_value):
                if not same(got, exp):
                    failures.put((rank, name))
                if rank == 0:
                    stats_out.put((name, st))
        # wide integers: every rank looks up its slice of the elements of all residues, then all-gathers per residue
        clear_engine.install_wide()
        for name in ("wide_matmul_relu_14bit", "wide_multi_table_9bit", "wide_linear_tlu_16bit"):
            if name not in names:
                continue
            doc = load_golden(name)
            for case in doc["cases"]:
                got = clear_engine.run_crt(X, doc, [to_value(a) for a in case["args"]], group=dist.group.WORLD)
                if not same(got, to_value(case["expected"])):
                    failures.put((rank, name))
    finally:
        dist.destroy_process_group()


def _launch(world, names, shard_min):
    ctx = mp.get_context("spawn")
    failures, stats = ctx.Queue(), ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, shard_min, failures, stats)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    # drain the statistics while the workers run (a full queue would block them)
    alive = True
    while alive:
        alive = any(p.is_alive() for p in procs)
        while not stats.empty():
            name, st = stats.get()
            out.setdefault(name, st)
        if alive:
            procs[0].join(0.05)
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert failures.empty(), failures.get()
    return out


def test_sharded_circuits_world2():
    """BASELINE-shaped circuits at the default sharding threshold: what is exchanged is what SURVEY 8e predicts."""
    names = ["cfg1_lut4_1024", "cfg3_matmul6_relu", "cfg4_conv2d_relu", "cfg5_kvdb128_query", "cfg5_kvdb8_insert",
             "cfg5_kvdb8_query", "ops_uni_chain", "wide_matmul_relu_14bit", "wide_multi_table_9bit", "wide_linear_tlu_16bit"]
    st = _launch(2, names, 64)
    # cfg1: one lookup on 1024 ciphertexts -> only the final gather of the result
    assert st["cfg1_lut4_1024"]["allgather_calls"] == 1 and st["cfg1_lut4_1024"]["allreduce_calls"] == 0
    # cfg3: x @ W sharded by rows of x, + 32, ReLU lookup: no exchange before the result is returned
    assert st["cfg3_matmul6_relu"]["allgather_calls"] == 1, st["cfg3_matmul6_relu"]
    # cfg4: the first lookup feeds nothing else; conv2d consumes the (replicated) input, lookups stay sharded
    assert st["cfg4_conv2d_relu"]["allgather_calls"] == 1, st["cfg4_conv2d_relu"]
    # key-value query: np.sum(axis=0) over the sharded entry axis is an all-reduce of partial sums
    assert st["cfg5_kvdb128_query"]["allreduce_calls"] >= 1, st["cfg5_kvdb128_query"]
    # the sequential insert chain (8 scalar lookups) is replicated; only its two 8x8 lookups (64 ciphertexts) are sharded
    assert st["cfg5_kvdb8_insert"]["allgather_calls"] <= 2 and st["cfg5_kvdb8_insert"]["allreduce_calls"] == 0
    # the 8-entry query (lookups on <= 64 ciphertexts ... all below or at the threshold) never all-reduces more than its sums
    assert st["cfg5_kvdb8_query"]["allgather_calls"] <= 3


@pytest.mark.parametrize("world", [2, 3])
def test_every_operator_family_sharded(world):
    """All golden circuits with the threshold forced down to 2 ciphertexts: every handler runs its sharded path
    (uneven and empty blocks included) and every rank still gets the reference's clear result."""
    sys.path.insert(0, HERE)
    from golden_util import golden_names

    names = [n for n in golden_names() if not n.startswith(("cfg2", "cfg3b", "cfg5_kvdb128", "wide_", "lin_", "cfg4", "cfg3"))]
    _launch(world, names, 2)


CHAIN_MLIR = """module {
  func.func @main(%arg0: tensor<8x4x!FHE.eint<5>>) -> tensor<2x!FHE.eint<5>> {
    %t = arith.constant dense<[0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3]> : tensor<32xi64>
    %w = arith.constant dense<[[1, 0], [0, 1], [1, 1], [0, 2]]> : tensor<4x2xi6>
    %0 = "FHELinalg.apply_lookup_table"(%arg0, %t) : (tensor<8x4x!FHE.eint<5>>, tensor<32xi64>) -> tensor<8x4x!FHE.eint<5>>
    %1 = "FHELinalg.matmul_eint_int"(%0, %w) : (tensor<8x4x!FHE.eint<5>>, tensor<4x2xi6>) -> tensor<8x2x!FHE.eint<5>>
    %2 = "FHELinalg.apply_lookup_table"(%1, %t) : (tensor<8x2x!FHE.eint<5>>, tensor<32xi64>) -> tensor<8x2x!FHE.eint<5>>
    %3 = "FHELinalg.sum"(%2) {axes = [0], keep_dims = false} : (tensor<8x2x!FHE.eint<5>>) -> tensor<2x!FHE.eint<5>>
    return %3 : tensor<2x!FHE.eint<5>>
  }
}"""


def _chain_worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CNP_B200_SHARD_MIN"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clear_engine
    from concrete_numpy_b200 import executor as X
    from golden_util import to_value

    X.E = clear_engine
    try:
        x = np.arange(32).reshape(8, 4) % 7
        doc = {"mlir": CHAIN_MLIR, "input_signs": [False], "output_signs": [False],
               "cases": [{"args": [{"data": x.reshape(-1).tolist(), "shape": [8, 4]}], "expected": {"data": [0, 0], "shape": [2]}}]}
        (got, _, st), = _run_clear(X, clear_engine, doc, dist.group.WORLD, to_value)
        out.put((rank, np.asarray(got).tolist(), st))
    finally:
        dist.destroy_process_group()


def test_lookup_matmul_sum_chain_stays_sharded():
    """lookup -> matmul (rows inside the blocks: local) -> lookup -> sum over the sharded axis (all-reduce): the only
    collective of the whole circuit is ONE all-reduce of the two result ciphertexts; no all-gather at all."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_chain_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    x = np.arange(32).reshape(8, 4) % 7
    w = np.array([[1, 0], [0, 1], [1, 1], [0, 2]])
    exp = (((x % 4) @ w) % 4).sum(axis=0)
    for rank, res, st in got:
        assert res == exp.tolist(), (rank, res, exp)
        assert st["allgather_calls"] == 0 and st["allreduce_calls"] == 1, st
