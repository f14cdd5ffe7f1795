"""CPU tests of the host logic against the reference-generated golden circuits: the MLIR text the
reference frontend emitted is parsed, analysed and executed by executor.py on clear torus words
(tests/clear_engine.py); results must equal the reference's clear graph evaluation bit for bit."""
import numpy as np
import pytest
import torch

from concrete_numpy_b200 import executor as X
from concrete_numpy_b200.engine import CryptoParams
from concrete_numpy_b200.mlir_text.parser import parse_module

import clear_engine
from golden_util import golden_names, load_golden, same, to_value

SMALL = [n for n in golden_names() if not n.startswith("cfg5_kvdb128")]


@pytest.fixture()
def clear_backend(monkeypatch):
    monkeypatch.setattr(X, "E", clear_engine)


def run_clear(doc, args):
    program = parse_module(doc["mlir"])
    info = X.analyze(program)
    p = info.precision
    params = CryptoParams(p=p, n=1, k=0, N=1 << 11, pbs_level=1, pbs_base_log=1, ks_level=1, ks_base_log=1,
                          lwe_std=0.0, glwe_std=0.0)
    ex = X.Executor(program, params, "cpu")
    vals = []
    for (name, t), a, signed in zip(program.args, args, doc["input_signs"]):
        a = np.asarray(a, dtype=np.int64)
        if t.encrypted:
            off = (1 << (t.width - 1)) if signed else 0            # client.py:161-168
            m = ((a + off).astype(np.uint64) << np.uint64(63 - p)).reshape(-1, 1)
            vals.append(X.Ct(torch.from_numpy(m.view(np.int64)), t.dims))
        else:
            vals.append(a)
    outs = ex.run(clear_engine.ClearEval(params), vals)
    res = []
    for o, t, signed in zip(outs, program.result_types, doc["output_signs"]):
        m = clear_engine.decode(o.data.numpy().view(np.uint64)[:, -1], p).astype(np.int64)
        if signed:                                                  # client.py:237-250
            m %= 1 << t.width
            m = np.where(m < (1 << (t.width - 1)), m, m - (1 << t.width))
        res.append(int(m[0]) if t.dims == () else m.reshape(t.dims))
    return (res[0] if len(res) == 1 else tuple(res)), info


@pytest.mark.parametrize("name", SMALL)
def test_golden_clear_execution(name, clear_backend):
    doc = load_golden(name)
    for case in doc["cases"]:
        got, _ = run_clear(doc, [to_value(a) for a in case["args"]])
        assert same(got, to_value(case["expected"])), name


def test_pbs_counts_match_survey(clear_backend):
    """PBS counts / depths probed on the reference graphs."""
    expect = {"cfg1_lut4_1024": (1024, 1, 4), "cfg2_univariate8_64x64": (8192, 2, 8),
              "cfg3_matmul6_relu": (4096, 1, 6), "cfg4_conv2d_relu": (12544, 2, 6)}
    for name, (n_pbs, depth, p) in expect.items():
        info = X.analyze(parse_module(load_golden(name)["mlir"]))
        assert (info.n_pbs, info.tlu_depth, info.precision) == (n_pbs, depth, p), name


WIDE = golden_names("wide_") + ["cfg3b_matmul6_dense_relu"]


@pytest.mark.parametrize("name", WIDE)
def test_golden_wide_crt_lanes(name, clear_backend, monkeypatch):
    """9..16-bit circuits on clear CRT residues: residue encodings, per-residue weight reduction, lane execution,
    table expansion over the extracted-bit patterns and CRT decoding reproduce the reference's clear evaluation."""
    clear_engine.install_wide(monkeypatch)
    doc = load_golden(name)
    for case in doc["cases"]:
        got = clear_engine.run_crt(X, doc, [to_value(a) for a in case["args"]])
        assert same(got, to_value(case["expected"])), name


def test_every_operator_family_runs_on_crt_lanes(clear_backend, monkeypatch):
    """All small reference-generated circuits, forced onto CRT residues (moduli for their own width): every executor
    handler (broadcasting, indexing, concat, conv2d, sums, mapped tables, ...) works lane-wise and the table expansion
    over extracted-bit patterns reproduces the reference's clear evaluation."""
    clear_engine.install_wide(monkeypatch)
    names = [n for n in SMALL if not n.startswith(("cfg1", "cfg2", "cfg3_", "cfg4", "lin_"))]
    assert len(names) > 100
    for name in names:
        doc = load_golden(name)
        case = doc["cases"][0]
        got = clear_engine.run_crt(X, doc, [to_value(a) for a in case["args"]])
        assert same(got, to_value(case["expected"])), name
