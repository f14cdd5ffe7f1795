"""The drop-in, proven through the REFERENCE'S OWN classes: the unmodified `concrete.numpy` package from
/root/reference (Compiler, Circuit, Client, Server, ClientSpecs, EvaluationKeys) runs on top of this backend's
`concrete.compiler` module (concrete-numpy_b200/compiler.py) and MLIR stand-ins (mlir_text/).  The test bodies restate
/root/reference/tests/compilation/test_circuit.py (bad inputs :60-129, client/server API :132-247, feedback :37-58,
unused argument :306-326, dataflow :329-345) and tests/mlir/test_graph_converter.py:462-501 (golden MLIR).

No GPU here: the arithmetic engine behind compiler.py / executor.py is swapped for tests/clear_engine.py (one clear
torus word per ciphertext), so this pins the host logic -- options, parameter search, client parameters, argument
validation, (un)serialisation, save / load, executor -- not the kernels (tests/test_golden_gpu.py does that).
/root/reference exists in the development container only; the test is skipped elsewhere."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_frontend  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_frontend.available(), reason="reference frontend (/root/reference) not present")


@pytest.fixture()
def cnp(monkeypatch, tmp_path):
    import clear_engine
    from concrete_numpy_b200 import compiler as cc, executor as X

    mod, _ = ref_frontend.load()
    monkeypatch.setattr(cc, "E", clear_engine)
    monkeypatch.setattr(X, "E", clear_engine)
    monkeypatch.setattr(cc, "default_device", lambda: torch.device("cpu"))
    monkeypatch.delenv("CNP_B200_KEY_SEED", raising=False)
    mod._test_cache = str(tmp_path / "keycache")
    return mod


def configuration(cnp):
    """tests/conftest.py:109-119 of the reference"""
    return cnp.Configuration(
        dump_artifacts_on_unexpected_failures=False, enable_unsafe_features=True, use_insecure_key_cache=True,
        loop_parallelize=True, dataflow_parallelize=False, auto_parallelize=False, jit=True,
        insecure_key_cache_location=cnp._test_cache, global_p_error=(1 / 10_000),
    )


def test_reference_classes_are_the_unmodified_ones(cnp):
    import concrete.compiler as wheel
    from concrete.numpy.compilation import circuit, client, server, specs

    from concrete_numpy_b200 import compiler as cc

    assert wheel is cc
    for m in (circuit, client, server, specs):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(ref_frontend.REFERENCE))
    assert server.JITSupport is cc.JITSupport and client.ClientSupport is cc.ClientSupport


def test_circuit_bad_run(cnp):
    @cnp.compiler({"x": "encrypted", "y": "encrypted"})
    def f(x, y):
        return x + y

    rng = np.random.default_rng(0)
    inputset = [(int(rng.integers(0, 2**4)), int(rng.integers(0, 2**5))) for _ in range(100)] + [(15, 31)]
    circuit = f.compile(inputset, configuration(cnp))
    for args, msg in [((1,), "Expected 2 inputs but got 1"), ((1, 2, 3), "Expected 2 inputs but got 3"),
                      ((-1, 11), "Expected argument 0 to be EncryptedScalar<uint6> but it's EncryptedScalar<int1>"),
                      ((1, -11), "Expected argument 1 to be EncryptedScalar<uint6> but it's EncryptedScalar<int5>"),
                      ((100, 10), "Expected argument 0 to be EncryptedScalar<uint6> but it's EncryptedScalar<uint7>"),
                      ((1, 100), "Expected argument 1 to be EncryptedScalar<uint6> but it's EncryptedScalar<uint7>")]:
        with pytest.raises(ValueError) as excinfo:
            circuit.encrypt_run_decrypt(*args)
        assert str(excinfo.value) == msg
    assert circuit.encrypt_run_decrypt(7, 20) == 27


def test_circuit_feedback(cnp):
    @cnp.compiler({"x": "encrypted", "y": "encrypted"})
    def f(x, y):
        return np.sqrt(((x + y) ** 2) + 10).astype(np.int64)

    inputset = [(i % 4, (i // 4) % 4) for i in range(100)]
    circuit = f.compile(inputset, configuration(cnp), p_error=0.1, global_p_error=0.05)
    assert isinstance(circuit.complexity, float)
    for name in ("size_of_secret_keys", "size_of_bootstrap_keys", "size_of_keyswitch_keys", "size_of_inputs", "size_of_outputs"):
        assert isinstance(getattr(circuit, name), int)
    assert isinstance(circuit.p_error, float) and isinstance(circuit.global_p_error, float)
    assert circuit.p_error <= 0.1 and circuit.global_p_error <= 0.05
    assert circuit.encrypt_run_decrypt(3, 2) == int(np.sqrt((3 + 2) ** 2 + 10))


@pytest.mark.parametrize("via_mlir", [False, True])
def test_client_server_api(cnp, via_mlir):
    from concrete.numpy import Client, ClientSpecs, EvaluationKeys, Server

    cfg = configuration(cnp)

    @cnp.compiler({"x": "encrypted"})
    def function(x):
        return x + 42

    rng = np.random.default_rng(1)
    inputset = [rng.integers(0, 10, size=(3,)) for _ in range(10)] + [np.array([9, 9, 9]), np.array([0, 0, 0])]
    circuit = function.compile(inputset, cfg.fork(jit=False))
    circuit.keygen()
    with tempfile.TemporaryDirectory() as tmp_dir:
        tmp_dir_path = Path(tmp_dir)
        server_path = tmp_dir_path / "server.zip"
        circuit.server.save(server_path, via_mlir=via_mlir)
        client_path = tmp_dir_path / "client.zip"
        circuit.client.save(client_path)
        circuit.cleanup()
        server = Server.load(server_path)
        client_specs = ClientSpecs.unserialize(server.client_specs.serialize())
        clients = [Client(client_specs, cfg.insecure_key_cache_location),
                   Client.load(client_path, cfg.insecure_key_cache_location)]
        for client in clients:
            args = client.encrypt([3, 8, 1])
            serialized_args = client.specs.serialize_public_args(args)
            serialized_evaluation_keys = client.evaluation_keys.serialize()
            unserialized_args = server.client_specs.unserialize_public_args(serialized_args)
            unserialized_evaluation_keys = EvaluationKeys.unserialize(serialized_evaluation_keys)
            result = server.run(unserialized_args, unserialized_evaluation_keys)
            serialized_result = server.client_specs.serialize_public_result(result)
            output = client.decrypt(client.specs.unserialize_public_result(serialized_result))
            assert np.array_equal(output, [45, 50, 43])
        if not via_mlir:
            with pytest.raises(RuntimeError) as excinfo:
                server.save("UNUSED", via_mlir=True)
            assert str(excinfo.value) == "Loaded server objects cannot be saved again via MLIR"
        server.cleanup()


def test_bad_server_save(cnp):
    @cnp.compiler({"x": "encrypted"})
    def function(x):
        return x + 42

    circuit = function.compile(range(10), configuration(cnp))
    with pytest.raises(RuntimeError) as excinfo:
        circuit.server.save("test.zip")
    assert str(excinfo.value) == "Just-in-Time compilation cannot be saved"


def test_circuit_run_with_unused_arg_and_dataflow(cnp):
    @cnp.compiler({"x": "encrypted", "y": "encrypted"})
    def f(x, y):  # pylint: disable=unused-argument
        return x + 10

    inputset = [(8 + i % 8, 16 + i % 16) for i in range(100)]
    circuit = f.compile(inputset, configuration(cnp))
    with pytest.raises(ValueError, match="Expected 2 inputs but got 1"):
        circuit.encrypt_run_decrypt(10)
    assert [circuit.encrypt_run_decrypt(10, y) for y in (0, 10, 20)] == [20, 20, 20]

    @cnp.compiler({"x": "encrypted", "y": "encrypted"})
    def g(x, y):
        return (x**2) + (y // 2)

    inputset = [(i % 8, (i // 8) % 8) for i in range(100)]
    circuit = g.compile(inputset, configuration(cnp).fork(dataflow_parallelize=True))
    assert circuit.encrypt_run_decrypt(5, 6) == 28


def test_tensor_circuit_with_tables_through_reference_classes(cnp):
    """A multi-op tensor circuit (lookup + matmul + comparison lowering) through the reference Circuit."""
    table = cnp.LookupTable([2, 0, 3, 1, 1, 3, 0, 2])

    @cnp.compiler({"x": "encrypted", "y": "encrypted"})
    def f(x, y):
        return (table[x] @ np.array([[1, 0], [1, 2]])) + (x < y)

    rng = np.random.default_rng(3)
    inputset = [(rng.integers(0, 8, size=(2, 2)), rng.integers(0, 8, size=(2, 2))) for _ in range(30)]
    inputset += [(np.full((2, 2), 7), np.zeros((2, 2), dtype=np.int64)), (np.zeros((2, 2), dtype=np.int64), np.full((2, 2), 7))]
    circuit = f.compile(inputset, configuration(cnp))
    for x, y in inputset[:6]:
        expected = (np.array([2, 0, 3, 1, 1, 3, 0, 2])[x] @ np.array([[1, 0], [1, 2]])) + (x < y)
        assert np.array_equal(circuit.encrypt_run_decrypt(x, y), expected)


GOLDEN_MLIR = """
module {
  func.func @main(%arg0: !FHE.eint<3>) -> !FHE.eint<3> {
    %c1_i4 = arith.constant 1 : i4
    %cst = arith.constant dense<[4, 1, 2, 3, 3, 3, 3, 3]> : tensor<8xi64>
    %0 = "FHE.apply_lookup_table"(%arg0, %cst) : (!FHE.eint<3>, tensor<8xi64>) -> !FHE.eint<3>
    %1 = "FHE.add_eint_int"(%arg0, %c1_i4) : (!FHE.eint<3>, i4) -> !FHE.eint<3>
    %2 = "FHE.add_eint_int"(%0, %c1_i4) : (!FHE.eint<3>, i4) -> !FHE.eint<3>
    %3 = "FHE.apply_lookup_table"(%1, %cst) : (!FHE.eint<3>, tensor<8xi64>) -> !FHE.eint<3>
    %4 = "FHE.add_eint"(%2, %3) : (!FHE.eint<3>, !FHE.eint<3>) -> !FHE.eint<3>
    return %4 : !FHE.eint<3>
  }
}
"""


def test_shim_reproduces_the_reference_golden_mlir(cnp):
    """tests/mlir/test_graph_converter.py:462-501 (test_constant_cache): the MLIR text emitted through mlir_text/ir.py
    is the reference's golden text, character for character (modulo surrounding whitespace, as `helpers.check_str`)."""
    compiler = cnp.Compiler(lambda x: 1 + cnp.LookupTable([4, 1, 2, 3])[x] + cnp.LookupTable([4, 1, 2, 3])[x + 1],
                            {"x": "encrypted"})
    circuit = compiler.compile(range(3), configuration(cnp))
    got = [line.rstrip() for line in circuit.mlir.strip().splitlines()]
    exp = [line.rstrip() for line in GOLDEN_MLIR.strip().splitlines()]
    assert got == exp
    assert [circuit.encrypt_run_decrypt(x) for x in range(3)] == [1 + [4, 1, 2, 3][x] + [4, 1, 2, 3][x + 1] for x in range(3)]
