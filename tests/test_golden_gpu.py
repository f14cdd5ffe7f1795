"""GPU parity tests proper: every reference-generated golden circuit (MLIR text emitted by the
reference frontend) is compiled by this backend, keys are generated on device, inputs encrypted,
the program executed by the CUDA kernels through the C ABI and the result decrypted; the decrypted
output must equal the reference's clear graph evaluation BIT FOR BIT (tests/conftest.py:221-287 of
the reference, `check_execution`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from concrete_numpy_b200 import api

from golden_util import golden_names, load_golden, same, to_value

os.environ.setdefault("CNP_B200_KEY_SEED", "20261019")

FAST = [n for n in golden_names() if n.startswith(("ops_", "ka_", "lin_", "cfg5_kvdb8", "cfg1", "cfg3"))]
# full parameterisations of the reference's comparison / shift / bitwise / matmul / sum / static indexing / static
# assignment tests, generated from the reference's own parametrize lists (oracle/gen_golden.py: sweep)
SWEEP = [n for n in golden_names() if n.startswith("sw_")]
HEAVY = [n for n in golden_names() if n.startswith(("cfg2", "cfg4", "cfg5_kvdb128"))]


WIDE = [n for n in golden_names() if n.startswith("wide_")]      # (cfg3b, 11 bits at full size, runs with FAST)      # 9..16-bit circuits: CRT residues + wide lookup


def check(name):
    doc = load_golden(name)
    cfg = api.Configuration(jit=True, global_p_error=doc["global_p_error"])
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], cfg)
    for case in doc["cases"]:
        args = [to_value(a) for a in case["args"]]
        got = circuit.encrypt_run_decrypt(*args)
        assert same(got, to_value(case["expected"])), name


@pytest.mark.parametrize("name", FAST)
def test_golden_encrypted_execution(cuda_device, name):
    check(name)


@pytest.mark.parametrize("name", SWEEP)
def test_golden_reference_parameter_sweep(cuda_device, name):
    check(name)


@pytest.mark.parametrize("name", HEAVY)
def test_golden_encrypted_execution_baseline_sizes(cuda_device, name):
    check(name)


@pytest.mark.parametrize("name", WIDE)
def test_golden_wide_integers(cuda_device, name):
    check(name)


def test_known_answer_client_server_roundtrip(cuda_device, tmp_path):
    """reference tests/compilation/test_circuit.py:132-192: [3, 8, 1] + 42 -> [45, 50, 43] through
    save/load and bytes serialisation of arguments, evaluation keys and results."""
    doc = load_golden("ka_client_server")
    cfg = api.Configuration(jit=False, global_p_error=doc["global_p_error"])
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], cfg)
    server_path = tmp_path / "server.zip"
    circuit.server.save(server_path)
    client_path = tmp_path / "client.zip"
    circuit.client.save(client_path)
    circuit.cleanup()

    server = api.Server.load(server_path)
    serialized_specs = server.client_specs.serialize()
    client_specs = api.ClientSpecs.unserialize(serialized_specs)
    clients = [api.Client(client_specs, tmp_path / "keys"), api.Client.load(client_path, tmp_path / "keys")]
    for client in clients:
        args = client.encrypt(np.array([3, 8, 1]))
        serialized_args = client.specs.serialize_public_args(args)
        serialized_keys = client.evaluation_keys.serialize()
        from concrete_numpy_b200.compiler import EvaluationKeys

        un_args = server.client_specs.unserialize_public_args(serialized_args)
        un_keys = EvaluationKeys.unserialize(serialized_keys)
        result = server.run(un_args, un_keys)
        serialized_result = server.client_specs.serialize_public_result(result)
        un_result = client.specs.unserialize_public_result(serialized_result)
        assert np.array_equal(client.decrypt(un_result), [45, 50, 43])
    server.cleanup()

    # via MLIR
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], cfg)
    circuit.server.save(tmp_path / "server_mlir.zip", via_mlir=True)
    server2 = api.Server.load(tmp_path / "server_mlir.zip")
    client = api.Client(server2.client_specs)
    res = server2.run(client.encrypt(np.array([3, 8, 1])), client.evaluation_keys)
    assert np.array_equal(client.decrypt(res), [45, 50, 43])


def test_bad_arguments(cuda_device):
    """reference tests/compilation/test_circuit.py:61-129 (error strings)."""
    doc = load_golden("ops_add_enc_enc")
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], api.Configuration(jit=True))
    with pytest.raises(ValueError) as e:
        circuit.encrypt_run_decrypt(1)
    assert str(e.value) == "Expected 2 inputs but got 1"
    with pytest.raises(ValueError) as e:
        circuit.encrypt_run_decrypt(-1, 11)
    assert str(e.value) == "Expected argument 0 to be EncryptedScalar<uint7> but it's EncryptedScalar<int1>"
    with pytest.raises(ValueError) as e:
        circuit.encrypt_run_decrypt(1, 1000)
    assert str(e.value) == "Expected argument 1 to be EncryptedScalar<uint7> but it's EncryptedScalar<uint10>"
    with pytest.raises(ValueError) as e:
        circuit.encrypt_run_decrypt(np.array([1, 2]), 3)
    assert str(e.value) == (
        "Expected argument 0 to be EncryptedScalar<uint7> but it's EncryptedTensor<uint2, shape=(2,)>")
    assert isinstance(circuit.complexity, float) and isinstance(circuit.size_of_bootstrap_keys, int)
    assert circuit.global_p_error <= 1e-5 * 1.0001
