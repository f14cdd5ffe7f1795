"""
Host-side mirror of the reference's compilation layer for the hot path, usable without the
reference being installed: `ClientSpecs`, `Client`, `Server`, `Circuit` with the same method
names, argument meaning and error behaviour as

  concrete/numpy/compilation/specs.py:11-127      ClientSpecs
  concrete/numpy/compilation/client.py:24-271     Client  (validation + signed offset + decode)
  concrete/numpy/compilation/server.py:31-350     Server  (create / run / save / load / feedback)
  concrete/numpy/compilation/circuit.py:19-218    Circuit (keygen / encrypt / run / decrypt)

The input of `Server.create` is the MLIR text the reference frontend emits (`Circuit.mlir`); when
the reference itself is installed it runs unmodified on `compiler.py` and this module is unused.
"""

import json
import os
import shutil
import tempfile
from dataclasses import asdict, dataclass
from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np

from . import compiler as cc

DEFAULT_GLOBAL_P_ERROR = 1 / 100_000      # reference configuration.py:9-10


@dataclass
class Configuration:
    """Subset of the reference's Configuration that reaches the hot path (configuration.py:20-34)."""

    jit: bool = False
    loop_parallelize: bool = True
    dataflow_parallelize: bool = False
    auto_parallelize: bool = False
    p_error: Optional[float] = None
    global_p_error: Optional[float] = None
    use_insecure_key_cache: bool = False
    insecure_key_cache_location: Optional[str] = None
    show_optimizer: bool = False


def _int_dtype_name(lo: int, hi: int) -> str:
    """Smallest (u)intN able to hold [lo, hi], named like the reference's Integer.__str__."""
    if lo < 0:
        bits = 1
        while not (-(1 << (bits - 1)) <= lo and hi <= (1 << (bits - 1)) - 1):
            bits += 1
        return f"int{bits}"
    bits = 1
    while hi > (1 << bits) - 1:
        bits += 1
    return f"uint{bits}"


def _describe(dtype: str, shape: Tuple[int, ...], encrypted: bool) -> str:
    kind = "Encrypted" if encrypted else "Clear"
    if shape == ():
        return f"{kind}Scalar<{dtype}>"
    return f"{kind}Tensor<{dtype}, shape={shape}>"


class ClientSpecs:
    def __init__(self, input_signs: List[bool], client_parameters: cc.ClientParameters, output_signs: List[bool]):
        self.input_signs = list(input_signs)
        self.client_parameters = client_parameters
        self.output_signs = list(output_signs)

    def serialize(self) -> str:
        return json.dumps({"input_signs": self.input_signs,
                           "client_parameters": json.loads(self.client_parameters.serialize()),
                           "output_signs": self.output_signs})

    @staticmethod
    def unserialize(text: str) -> "ClientSpecs":
        raw = json.loads(text)
        cp = cc.ClientParameters.unserialize(json.dumps(raw["client_parameters"]).encode("utf-8"))
        return ClientSpecs(raw["input_signs"], cp, raw["output_signs"])

    def serialize_public_args(self, args: cc.PublicArguments) -> bytes:
        return args.serialize()

    def unserialize_public_args(self, data: bytes) -> cc.PublicArguments:
        return cc.PublicArguments.unserialize(self.client_parameters, data)

    def serialize_public_result(self, result: cc.PublicResult) -> bytes:
        return result.serialize()

    def unserialize_public_result(self, data: bytes) -> cc.PublicResult:
        return cc.PublicResult.unserialize(self.client_parameters, data)


class Client:
    def __init__(self, client_specs: ClientSpecs, keyset_cache_directory: Optional[Union[str, Path]] = None):
        self.specs = client_specs
        self._keyset: Optional[cc.KeySet] = None
        self._keyset_cache = None
        if keyset_cache_directory is not None:
            self._keyset_cache = cc.KeySetCache.new(str(keyset_cache_directory))

    def save(self, path: Union[str, Path]):
        with tempfile.TemporaryDirectory() as tmp:
            with open(Path(tmp) / "client.specs.json", "w", encoding="utf-8") as f:
                f.write(self.specs.serialize())
            path = str(path)
            shutil.make_archive(path[:-4] if path.endswith(".zip") else path, "zip", tmp)

    @staticmethod
    def load(path: Union[str, Path], keyset_cache_directory: Optional[Union[str, Path]] = None) -> "Client":
        with tempfile.TemporaryDirectory() as tmp:
            shutil.unpack_archive(str(path), tmp, "zip")
            with open(Path(tmp) / "client.specs.json", "r", encoding="utf-8") as f:
                specs = ClientSpecs.unserialize(f.read())
        return Client(specs, keyset_cache_directory)

    def keygen(self, force: bool = False):
        if self._keyset is None or force:
            self._keyset = cc.ClientSupport.key_set(self.specs.client_parameters, self._keyset_cache)

    def encrypt(self, *args: Union[int, np.ndarray]) -> cc.PublicArguments:
        specs = self.specs.client_parameters.input_specs()
        if len(args) != len(specs):
            raise ValueError(f"Expected {len(specs)} inputs but got {len(args)}")
        sanitized = []
        for index, (arg, spec) in enumerate(zip(args, specs)):
            if isinstance(arg, list):
                arg = np.array(arg)
            width = spec["shape"]["width"]
            shape = tuple(spec["shape"]["dimensions"])
            encrypted = spec["encryption"] is not None
            signed = self.specs.input_signs[index]
            lo, hi = (-(1 << (width - 1)), (1 << (width - 1)) - 1) if signed else (0, (1 << width) - 1)
            expected = _describe(("int" if signed else "uint") + str(width), shape, encrypted)
            is_int = isinstance(arg, (int, np.integer)) or (
                isinstance(arg, np.ndarray) and np.issubdtype(arg.dtype, np.integer))
            if not is_int:
                raise ValueError(f"Expected argument {index} to be {expected} but it's {_value_of(arg, encrypted)}")
            a = np.asarray(arg)
            amin, amax = (int(a.min()), int(a.max())) if a.size else (0, 0)
            if amin < lo or amax > hi or a.shape != shape:
                raise ValueError(f"Expected argument {index} to be {expected} but it's {_value_of(arg, encrypted)}")
            offset = (1 << (width - 1)) if signed else 0          # client.py:161-168
            sanitized.append((a.astype(np.int64) + offset).astype(np.uint64))
        self.keygen(force=False)
        return cc.ClientSupport.encrypt_arguments(self.specs.client_parameters, self._keyset, sanitized)

    def decrypt(self, result: cc.PublicResult):
        self.keygen(force=False)
        outputs = cc.ClientSupport.decrypt_result(self.specs.client_parameters, self._keyset, result)
        if not isinstance(outputs, tuple):
            outputs = (outputs,)
        specs = self.specs.client_parameters.output_specs()
        cleaned = []
        for index, out in enumerate(outputs):
            n = specs[index]["shape"]["width"]
            crt = ((specs[index].get("encryption") or {}).get("encoding") or {}).get("crt") or []
            if self.specs.output_signs[index] and crt:            # client.py:217-235: values live mod prod(crt)
                M = int(np.prod([int(m) for m in crt], dtype=object))
                if isinstance(out, int):
                    cleaned.append(out if out < M // 2 else out - M)
                else:
                    o = out.astype(np.int64)
                    cleaned.append(np.where(o < M // 2, o, o - M).astype(np.int64))
            elif self.specs.output_signs[index]:                  # client.py:237-250
                if isinstance(out, int):
                    out %= 1 << n
                    cleaned.append(out if out < (1 << (n - 1)) else out - (1 << n))
                else:
                    o = out.astype(np.int64) % (1 << n)
                    cleaned.append(np.where(o < (1 << (n - 1)), o, o - (1 << n)).astype(np.int64))
            else:
                cleaned.append(out if isinstance(out, int) else out.astype(np.uint64))
        return cleaned[0] if len(cleaned) == 1 else tuple(cleaned)

    @property
    def evaluation_keys(self) -> cc.EvaluationKeys:
        self.keygen(force=False)
        return self._keyset.get_evaluation_keys()


def _value_of(arg, encrypted: bool) -> str:
    if isinstance(arg, (bool, np.bool_)):
        return _describe("uint1", (), encrypted)
    if isinstance(arg, (int, np.integer)):
        return _describe(_int_dtype_name(int(arg), int(arg)), (), encrypted)
    if isinstance(arg, (float, np.floating)):
        return _describe("float64", (), encrypted)
    if isinstance(arg, np.ndarray) and np.issubdtype(arg.dtype, np.integer):
        lo, hi = (int(arg.min()), int(arg.max())) if arg.size else (0, 0)
        return _describe(_int_dtype_name(lo, hi), arg.shape, encrypted)
    if isinstance(arg, np.ndarray) and np.issubdtype(arg.dtype, np.floating):
        return _describe(f"float{arg.dtype.itemsize * 8}", arg.shape, encrypted)
    raise ValueError(f"Value cannot represent {arg!r}")


class Server:
    def __init__(self, client_specs: ClientSpecs, output_dir, support, compilation_result, server_lambda):
        self.client_specs = client_specs
        self._output_dir = output_dir
        self._support = support
        self._compilation_result = compilation_result
        self._feedback = support.load_compilation_feedback(compilation_result)
        self._server_lambda = server_lambda
        self._mlir: Optional[str] = None
        self._configuration: Optional[Configuration] = None

    @staticmethod
    def create(mlir: str, input_signs: List[bool], output_signs: List[bool],
               configuration: Optional[Configuration] = None, crypto_parameters=None) -> "Server":
        cfg = configuration or Configuration()
        options = cc.CompilationOptions.new("main")
        options.set_loop_parallelize(cfg.loop_parallelize)
        options.set_dataflow_parallelize(cfg.dataflow_parallelize)
        options.set_auto_parallelize(cfg.auto_parallelize)
        # same precedence as reference server.py:102-126
        if cfg.global_p_error is not None and cfg.p_error is not None:
            options.set_global_p_error(cfg.global_p_error)
            options.set_p_error(cfg.p_error)
        elif cfg.global_p_error is not None:
            options.set_global_p_error(cfg.global_p_error)
        elif cfg.p_error is not None:
            options.set_p_error(cfg.p_error)
        else:
            options.set_global_p_error(DEFAULT_GLOBAL_P_ERROR)
        options.set_display_optimizer_choice(cfg.show_optimizer)
        if crypto_parameters is not None:
            options.set_crypto_parameters(crypto_parameters)
        output_dir = None
        if cfg.jit:
            support = cc.JITSupport.new()
        else:
            output_dir = tempfile.TemporaryDirectory()
            support = cc.LibrarySupport.new(output_dir.name, generateCppHeader=False, generateStaticLib=False)
        result = support.compile(mlir, options)
        server_lambda = support.load_server_lambda(result)
        specs = ClientSpecs(input_signs, support.load_client_parameters(result), output_signs)
        server = Server(specs, output_dir, support, result, server_lambda)
        server._mlir, server._configuration = mlir, cfg
        return server

    def save(self, path: Union[str, Path], via_mlir: bool = False):
        path = str(path)
        path = path[:-4] if path.endswith(".zip") else path
        if via_mlir:
            if self._mlir is None or self._configuration is None:
                raise RuntimeError("Loaded server objects cannot be saved again via MLIR")
            with tempfile.TemporaryDirectory() as tmp:
                (Path(tmp) / "circuit.mlir").write_text(self._mlir, encoding="utf-8")
                (Path(tmp) / "input_signs.json").write_text(json.dumps(self.client_specs.input_signs))
                (Path(tmp) / "output_signs.json").write_text(json.dumps(self.client_specs.output_signs))
                (Path(tmp) / "configuration.json").write_text(json.dumps(asdict(self._configuration)))
                shutil.make_archive(path, "zip", tmp)
            return
        if self._output_dir is None:
            raise RuntimeError("Just-in-Time compilation cannot be saved")
        (Path(self._output_dir.name) / "client.specs.json").write_text(self.client_specs.serialize())
        shutil.make_archive(path, "zip", self._output_dir.name)

    @staticmethod
    def load(path: Union[str, Path]) -> "Server":
        output_dir = tempfile.TemporaryDirectory()
        root = Path(output_dir.name)
        shutil.unpack_archive(str(path), str(root), "zip")
        if (root / "circuit.mlir").exists():
            cfg = Configuration(**json.loads((root / "configuration.json").read_text()))
            return Server.create((root / "circuit.mlir").read_text(encoding="utf-8"),
                                 json.loads((root / "input_signs.json").read_text()),
                                 json.loads((root / "output_signs.json").read_text()), cfg)
        specs = ClientSpecs.unserialize((root / "client.specs.json").read_text())
        support = cc.LibrarySupport.new(str(root), generateCppHeader=False, generateStaticLib=False)
        result = support.reload("main")
        return Server(specs, output_dir, support, result, support.load_server_lambda(result))

    def run(self, args: cc.PublicArguments, evaluation_keys: cc.EvaluationKeys) -> cc.PublicResult:
        return self._support.server_call(self._server_lambda, args, evaluation_keys)

    def cleanup(self):
        if self._output_dir is not None:
            self._output_dir.cleanup()

    complexity = property(lambda self: self._feedback.complexity)
    size_of_secret_keys = property(lambda self: self._feedback.total_secret_keys_size)
    size_of_bootstrap_keys = property(lambda self: self._feedback.total_bootstrap_keys_size)
    size_of_keyswitch_keys = property(lambda self: self._feedback.total_keyswitch_keys_size)
    size_of_inputs = property(lambda self: self._feedback.total_inputs_size)
    size_of_outputs = property(lambda self: self._feedback.total_output_size)
    p_error = property(lambda self: self._feedback.p_error)
    global_p_error = property(lambda self: self._feedback.global_p_error)


class Circuit:
    """`Circuit` of the reference minus the graph (which belongs to the frontend)."""

    def __init__(self, mlir: str, input_signs: List[bool], output_signs: List[bool],
                 configuration: Optional[Configuration] = None, crypto_parameters=None):
        self.mlir = mlir
        self.configuration = configuration or Configuration()
        self.server = Server.create(mlir, input_signs, output_signs, self.configuration, crypto_parameters)
        cache = None
        if self.configuration.use_insecure_key_cache:
            cache = self.configuration.insecure_key_cache_location or os.path.join(
                tempfile.gettempdir(), "cnp_b200_keycache")
        self.client = Client(self.server.client_specs, cache)

    def __str__(self):
        return self.mlir

    def keygen(self, force: bool = False):
        self.client.keygen(force)

    def encrypt(self, *args):
        return self.client.encrypt(*args)

    def run(self, args: cc.PublicArguments) -> cc.PublicResult:
        self.keygen(force=False)
        return self.server.run(args, self.client.evaluation_keys)

    def decrypt(self, result: cc.PublicResult):
        return self.client.decrypt(result)

    def encrypt_run_decrypt(self, *args):
        return self.decrypt(self.run(self.encrypt(*args)))

    def cleanup(self):
        self.server.cleanup()

    complexity = property(lambda self: self.server.complexity)
    size_of_secret_keys = property(lambda self: self.server.size_of_secret_keys)
    size_of_bootstrap_keys = property(lambda self: self.server.size_of_bootstrap_keys)
    size_of_keyswitch_keys = property(lambda self: self.server.size_of_keyswitch_keys)
    size_of_inputs = property(lambda self: self.server.size_of_inputs)
    size_of_outputs = property(lambda self: self.server.size_of_outputs)
    p_error = property(lambda self: self.server.p_error)
    global_p_error = property(lambda self: self.server.global_p_error)
