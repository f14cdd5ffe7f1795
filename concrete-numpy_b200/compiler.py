"""
`concrete.compiler`-compatible module: every name the reference imports from the absent
concrete-compiler wheel, implemented on the B200 engine.

Call sites mirrored (reference):
  concrete/numpy/compilation/server.py:93-166   CompilationOptions, JITSupport / LibrarySupport
  concrete/numpy/compilation/server.py:286      support.server_call(lambda, args, eval_keys)
  concrete/numpy/compilation/client.py:48,106,178-182,201,271   KeySetCache, ClientSupport, KeySet
  concrete/numpy/compilation/specs.py:39-127    ClientParameters / PublicArguments / PublicResult
  tests/compilation/test_circuit.py:174-177     EvaluationKeys.serialize / unserialize

Errors surface as RuntimeError, like the wheel's.
"""

import hashlib
import io
import json
import math
import os
import struct
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import engine as E
from . import params as P
from .engine import CryptoParams
from .executor import Ct, Executor, ProgramInfo, analyze
from .mlir_text.parser import Program, parse_module

_MAGIC = b"CNPB200\x01"


def default_device() -> torch.device:
    """Device of this process: LOCAL_RANK under torchrun, else CNP_B200_DEVICE, else cuda:0."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "no CUDA device: the B200 backend has no CPU fallback for keygen / encrypt / run / decrypt"
        )
    idx = int(os.environ.get("CNP_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return torch.device(f"cuda:{idx}")


def init_dfr() -> None:
    """The wheel starts its HPX dataflow runtime here (server.py:99-100); nothing to start."""


# ---------------------------------------------------------------- options ---------

class CompilationOptions:
    def __init__(self, function_name: str):
        self.function_name = function_name
        self.loop_parallelize = False
        self.dataflow_parallelize = False
        self.auto_parallelize = False
        self.p_error = 1.0
        self.global_p_error = 1.0
        self.display_optimizer_choice = False
        self.parameters_override: Optional[CryptoParams] = None

    @staticmethod
    def new(function_name: str = "main") -> "CompilationOptions":
        return CompilationOptions(function_name)

    def set_loop_parallelize(self, v: bool):
        self.loop_parallelize = bool(v)

    def set_dataflow_parallelize(self, v: bool):
        self.dataflow_parallelize = bool(v)

    def set_auto_parallelize(self, v: bool):
        self.auto_parallelize = bool(v)

    def set_p_error(self, v: float):
        if not 0.0 < v <= 1.0:
            raise RuntimeError("p_error should be a probability in ]0; 1]")
        self.p_error = float(v)

    def set_global_p_error(self, v: float):
        if not 0.0 < v <= 1.0:
            raise RuntimeError("global_p_error should be a probability in ]0; 1]")
        self.global_p_error = float(v)

    def set_display_optimizer_choice(self, v: bool):
        self.display_optimizer_choice = bool(v)

    # extension (not in the wheel): pin a parameter set instead of searching one
    def set_crypto_parameters(self, params: CryptoParams):
        self.parameters_override = params


# ---------------------------------------------------------------- parameters ------

class ClientParameters:
    """JSON document read by the reference's client (client.py:121-141, 207-215)."""

    def __init__(self, doc: dict):
        self.doc = doc

    def serialize(self) -> str:
        return json.dumps(self.doc, sort_keys=True)

    @staticmethod
    def unserialize(data: Union[str, bytes]) -> "ClientParameters":
        if isinstance(data, (bytes, bytearray)):
            data = bytes(data).decode("utf-8")
        return ClientParameters(json.loads(data))

    @property
    def crypto(self) -> CryptoParams:
        return CryptoParams(**self.doc["crypto"])

    def input_specs(self) -> List[dict]:
        return self.doc["inputs"]

    def output_specs(self) -> List[dict]:
        return self.doc["outputs"]

    def cache_key(self) -> str:
        return hashlib.sha256(json.dumps(self.doc["crypto"], sort_keys=True).encode()).hexdigest()[:32]


def _gate(t, crypto: CryptoParams) -> dict:
    enc = None
    if t.encrypted:
        enc = {"secretKeyID": "big", "variance": crypto.glwe_std ** 2,
               "encoding": {"precision": t.width, "crt": [int(m) for m in crypto.moduli]}}
    return {"shape": {"width": t.width, "dimensions": list(t.dims), "size": t.size, "sign": False},
            "encryption": enc}


class CompilationFeedback:
    """Attributes read by Server (server.py:296-350)."""

    def __init__(self, info: ProgramInfo, crypto: CryptoParams, program: Program, p_error: float):
        L = crypto.ct_words * max(len(crypto.moduli), 1)          # words per encrypted element (all residues)
        self.total_secret_keys_size = int(8 * (crypto.n + crypto.big_dim))
        self.total_bootstrap_keys_size = int(8 * crypto.n * crypto.pbs_level * (crypto.k + 1) ** 2 * crypto.N)
        self.total_keyswitch_keys_size = int(8 * crypto.big_dim * crypto.ks_level * (crypto.n + 1))
        if crypto.linear_only:
            self.complexity = float(sum(t.size for t in program.result_types) * crypto.ct_words)
        elif crypto.wide:
            from . import wide

            self.complexity = float(info.n_pbs * P.wide_lookup_flops(crypto))
            self.total_keyswitch_keys_size += int(8 * (crypto.big_dim + wide.PF_PAD) * crypto.pf_level
                                                  * (crypto.k + 1) ** 2 * crypto.N)
        else:
            self.complexity = float(P.complexity(crypto, info.n_pbs))
        self.total_inputs_size = int(sum(8 * L * t.size if t.encrypted else 8 * t.size for _, t in program.args))
        self.total_output_size = int(sum(8 * L * t.size if t.encrypted else 8 * t.size for t in program.result_types))
        self.p_error = float(p_error)
        self.global_p_error = float(1.0 - (1.0 - min(p_error, 1.0)) ** max(info.n_pbs, 1))
        self.n_pbs = info.n_pbs
        self.tlu_depth = info.tlu_depth
        self.norm2 = info.norm2


class CompilationResult:
    def __init__(self, mlir: str, program: Program, info: ProgramInfo, crypto: CryptoParams,
                 client_parameters: ClientParameters, feedback: CompilationFeedback):
        self.mlir = mlir
        self.program = program
        self.info = info
        self.crypto = crypto
        self.client_parameters = client_parameters
        self.feedback = feedback


class JITCompilationResult(CompilationResult):
    pass


class LibraryCompilationResult(CompilationResult):
    pass


_SHARD_GROUP = None


def set_shard_group(group) -> None:
    """Shard every table lookup of subsequent `server_call`s by ciphertext over the ranks of `group`
    (a torch.distributed process group, e.g. `dist.group.WORLD`; None switches sharding off).  All ranks
    must then call `Server.run` with the same arguments; keys are replicated (deterministic keygen)."""
    global _SHARD_GROUP
    _SHARD_GROUP = group


def _shard_group():
    if _SHARD_GROUP is not None:
        return _SHARD_GROUP
    if os.environ.get("CNP_B200_SHARD", "0") not in ("", "0"):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.group.WORLD
    return None


class ServerLambda:
    """Executable form of a compiled program: one Executor per (device, shard group), created lazily."""

    def __init__(self, result: CompilationResult):
        self.result = result
        self._executors: Dict[tuple, Executor] = {}

    def executor(self, device) -> Executor:
        group = _shard_group()
        key = (str(device), group)          # the group object itself: an id() can be reused after collection
        if key not in self._executors:
            self._executors[key] = Executor(self.result.program, self.result.crypto, device, group=group)
        return self._executors[key]


class JITLambda(ServerLambda):
    pass


class LibraryLambda(ServerLambda):
    pass


def _compile(mlir: str, options: CompilationOptions, cls):
    program = parse_module(mlir)
    info = analyze(program)
    wide_mode = info.precision > 8          # beyond the native bootstrap: CRT residues + without-padding lookup (wide.py)
    if options.parameters_override is not None:
        crypto = options.parameters_override
        if crypto.p != info.precision:
            raise RuntimeError(f"pinned parameters carry {crypto.p} bits, circuit needs {info.precision}")
        if crypto.wide:
            info = analyze(program, crypto.moduli)
        elif wide_mode and info.n_pbs > 0:
            raise RuntimeError(f"{info.precision}-bit circuits need CRT parameters (moduli)")
    else:
        n = max(info.n_pbs, 1)
        per_tlu = 1.0 - (1.0 - options.global_p_error) ** (1.0 / n) if options.global_p_error < 1.0 else 1.0
        target = min(options.p_error, per_tlu)
        if target >= 1.0:
            target = 1.0 - (1.0 - 1e-5) ** (1.0 / n)     # the reference's DEFAULT_GLOBAL_P_ERROR
        # canonical buckets (norm2 up to a power of two, p_error down to a power of ten): fewer
        # distinct parameter sets, always on the conservative side
        norm2_q = float(2 ** math.ceil(math.log2(max(info.norm2, 1.0))))
        target_q = float(10.0 ** math.floor(math.log10(target) + 1e-9))
        try:
            if info.n_pbs == 0:
                # no table lookup: any width that fits a torus word, no evaluation keys (graph_converter.py:288-306)
                crypto = P.select_linear_parameters(info.precision, norm2_q, target_q)
            elif wide_mode:
                from . import wide

                if info.precision > 16:
                    raise ValueError(f"{info.precision}-bit integers exceed the 16-bit limit of table lookups")
                info = analyze(program, wide.crt_moduli(info.precision))
                norm2_q = float(2 ** math.ceil(math.log2(max(info.norm2, 1.0))))
                crypto = P.select_wide_parameters(info.precision, norm2_q, target_q)
            else:
                crypto = P.select_parameters(info.precision, norm2_q, target_q)
        except ValueError as error:
            raise RuntimeError(f"Unfeasible parameter search: {error}") from error
    if crypto.wide:
        p_err = P.estimate_wide_p_error(crypto, info.norm2)
    elif crypto.linear_only:
        if info.n_pbs:
            raise RuntimeError("parameters without bootstrap keys cannot run a circuit with table lookups")
        p_err = P.p_error_of(crypto.p, max(info.norm2, 1.0) * crypto.glwe_std ** 2)
    else:
        p_err = P.estimate_p_error(crypto, info.norm2)
    if options.display_optimizer_choice:
        print(f"--- Circuit\n  {info.precision} bits integers\n  {info.norm2:g} manp (squared 2-norm)\n"
              f"  {info.n_pbs} table lookups, depth {info.tlu_depth}\n--- Optimizer choice\n  {crypto}\n"
              f"  p_error per lookup {p_err:.3g}")
    doc = {
        "functionName": program.name,
        "inputs": [_gate(t, crypto) for _, t in program.args],
        "outputs": [_gate(t, crypto) for t in program.result_types],
        "crypto": crypto.to_dict(),
    }
    cp = ClientParameters(doc)
    return cls(mlir, program, info, crypto, cp, CompilationFeedback(info, crypto, program, p_err))


class JITSupport:
    @staticmethod
    def new(runtime_lib_path: Optional[str] = None) -> "JITSupport":
        return JITSupport()

    def compile(self, mlir_program: str, options: Optional[CompilationOptions] = None) -> JITCompilationResult:
        return _compile(mlir_program, options or CompilationOptions.new("main"), JITCompilationResult)

    def load_server_lambda(self, result: CompilationResult) -> JITLambda:
        return JITLambda(result)

    def load_client_parameters(self, result: CompilationResult) -> ClientParameters:
        return result.client_parameters

    def load_compilation_feedback(self, result: CompilationResult) -> CompilationFeedback:
        return result.feedback

    def server_call(self, server_lambda: ServerLambda, public_arguments: "PublicArguments",
                    evaluation_keys: "EvaluationKeys") -> "PublicResult":
        return _server_call(server_lambda, public_arguments, evaluation_keys)


class LibrarySupport(JITSupport):
    """Ahead-of-time flavour: the 'library' on disk is the MLIR text + the chosen parameters."""

    def __init__(self, output_path: str):
        self.output_path = output_path

    @staticmethod
    def new(output_path: str = "./out", generateSharedLib: bool = True, generateStaticLib: bool = False,
            generateClientParameters: bool = True, generateCompilationFeedback: bool = True,
            generateCppHeader: bool = False) -> "LibrarySupport":
        os.makedirs(output_path, exist_ok=True)
        return LibrarySupport(output_path)

    def compile(self, mlir_program: str, options: Optional[CompilationOptions] = None) -> LibraryCompilationResult:
        res = _compile(mlir_program, options or CompilationOptions.new("main"), LibraryCompilationResult)
        with open(os.path.join(self.output_path, "program.mlir"), "w", encoding="utf-8") as f:
            f.write(mlir_program)
        with open(os.path.join(self.output_path, "client_parameters.concrete.params.json"), "w", encoding="utf-8") as f:
            f.write(res.client_parameters.serialize())
        return res

    def reload(self, func_name: str = "main") -> LibraryCompilationResult:
        with open(os.path.join(self.output_path, "program.mlir"), "r", encoding="utf-8") as f:
            mlir = f.read()
        with open(os.path.join(self.output_path, "client_parameters.concrete.params.json"), "r", encoding="utf-8") as f:
            cp = ClientParameters.unserialize(f.read())
        options = CompilationOptions.new(func_name)
        options.set_crypto_parameters(cp.crypto)
        return _compile(mlir, options, LibraryCompilationResult)

    def load_server_lambda(self, result: CompilationResult) -> LibraryLambda:
        return LibraryLambda(result)


# ---------------------------------------------------------------- keys ------------

class KeySetCache:
    """Insecure key cache (client.py:47-48): keys are a pure function of (key material, parameters), so the
    cache stores the 64 bytes of key material under a digest of the crypto parameters, in a file only its owner
    can read.  Like the reference's cache it re-uses KEYS; encryption randomness is fresh on every call."""

    def __init__(self, path: str):
        self.path = path

    @staticmethod
    def new(path: str) -> "KeySetCache":
        os.makedirs(path, exist_ok=True)
        return KeySetCache(path)

    def seed_for(self, client_parameters: ClientParameters) -> bytes:
        fn = os.path.join(self.path, client_parameters.cache_key() + ".key")
        if os.path.exists(fn):
            with open(fn, "rb") as f:
                data = f.read()
            if len(data) == 64:
                return data
        seed = E.fresh_key()
        fd = os.open(fn, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
        with os.fdopen(fd, "wb") as f:
            f.write(seed)
        os.chmod(fn, 0o600)
        return seed


class EvaluationKeys:
    """BSK (Fourier) + KSK, device resident on the server; bytes on the wire."""

    def __init__(self, crypto: CryptoParams, bsk_f: torch.Tensor, ksk: torch.Tensor,
                 pfksk: Optional[torch.Tensor] = None):
        self.crypto = crypto
        self.bsk_f = bsk_f
        self.ksk = ksk
        self.pfksk = pfksk          # packing keyswitch key of the wide-integer lookup (None for native widths)
        self._per_device: Dict[str, E.EvalKeys] = {}

    def on(self, device) -> E.EvalKeys:
        key = str(torch.device(device))
        if key not in self._per_device:
            self._per_device[key] = E.EvalKeys(self.crypto, self.bsk_f.to(device), self.ksk.to(device),
                                               None if self.pfksk is None else self.pfksk.to(device))
        return self._per_device[key]

    def serialize(self) -> bytes:
        header = json.dumps({"crypto": self.crypto.to_dict(), "bsk_shape": list(self.bsk_f.shape),
                             "ksk_shape": list(self.ksk.shape),
                             "pfksk_shape": None if self.pfksk is None else list(self.pfksk.shape)}).encode()
        buf = io.BytesIO()
        buf.write(_MAGIC + b"EVK")
        buf.write(struct.pack("<I", len(header)))
        buf.write(header)
        buf.write(self.bsk_f.detach().cpu().numpy().tobytes())
        buf.write(self.ksk.detach().cpu().numpy().tobytes())
        if self.pfksk is not None:
            buf.write(self.pfksk.detach().cpu().numpy().tobytes())
        return buf.getvalue()

    @staticmethod
    def unserialize(data: bytes, device=None) -> "EvaluationKeys":
        if data[:len(_MAGIC) + 3] != _MAGIC + b"EVK":
            raise RuntimeError("not a serialized EvaluationKeys blob of this backend")
        off = len(_MAGIC) + 3
        (hl,) = struct.unpack_from("<I", data, off)
        off += 4
        h = json.loads(data[off:off + hl].decode())
        off += hl
        crypto = CryptoParams(**h["crypto"])
        nb = int(np.prod(h["bsk_shape"])) * 8
        bsk = np.frombuffer(data, dtype=np.float64, count=nb // 8, offset=off).reshape(h["bsk_shape"])
        off += nb
        ksk = np.frombuffer(data, dtype=np.int64, count=int(np.prod(h["ksk_shape"])), offset=off).reshape(h["ksk_shape"])
        off += ksk.size * 8
        dev = default_device() if device is None else device
        pfksk = None
        if h.get("pfksk_shape"):
            pf = np.frombuffer(data, dtype=np.int64, count=int(np.prod(h["pfksk_shape"])), offset=off).reshape(h["pfksk_shape"])
            pfksk = torch.from_numpy(pf.copy()).to(dev)
        return EvaluationKeys(crypto, torch.from_numpy(bsk.copy()).to(dev), torch.from_numpy(ksk.copy()).to(dev), pfksk)


_KEYSET_MEMO: Dict[tuple, E.KeySet] = {}


class KeySet:
    def __init__(self, keys: E.KeySet):
        self.keys = keys
        self._enc_counter = 0

    def get_evaluation_keys(self) -> EvaluationKeys:
        return EvaluationKeys(self.keys.params, self.keys.bsk_f, self.keys.ksk, self.keys.pfksk)


# ---------------------------------------------------------- arguments / results ---

def _pack(kind: bytes, entries: List[dict], blobs: List[bytes]) -> bytes:
    header = json.dumps(entries).encode()
    return _MAGIC + kind + struct.pack("<I", len(header)) + header + b"".join(blobs)


def _unpack(kind: bytes, data: bytes):
    if data[:len(_MAGIC) + 3] != _MAGIC + kind:
        raise RuntimeError(f"not a serialized {kind.decode()} blob of this backend")
    off = len(_MAGIC) + 3
    (hl,) = struct.unpack_from("<I", data, off)
    off += 4
    entries = json.loads(data[off:off + hl].decode())
    return entries, off + hl


class _ValueList:
    KIND = b"???"

    def __init__(self, values: List[Union[Ct, np.ndarray]]):
        self.values = values

    def serialize(self) -> bytes:
        entries, blobs = [], []
        for v in self.values:
            if isinstance(v, Ct):
                raw = v.data.detach().cpu().numpy().tobytes()
                entries.append({"enc": True, "shape": list(v.shape), "rows": int(v.data.shape[0]),
                                "words": int(v.data.shape[1]), "bytes": len(raw)})
            else:
                a = np.ascontiguousarray(np.asarray(v, dtype=np.int64))
                raw = a.tobytes()
                entries.append({"enc": False, "shape": list(a.shape), "bytes": len(raw)})
            blobs.append(raw)
        return _pack(self.KIND, entries, blobs)

    @classmethod
    def unserialize(cls, client_parameters: ClientParameters, data: bytes, device=None):
        entries, off = _unpack(cls.KIND, data)
        dev = None
        values = []
        for e in entries:
            a = np.frombuffer(data, dtype=np.int64, count=e["bytes"] // 8, offset=off)
            off += e["bytes"]
            if e["enc"]:
                if dev is None:
                    dev = default_device() if device is None else device
                t = torch.from_numpy(a.reshape(e["rows"], e["words"]).copy()).to(dev)
                values.append(Ct(t, tuple(e["shape"])))
            else:
                values.append(a.reshape(e["shape"]).copy())
        return cls(values)


class PublicArguments(_ValueList):
    KIND = b"ARG"


class PublicResult(_ValueList):
    KIND = b"RES"


# ---------------------------------------------------------------- client ----------

class ClientSupport:
    @staticmethod
    def key_set(client_parameters: ClientParameters, keyset_cache: Optional[KeySetCache] = None,
                seed: Optional[int] = None, device=None) -> KeySet:
        dev = torch.device(default_device() if device is None else device)
        memo = False
        if seed is None:
            env_seed = os.environ.get("CNP_B200_KEY_SEED")
            if env_seed is not None:
                seed, memo = int(env_seed), True       # deterministic test mode: reuse device keys
            elif keyset_cache is not None:
                seed = keyset_cache.seed_for(client_parameters)
            else:
                seed = E.fresh_key()                   # 256-bit secret key + 256-bit public-mask key from the OS
        if not memo:
            return KeySet(E.KeySet(client_parameters.crypto, dev, seed))
        key = (client_parameters.crypto, seed, str(dev))
        if key not in _KEYSET_MEMO:
            if len(_KEYSET_MEMO) >= 6:                 # bound HBM held by memoised key sets
                _KEYSET_MEMO.pop(next(iter(_KEYSET_MEMO)))
            _KEYSET_MEMO[key] = E.KeySet(client_parameters.crypto, dev, seed)
        return KeySet(_KEYSET_MEMO[key])

    @staticmethod
    def encrypt_arguments(client_parameters: ClientParameters, keyset: KeySet,
                          args: Sequence[Union[int, np.ndarray]]) -> PublicArguments:
        specs = client_parameters.input_specs()
        if len(args) != len(specs):
            raise RuntimeError(f"function has arity {len(specs)} but is applied to {len(args)} arguments")
        crypto = keyset.keys.params
        values: List[Union[Ct, np.ndarray]] = []
        for arg, spec in zip(args, specs):
            shape = tuple(spec["shape"]["dimensions"])
            a = np.asarray(arg)
            if a.shape != shape:
                raise RuntimeError(f"argument shape {a.shape} does not match the expected {shape}")
            if spec["encryption"] is None:
                values.append(a.astype(np.int64))
                continue
            width = spec["shape"]["width"]
            m = a.astype(np.uint64).reshape(-1)
            if m.size and int(m.max()) >= (1 << width):
                raise RuntimeError(f"argument value does not fit {width} bits")
            # mask / noise streams: fresh OS entropy for every call; only a key set made from a TEST seed (explicit
            # `seed=` / CNP_B200_KEY_SEED: bit-reproducible runs, identical ciphertexts on every rank) derives them
            # from the seed and a running index
            det = keyset.keys.deterministic
            enc_seed = E.derive_key(keyset.keys.seed, 0x5EEDC0DEC0FFEE) if det else None
            if crypto.wide:
                # one ciphertext per CRT residue, rows residue-major: [residue][element]
                from . import wide

                pts = np.concatenate([wide.encode_residue(m.astype(np.int64), mod) for mod in crypto.moduli])
                ct = E.encrypt_plaintexts(keyset.keys, pts, first_index=keyset._enc_counter if det else 0, seed=enc_seed)
                keyset._enc_counter += int(pts.size)
            else:
                ct = E.encrypt(keyset.keys, m, first_index=keyset._enc_counter if det else 0, seed=enc_seed)
                keyset._enc_counter += int(m.size)
            values.append(Ct(ct, shape))
        return PublicArguments(values)

    @staticmethod
    def decrypt_result(client_parameters: ClientParameters, keyset: KeySet, public_result: PublicResult):
        specs = client_parameters.output_specs()
        outs = []
        for v, spec in zip(public_result.values, specs):
            shape = tuple(spec["shape"]["dimensions"])
            crypto = keyset.keys.params
            if isinstance(v, Ct) and crypto.wide:
                # CRT: decode every residue, recombine; the value lives mod prod(moduli) (client.py:217-235)
                from . import wide

                _, phase = E.decrypt(keyset.keys, v.data.to(keyset.keys.device), return_phase=True)
                per = phase.reshape(len(crypto.moduli), -1)
                m = wide.crt_combine([wide.decode_residue(per[i], mod) for i, mod in enumerate(crypto.moduli)],
                                     crypto.moduli).astype(np.uint64)
            elif isinstance(v, Ct):
                m = E.decrypt(keyset.keys, v.data.to(keyset.keys.device))
                width = spec["shape"]["width"]
                m = m & np.uint64((1 << (width + 1)) - 1)
            else:
                m = np.asarray(v).astype(np.uint64).reshape(-1)
            outs.append(int(m[0]) if shape == () else m.reshape(shape))
        return outs[0] if len(outs) == 1 else tuple(outs)


def _server_call(server_lambda: ServerLambda, public_arguments: PublicArguments,
                 evaluation_keys: EvaluationKeys) -> PublicResult:
    result = server_lambda.result
    if evaluation_keys.crypto != result.crypto:
        raise RuntimeError("evaluation keys were generated for different crypto parameters")
    enc = [v for v in public_arguments.values if isinstance(v, Ct)]
    device = enc[0].data.device if enc else evaluation_keys.bsk_f.device
    ex = server_lambda.executor(device)
    with _device_context(device):
        outs = ex.run(evaluation_keys.on(device), list(public_arguments.values))
    return PublicResult(outs)


def _device_context(device):
    """torch.cuda.device(...) for a CUDA device; a no-op for the CPU tensors of the host-logic tests (which swap the
    engine for tests/clear_engine.py -- the product engine itself raises without CUDA)."""
    import contextlib

    dev = torch.device(device)
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()
