"""
Executor: walks the FHE / FHELinalg program parsed from the reference's MLIR text and launches
the CUDA kernels through the C ABI (engine.py).  This is the body of `Server.run`
(reference: concrete/numpy/compilation/server.py:270-286 -> `server_call`).

Value model
  * clear value      : numpy int64 array (shape from the MLIR type)
  * encrypted value  : `Ct(data, shape)` with `data` a torch.int64 `[E, k*N+1]` device tensor,
                       E = prod(shape), elements in row-major order

Op semantics follow the reference's lowering, cited per handler
(concrete/numpy/mlir/node_converter.py).  Linear / data-movement ops with numpy broadcasting are
resolved ON THE HOST into int32 row maps once per program ("plans", cached), so each op is one
kernel launch on device.
"""

import os
import threading
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import engine as E
from .engine import CryptoParams, EvalKeys
from .mlir_text.parser import Op, Program, TType


@dataclass
class Ct:
    data: torch.Tensor            # [E, L] int64; when `sharded`: this rank's block of rows only (Executor._part)
    shape: Tuple[int, ...]
    sharded: bool = False


def _numel(shape) -> int:
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


# ------------------------------------------------------------------- analysis -----

@dataclass
class ProgramInfo:
    precision: int        # circuit-wide eint width p
    n_pbs: int            # table lookups per run (keyswitch + bootstrap each)
    tlu_depth: int
    norm2: float          # largest squared 2-norm of the clear linear maps feeding a lookup / output
    n_inputs_encrypted: int


TLU_OPS = {"FHE.apply_lookup_table", "FHELinalg.apply_lookup_table", "FHELinalg.apply_mapped_lookup_table"}


def _centered(c: np.ndarray, m: int) -> np.ndarray:
    """Representative of c mod m in [-m/2, m/2): what a residue ciphertext is actually multiplied by."""
    c = np.asarray(c, dtype=np.int64)
    return np.mod(c + m // 2, m) - m // 2


def analyze(program: Program, moduli: Tuple[int, ...] = ()) -> ProgramInfo:
    """Static pass: precision, PBS count / depth and the noise growth of the linear parts.  With CRT `moduli`
    (wide integers) every clear weight acts modulo each modulus, so the norms are taken over the centred residues
    of the weights (worst modulus)."""
    widths = {t.width for _, t in program.args if t.encrypted}
    for op in program.ops:
        if op.result_type is not None and op.result_type.encrypted:
            widths.add(op.result_type.width)
    if not widths:
        raise RuntimeError("program has no encrypted value")
    precision = max(widths)
    nu: Dict[str, float] = {}      # variance multiplier (in units of a fresh/bootstrapped ct)
    depth: Dict[str, int] = {}
    const: Dict[str, np.ndarray] = {}
    clear_bound: Dict[str, float] = {}
    for name, t in program.args:
        if t.encrypted:
            nu[name], depth[name] = 1.0, 0
        else:
            clear_bound[name] = float(2 ** t.width)
    n_pbs, worst = 0, 1.0

    def reductions(c: np.ndarray) -> List[np.ndarray]:
        return [_centered(c, m) for m in moduli] if moduli else [np.asarray(c)]

    def stat(c: np.ndarray, fn) -> float:
        return max(float(fn(r.astype(np.float64))) for r in reductions(c))

    def cmax(name: str) -> float:
        if name in const:
            c = const[name]
            return stat(c, lambda r: np.max(np.abs(r))) if c.size else 0.0
        bound = clear_bound.get(name, 1.0)
        return min(bound, max(moduli) / 2.0) if moduli else bound

    for op in program.ops:
        r, o = op.result, op.operands
        enc_in = [x for x in o if x in nu]
        d = max([depth[x] for x in enc_in], default=0)
        if op.name == "arith.constant":
            const[r] = np.asarray(op.attrs["value"], dtype=np.int64)
            continue
        if op.name in TLU_OPS:
            n_pbs += op.result_type.size
            worst = max(worst, nu[o[0]])
            nu[r], depth[r] = 1.0, d + 1
            continue
        base = op.name.split(".", 1)[1]
        if base in ("add_eint", "sub_eint"):
            v = nu[o[0]] + nu[o[1]]
        elif base in ("add_eint_int", "sub_eint_int", "neg_eint"):
            v = nu[o[0]]
        elif base == "sub_int_eint":
            v = nu[o[1]]
        elif base == "mul_eint_int":
            v = nu[o[0]] * cmax(o[1]) ** 2
        elif base in ("zero", "zero_tensor"):
            v = 0.0
        elif base == "dot_eint_int":
            c = const.get(o[1])
            v = nu[o[0]] * (stat(c, lambda r: np.sum(r ** 2)) if c is not None
                            else cmax(o[1]) ** 2 * op.operand_types[0].size)
        elif base in ("matmul_eint_int", "matmul_int_eint"):
            ct_i, cl_i = (0, 1) if base == "matmul_eint_int" else (1, 0)
            c = const.get(o[cl_i])
            axis = -2 if base == "matmul_eint_int" else -1
            if c is not None:
                s = stat(c, lambda r: np.max((r ** 2).sum(axis=axis if r.ndim > 1 else 0))) if c.size else 0.0
            else:
                kdim = op.operand_types[cl_i].dims[axis if len(op.operand_types[cl_i].dims) > 1 else 0]
                s = cmax(o[cl_i]) ** 2 * kdim
            v = nu[o[ct_i]] * s
        elif base == "conv2d":
            c = const.get(o[1])
            if c is not None:
                s = stat(c, lambda r: np.max((r ** 2).reshape(r.shape[0], -1).sum(axis=1)))
            else:
                wt = op.operand_types[1]
                s = cmax(o[1]) ** 2 * (wt.size // wt.dims[0])
            v = nu[o[0]] * s
        elif base == "sum":
            v = nu[o[0]] * max(op.operand_types[0].size // max(op.result_type.size, 1), 1)
        elif base in ("concat", "from_elements", "insert_slice"):
            v = max([nu[x] for x in enc_in], default=0.0)
        else:   # transpose, tensor.* views
            v = nu[enc_in[0]] if enc_in else 0.0
        if op.result_type is not None and op.result_type.encrypted:
            nu[r], depth[r] = v, d
        elif r is not None and op.name.startswith("tensor."):
            # clear tensors moved around: keep constants foldable for later bounds
            clear_bound[r] = max([cmax(x) for x in o], default=1.0)
    for name in program.returns:
        if name in nu:
            worst = max(worst, nu[name])
    tlu_depth = max(depth.values(), default=0)
    return ProgramInfo(precision, n_pbs, tlu_depth, worst, sum(1 for _, t in program.args if t.encrypted))


# ------------------------------------------------------------------ helpers -------

def _row_map(src_shape, dst_shape) -> Optional[np.ndarray]:
    """int32 map dst element -> src element for numpy broadcasting (None = identity)."""
    if tuple(src_shape) == tuple(dst_shape):
        return None
    n = int(np.prod(src_shape, dtype=np.int64)) if len(src_shape) else 1
    idx = np.arange(n, dtype=np.int32).reshape(src_shape)
    return np.ascontiguousarray(np.broadcast_to(idx, dst_shape)).reshape(-1)


def _static_slices(offsets, sizes, strides):
    out = []
    for o, s, t in zip(offsets, sizes, strides):
        stop = o + s * t
        out.append(slice(o, stop if stop >= 0 else None, t))
    return tuple(out)


class Executor:
    """Runs one parsed program on one device with one parameter set."""

    def __init__(self, program: Program, params: CryptoParams, device, group=None):
        """`group`: optional torch.distributed process group.  With world size G > 1 every table lookup on at least
        `shard_min` ciphertexts is sharded BY CIPHERTEXT (keys are replicated per GPU): rank r bootstraps one
        contiguous block of rows and the result STAYS sharded.  Element-wise ops, further lookups and matmuls whose
        rows fall inside the blocks consume the local block with no communication; a sum over the sharded axis adds
        local partial sums with an all-reduce (the u64 wrap-around add is the torus add); an op that reads other
        elements (conv2d, index / concat / transpose, returning the result) all-gathers the value once (NCCL over
        NVLink on GPUs; cached on the value).  Values that never went through a sharded lookup -- the arguments, the
        cheap HBM-bound linear ops on them, and small tensors such as every step of the sequential key-value insert
        chain -- are replicated: every rank computes them and nothing is sent."""
        self.program = program
        self.params = params
        self.device = torch.device(device)
        self.group = group
        self.world, self.rank = 1, 0
        if group is not None:
            import torch.distributed as dist

            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.L = params.ct_words
        # wide integers: every encrypted value is a list of CRT residues; the linear ops run once per residue
        # ("lane", self._res) with that residue's plaintext encoding and weights reduced mod m, lookups see all lanes
        self.moduli = tuple(params.moduli)
        self.lanes = max(len(self.moduli), 1)
        self._res = 0
        # tensors below this many ciphertexts are replicated, not sharded (CRT circuits shard inside the lookup only)
        self.shard_min = (1 << 62) if self.moduli else int(os.environ.get("CNP_B200_SHARD_MIN", "64"))
        self._run_lock = threading.Lock()
        self.time_comm = False                      # bench.py: CUDA events around every collective
        self._comm_events: List[Tuple] = []
        self._comm = {"allgather_calls": 0, "allgather_bytes": 0, "allreduce_calls": 0, "allreduce_bytes": 0}
        self._plans: Dict[int, dict] = {}          # per-op cached host plans + device uploads
        self._consts: Dict[str, np.ndarray] = {}
        self._no_clear_args = all(t.encrypted for _, t in program.args)
        for op in program.ops:
            if op.name == "arith.constant":
                self._consts[op.result] = np.asarray(op.attrs["value"], dtype=np.int64).reshape(op.result_type.dims)
        self.handlers: Dict[str, Callable] = {
            "add_eint": self._add_eint, "sub_eint": self._sub_eint,
            "add_eint_int": self._add_eint_int, "sub_eint_int": self._sub_eint_int,
            "sub_int_eint": self._sub_int_eint, "neg_eint": self._neg_eint,
            "mul_eint_int": self._mul_eint_int,
            "apply_lookup_table": self._tlu, "apply_mapped_lookup_table": self._tlu,
            "zero": self._zero, "zero_tensor": self._zero,
            "dot_eint_int": self._dot, "matmul_eint_int": self._matmul, "matmul_int_eint": self._matmul,
            "conv2d": self._conv2d, "sum": self._sum, "concat": self._concat, "transpose": self._transpose,
            "extract": self._extract, "extract_slice": self._extract_slice, "insert_slice": self._insert_slice,
            "collapse_shape": self._reshape, "expand_shape": self._reshape, "from_elements": self._from_elements,
        }
        self.last_stats = {}

    # -- sharding by ciphertext (world > 1): SURVEY 8e ------------------------------------------------------
    def _part(self, n: int, rank: Optional[int] = None) -> Tuple[int, int, int]:
        """Block partition of n rows: (rows per block, first row, end row) of `rank` (default: this rank).  Blocks
        start on even rows: rows are an odd number of words long, so an odd start would break the 16-byte alignment
        the kernels vectorise on when a block is a view of a replicated tensor."""
        r = self.rank if rank is None else rank
        per = -(-n // self.world)
        per += per & 1
        lo = min(r * per, n)
        return per, lo, min(lo + per, n)

    def _shardable(self, n: int) -> bool:
        return self.world > 1 and n >= self.shard_min

    def _collective(self, fn, kind: str, nbytes: int):
        self._comm[kind + "_calls"] += 1
        self._comm[kind + "_bytes"] += int(nbytes)
        if self.time_comm and self.device.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            self._comm_events.append((e0, e1))
        else:
            fn()

    def _full(self, x: Ct) -> torch.Tensor:
        """All rows of a value on this rank (all-gather of the blocks when it is sharded; cached on the value)."""
        if not x.sharded:
            return x.data
        g = getattr(x, "_gathered", None)
        if g is None:
            import torch.distributed as dist

            per = x.data.shape[0]
            full = torch.empty((per * self.world, self.L), dtype=torch.int64, device=x.data.device)
            self._collective(lambda: dist.all_gather_into_tensor(full, x.data, group=self.group), "allgather",
                             full.numel() * 8)
            g = full[: _numel(x.shape)]
            x._gathered = g
        return g

    def _mine(self, x: Ct, lo: int, hi: int) -> torch.Tensor:
        """Rows [lo, hi) (this rank's block) of a value with as many rows as the result being computed."""
        return x.data[: hi - lo] if x.sharded else x.data[lo:hi]

    def _sharded(self, rows: torch.Tensor, shape) -> Ct:
        """Wrap this rank's block of result rows; blocks are padded to the common block size (all-gather needs it)."""
        per, lo, hi = self._part(_numel(shape))
        if rows.shape[0] != per:
            pad = torch.zeros((per, self.L), dtype=torch.int64, device=rows.device)
            if hi > lo:
                pad[: hi - lo] = rows[: hi - lo]
            rows = pad
        return Ct(rows, tuple(shape), True)

    def _operand(self, x: Ct, row_map: Optional[torch.Tensor], lo: int, hi: int):
        """(rows, map) with which result rows [lo, hi) read `x`: its own block when the map is the identity, else
        all of x (gathered if need be) through the block's slice of the map."""
        if row_map is None:
            return self._mine(x, lo, hi), None
        return self._full(x), row_map[lo:hi]

    def _rowwise(self, x: Ct, row_map, shape, fn) -> Ct:
        """out[e] = fn(x[row_map[e]], e): one result row per (mapped) operand row; fn(rows, map, lo, hi) -> rows."""
        n = _numel(shape)
        if not (x.sharded and self._shardable(n)):      # replicated operand: every rank computes all rows, nothing is sent
            return Ct(fn(self._full(x), row_map, 0, n), tuple(shape))
        per, lo, hi = self._part(n)
        d, m = self._operand(x, row_map, lo, hi)
        return self._sharded(fn(d, m, lo, hi), shape)

    # -- device upload helpers (cached per op)
    def _dev(self, arr: Optional[np.ndarray], dtype=None) -> Optional[torch.Tensor]:
        if arr is None:
            return None
        a = np.array(arr if dtype is None else arr.astype(dtype), copy=True, order="C")
        return torch.from_numpy(a).to(self.device)

    def _static_clear(self, names) -> bool:
        """Are these clear operands the same in every run?  Compile-time constants are; and when the function has no
        clear argument at all, every clear value is derived from constants only.  (A table, weight or index that comes
        from a clear ARGUMENT must be rebuilt per call.)"""
        return self._no_clear_args or all(n in self._consts for n in names)

    def _plan(self, op: Op, build: Callable[[], dict]) -> dict:
        key = (id(op), self._res)
        if key not in self._plans:
            self._plans[key] = build()
        return self._plans[key]

    def _encode(self, c: np.ndarray) -> np.ndarray:
        """clear ints -> plaintexts: (c mod 2^(p+1)) << (63 - p), or the residue encoding of the current lane
        (c mod m) * 2^64 / m; as int64 bit patterns."""
        if self.moduli:
            from . import wide

            return wide.encode_residue(c, self.moduli[self._res]).view(np.int64)
        p = self.params.p
        v = (np.asarray(c, dtype=np.int64).astype(np.uint64) & np.uint64((1 << (p + 1)) - 1)) << np.uint64(63 - p)
        return v.view(np.int64)

    def _weights(self, w: np.ndarray) -> np.ndarray:
        """Clear multipliers as applied to the current lane (centred residues mod m keep the noise growth small)."""
        w = np.asarray(w, dtype=np.int64)
        return _centered(w, self.moduli[self._res]) if self.moduli else w

    # ---------------------------------------------------------------- run ---------
    def run(self, ev: EvalKeys, args: List) -> List:
        """args: one entry per function argument (Ct for encrypted, numpy array / int for clear).  One run at a
        time per executor (the evaluation keys and the CRT lane of a run are executor state)."""
        with self._run_lock:
            return self._run(ev, args)

    def _run(self, ev: EvalKeys, args: List) -> List:
        R = self.lanes
        envs: List[Dict[str, object]] = [{} for _ in range(R)]
        for (name, t), a in zip(self.program.args, args):
            if t.encrypted:
                if not isinstance(a, Ct):
                    raise RuntimeError(f"argument {name} must be encrypted")
                if tuple(a.shape) != t.dims:
                    raise RuntimeError(f"argument {name}: shape {a.shape} != {t.dims}")
                rows = a.data.shape[0] // R
                if rows * R != a.data.shape[0] or rows != max(t.size, 1):
                    raise RuntimeError(f"argument {name}: {a.data.shape[0]} ciphertexts for {t.size} elements x {R} residues")
                for i in range(R):       # residue-major rows: [residue][element]; own buffers (rows are an odd
                    # number of words long, a row slice would break the 16-byte alignment the kernels vectorise on)
                    envs[i][name] = a if R == 1 else Ct(a.data[i * rows:(i + 1) * rows].clone(), a.shape)
            else:
                for i in range(R):
                    envs[i][name] = np.asarray(a, dtype=np.int64).reshape(t.dims)
        for env in envs:
            env.update(self._consts)
        self._ev = ev
        n_pbs = 0
        self._comm = {k: 0 for k in self._comm}
        self._comm_events = []
        for op in self.program.ops:
            if op.name == "arith.constant":
                continue
            base = op.name.split(".", 1)[1]
            h = self.handlers.get(base)
            if h is None:
                raise RuntimeError(f"operation {op.name} is not supported by the B200 backend")
            if self.moduli and base in ("apply_lookup_table", "apply_mapped_lookup_table"):
                outs = self._tlu_wide(op, [[env[x] for x in op.operands] for env in envs])
                for env, o in zip(envs, outs):
                    env[op.result] = o
            else:
                for i, env in enumerate(envs):
                    self._res = i
                    env[op.result] = h(op, [env[x] for x in op.operands])
                self._res = 0
            if base in ("apply_lookup_table", "apply_mapped_lookup_table"):
                n_pbs += op.result_type.size
        results = []
        for r in self.program.returns:
            vals = [env[r] for env in envs]
            if isinstance(vals[0], Ct):        # every rank returns the complete result
                vals = [Ct(self._full(v), v.shape) for v in vals]
            if R > 1 and isinstance(vals[0], Ct):
                results.append(Ct(torch.cat([v.data for v in vals], dim=0), vals[0].shape))
            else:
                results.append(vals[0])
        self.last_stats = {"n_pbs": n_pbs, **self._comm}
        if self._comm_events:
            torch.cuda.current_stream().synchronize()
            self.last_stats["comm_ms"] = float(sum(a.elapsed_time(b) for a, b in self._comm_events))
        return results

    # ------------------------------------------------------------- elementwise ----
    def _binary_ct(self, op: Op, a: Ct, b: Ct, fn) -> Ct:
        shape = op.result_type.dims
        plan = self._plan(op, lambda: {"ia": self._dev(_row_map(a.shape, shape)), "ib": self._dev(_row_map(b.shape, shape))})
        n = _numel(shape)
        if not ((a.sharded or b.sharded) and self._shardable(n)):
            return Ct(fn(self._full(a), self._full(b), self.L, plan["ia"], plan["ib"], n), shape)
        per, lo, hi = self._part(n)
        da, ia = self._operand(a, plan["ia"], lo, hi)
        db, ib = self._operand(b, plan["ib"], lo, hi)
        return self._sharded(fn(da, db, self.L, ia, ib, hi - lo), shape)

    def _add_eint(self, op, v):          # node_converter.py:216-242
        return self._binary_ct(op, v[0], v[1], E.lin_add)

    def _sub_eint(self, op, v):          # node_converter.py:1062-1092
        return self._binary_ct(op, v[0], v[1], E.lin_sub)

    def _clear_operand(self, op: Op, c, shape, encode: bool, negate: bool = False):
        """Broadcast a clear operand to the result shape (encoded as plaintexts when `encode`);
        compile-time constants are uploaded once per program."""
        cname = [x for x, t in zip(op.operands, op.operand_types) if not t.encrypted][0]

        def build():
            cc = np.broadcast_to(np.asarray(c, dtype=np.int64), shape).reshape(-1)
            cc = -cc if negate else cc
            return self._dev(self._encode(cc) if encode else np.ascontiguousarray(cc))

        if not self._static_clear([cname]):
            return build()
        key = ("clear", id(op), self._res)
        if key not in self._plans:
            self._plans[key] = build()
        return self._plans[key]

    def _ct_map(self, op: Op, a: Ct, shape):
        key = ("ia", id(op), self._res)
        if key not in self._plans:
            self._plans[key] = self._dev(_row_map(a.shape, shape))
        return self._plans[key]

    def _add_eint_int(self, op, v):      # ct + clear: plaintext added to the body only
        shape = op.result_type.dims
        pt = self._clear_operand(op, v[1], shape, encode=True)
        return self._rowwise(v[0], self._ct_map(op, v[0], shape), shape,
                             lambda d, m, lo, hi: E.lin_add_plain(d, pt[lo:hi], self.L, m))

    def _sub_eint_int(self, op, v):      # ct - clear
        shape = op.result_type.dims
        pt = self._clear_operand(op, v[1], shape, encode=True, negate=True)
        return self._rowwise(v[0], self._ct_map(op, v[0], shape), shape,
                             lambda d, m, lo, hi: E.lin_add_plain(d, pt[lo:hi], self.L, m))

    def _sub_int_eint(self, op, v):      # clear - ct
        shape = op.result_type.dims
        pt = self._clear_operand(op, v[0], shape, encode=True)
        return self._rowwise(v[1], self._ct_map(op, v[1], shape), shape,
                             lambda d, m, lo, hi: E.lin_plain_sub(d, pt[lo:hi], self.L, m))

    def _neg_eint(self, op, v):          # node_converter.py:652-669
        return self._rowwise(v[0], None, v[0].shape, lambda d, m, lo, hi: E.lin_neg(d, self.L, None, hi - lo))

    def _mul_eint_int(self, op, v):      # ct * clear int (all words), node_converter.py:630-650
        shape = op.result_type.dims
        c = self._clear_operand(op, self._weights(v[1]), shape, encode=False)
        return self._rowwise(v[0], self._ct_map(op, v[0], shape), shape,
                             lambda d, m, lo, hi: E.lin_mul_clear(d, c[lo:hi], self.L, m))

    def _zero(self, op, v):              # trivial encryption of 0: all-zero LWE (node_converter.py:1232-1248)
        shape = op.result_type.dims
        return Ct(torch.zeros((_numel(shape), self.L), dtype=torch.int64, device=self.device), shape)

    # --------------------------------------------------------------- lookups ------
    def _tlu(self, op, v):               # keyswitch + PBS per element (node_converter.py:1129-1208)
        x: Ct = v[0]
        p = self.params

        def build():
            luts = np.asarray(v[1], dtype=np.int64)
            if luts.ndim == 1:
                luts = luts[None, :]
            if luts.shape[1] != (1 << p.p):
                raise RuntimeError(f"lookup table of {luts.shape[1]} entries on {p.p}-bit ciphertexts")
            acc = E.build_lut_accumulators(p, luts)
            idx = None
            if op.name.endswith("apply_mapped_lookup_table"):
                m = np.broadcast_to(np.asarray(v[2], dtype=np.int64), x.shape).reshape(-1)
                idx = self._dev(m.astype(np.int32))
            return {"luts": E.from_np_u64(acc, self.device), "idx": idx}

        # a table or map that arrives as a clear function argument can change between runs: cache constants only
        clear_names = [n for n, t in zip(op.operands, op.operand_types) if not t.encrypted]
        plan = self._plan(op, build) if self._static_clear(clear_names) else build()
        shape = op.result_type.dims
        n = _numel(shape)
        if not self._shardable(n):
            return Ct(E.table_lookup(self._ev, self._full(x), plan["luts"], plan["idx"]), shape)
        # sharded by ciphertext: this rank bootstraps its block, nothing is exchanged
        per, lo, hi = self._part(n)
        if hi > lo:
            idx = None if plan["idx"] is None else plan["idx"][lo:hi]
            rows = E.table_lookup(self._ev, self._mine(x, lo, hi), plan["luts"], idx)
        else:
            rows = torch.zeros((0, self.L), dtype=torch.int64, device=self.device)
        return self._sharded(rows, shape)

    def _tlu_wide(self, op, lanes_v):
        """Lookup on CRT residues (wide.py): bit extraction + circuit bootstrapping once, vertical packing per output
        residue.  lanes_v[i] = operand values of lane i (ciphertext residues differ, clear operands are shared)."""
        from . import wide

        p = self.params
        v0 = lanes_v[0]
        xs = [v[0] for v in lanes_v]
        self._res = 0

        def build():
            tables = np.asarray(v0[1], dtype=np.int64)
            if tables.ndim == 1:
                tables = tables[None, :]
            if tables.shape[1] != (1 << p.p):
                raise RuntimeError(f"lookup table of {tables.shape[1]} entries on {p.p}-bit ciphertexts")
            luts = wide.build_crt_luts(tables, self.moduli)
            idx = None
            if op.name.endswith("apply_mapped_lookup_table"):
                m = np.broadcast_to(np.asarray(v0[2], dtype=np.int64), xs[0].shape).reshape(-1)
                idx = self._dev(m.astype(np.int64))
            return {"pool": wide.lut_pool(p, luts, self.device), "polys": max(luts.shape[2] // p.N, 1), "idx": idx}

        clear_names = [n for n, t in zip(op.operands, op.operand_types) if not t.encrypted]
        plan = self._plan(op, build) if self._static_clear(clear_names) else build()
        rows = [self._full(x) for x in xs]
        n = rows[0].shape[0]
        if self.world == 1:
            outs = wide.wide_table_lookup(self._ev, rows, plan["pool"], plan["polys"], plan["idx"])
        else:
            import torch.distributed as dist

            per = (n + self.world - 1) // self.world
            lo = min(self.rank * per, n)
            hi = min(lo + per, n)
            mine = [torch.zeros((per, self.L), dtype=torch.int64, device=self.device) for _ in rows]
            if hi > lo:
                sub = wide.wide_table_lookup(self._ev, [r[lo:hi].clone() for r in rows], plan["pool"], plan["polys"],
                                             None if plan["idx"] is None else plan["idx"][lo:hi].contiguous())
                for m_, s_ in zip(mine, sub):
                    m_[: hi - lo] = s_
            outs = []
            for m_ in mine:
                full = torch.empty((per * self.world, self.L), dtype=torch.int64, device=self.device)
                dist.all_gather_into_tensor(full, m_, group=self.group)
                outs.append(full[:n].contiguous())
        return [Ct(o, op.result_type.dims) for o in outs]

    # ---------------------------------------------------- clear-weighted linear ---
    def _aligned(self, idx: np.ndarray, n_in: int) -> bool:
        """True when, on EVERY rank, the result rows of its block read only operand rows of its own block (so a
        sharded operand needs no all-gather).  Decided from the whole index table: all ranks take the same branch."""
        if not self._shardable(n_in):
            return False
        n_out = idx.shape[0]
        for r in range(self.world):
            _, lo, hi = self._part(n_out, r)
            _, xlo, xhi = self._part(n_in, r)
            sub = idx[lo:hi]
            ok = sub[sub >= 0]
            if ok.size and (int(ok.min()) < xlo or int(ok.max()) >= xhi):
                return False
        return True

    def _gather(self, op: Op, x: Ct, build: Callable[[], Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]],
                reduce_ok: bool = False) -> Ct:
        """out[e] = sum_j w[e, j] * x[idx[e, j]] (+ bias): dot / matmul / conv2d / sum and friends."""
        n_in = _numel(x.shape)

        def mk():
            idx, w, bias = build()
            idx = np.ascontiguousarray(idx).astype(np.int64)
            d = {"idx": self._dev(idx.astype(np.int32)), "w": self._dev(self._weights(w).astype(np.int64)),
                 "bias": None if bias is None else self._dev(self._encode(bias)), "n_out": idx.shape[0]}
            if self.world > 1:
                n_out = idx.shape[0]
                d["aligned"] = self._shardable(n_out) and self._aligned(idx, n_in)
                if d["aligned"]:
                    _, lo, hi = self._part(n_out)
                    _, xlo, _ = self._part(n_in)
                    sub = idx[lo:hi]
                    d["idx_local"] = self._dev(np.where(sub >= 0, sub - xlo, -1).astype(np.int32))
                # a reduction over the sharded axis: local partial sums + all-reduce instead of gathering the operand
                d["reduce"] = (reduce_ok and not d["aligned"] and self._shardable(n_in) and 2 * n_out <= n_in)
                if d["reduce"]:
                    _, xlo, xhi = self._part(n_in)
                    d["idx_part"] = self._dev(np.where((idx >= xlo) & (idx < xhi), idx - xlo, -1).astype(np.int32))
            return d

        clear_names = [n for n, t in zip(op.operands, op.operand_types) if not t.encrypted]
        if self._static_clear(clear_names):
            plan = self._plan(op, mk)
        else:
            plan = mk()
        shape = op.result_type.dims
        n_out = plan["n_out"]
        if self.world > 1 and x.sharded and plan.get("reduce"):
            import torch.distributed as dist

            _, xlo, xhi = self._part(n_in)
            part = E.lin_gather_mac(x.data[: xhi - xlo], plan["idx_part"], plan["w"], self.L, None)
            self._collective(lambda: dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group), "allreduce",
                             part.numel() * 8)
            if plan["bias"] is not None:
                part = E.lin_add_plain(part, plan["bias"], self.L, None)
            return Ct(part, shape)
        if self.world > 1 and x.sharded and plan.get("aligned"):
            # every block reads its own operand rows only (e.g. a matmul sharded by rows of x): stays sharded
            per, lo, hi = self._part(n_out)
            bias = None if plan["bias"] is None else plan["bias"][lo:hi]
            _, xlo, xhi = self._part(n_in)
            rows = E.lin_gather_mac(x.data[: max(xhi - xlo, 1)], plan["idx_local"], plan["w"][lo:hi], self.L, bias)
            return self._sharded(rows, shape)
        # reads rows of other blocks (conv2d, sums that do not reduce enough, ...): gather the operand once (cached on
        # the value) and compute every row on every rank -- the linear kernels are HBM-bound and cheap next to a lookup
        return Ct(E.lin_gather_mac(self._full(x), plan["idx"], plan["w"], self.L, plan["bias"]), shape)

    def _dot(self, op, v):               # node_converter.py:497-524 (vector . vector)
        x, c = v
        return self._gather(op, x, lambda: (np.arange(_numel(x.shape))[None, :], np.asarray(c).reshape(1, -1), None),
                            reduce_ok=True)

    def _matmul(self, op, v):            # node_converter.py:596-616, numpy matmul broadcasting
        eint_left = op.name.endswith("matmul_eint_int")
        x: Ct = v[0] if eint_left else v[1]
        W = np.asarray(v[1] if eint_left else v[0], dtype=np.int64)
        out_shape = op.result_type.dims
        if eint_left and len(x.shape) == 2 and W.ndim == 2 and W.shape[0] * 16 * 8 <= 48 * 1024:
            Mr, K = x.shape
            Nc = W.shape[1]
            rows_per = Mr // self.world
            # output row m needs input row m only: with whole rows per block the matmul stays local
            local_ok = (x.sharded and self._shardable(Mr * Nc) and Mr % self.world == 0
                        and (rows_per * K) % 2 == 0 and (rows_per * Nc) % 2 == 0)
            wname = op.operands[1]
            if self._static_clear([wname]):
                Wd = self._plan(op, lambda: {"W": self._dev(self._weights(W))})["W"]
            else:
                Wd = self._dev(self._weights(W))
            if local_ok:
                _, xlo, xhi = self._part(Mr * K)
                rows = E.matmul_ct_clear(self._mine(x, xlo, xhi), Wd, rows_per, K, Nc, self.L)
                return self._sharded(rows, out_shape)
            return Ct(E.matmul_ct_clear(self._full(x), Wd, Mr, K, Nc, self.L), out_shape)

        def build():
            xi = np.arange(_numel(x.shape), dtype=np.int64).reshape(x.shape)
            a, b = (xi, W) if eint_left else (W, xi)
            a2 = a[None, :] if a.ndim == 1 else a
            b2 = b[:, None] if b.ndim == 1 else b
            Mr, K, Nc = a2.shape[-2], a2.shape[-1], b2.shape[-1]
            batch = np.broadcast_shapes(a2.shape[:-2], b2.shape[:-2])
            A = np.broadcast_to(a2, batch + (Mr, K))[..., :, None, :]            # [.., M, 1, K]
            Bm = np.swapaxes(np.broadcast_to(b2, batch + (K, Nc)), -1, -2)[..., None, :, :]  # [.., 1, N, K]
            full = batch + (Mr, Nc, K)
            A, Bm = np.broadcast_to(A, full), np.broadcast_to(Bm, full)
            idx, w = (A, Bm) if eint_left else (Bm, A)
            return idx.reshape(-1, K), w.reshape(-1, K), None

        return self._gather(op, x, build)

    def _conv2d(self, op, v):            # node_converter.py:436-483; semantics concrete/onnx/convolution.py:433-520
        x: Ct = v[0]
        W = np.asarray(v[1], dtype=np.int64)
        bias = np.asarray(v[2], dtype=np.int64) if len(v) == 3 else None
        a = op.attrs
        pads = [int(t) for t in np.asarray(a["padding"][0] if isinstance(a.get("padding"), tuple) else a.get("padding", [0, 0, 0, 0])).reshape(-1)]
        strides = [int(t) for t in np.asarray(a["strides"][0] if isinstance(a.get("strides"), tuple) else a.get("strides", [1, 1])).reshape(-1)]
        dil = [int(t) for t in np.asarray(a["dilations"][0] if isinstance(a.get("dilations"), tuple) else a.get("dilations", [1, 1])).reshape(-1)]
        group = int(a.get("group", 1))
        out_shape = op.result_type.dims

        def build():
            Nb, Cin, H, Wd = x.shape
            F, Cg, KH, KW = W.shape
            _, _, OH, OW = out_shape
            xi = np.arange(_numel(x.shape), dtype=np.int64).reshape(x.shape)
            K = Cg * KH * KW
            idx = np.full((Nb, F, OH, OW, K), -1, dtype=np.int64)
            w = np.zeros((Nb, F, OH, OW, K), dtype=np.int64)
            fpg = F // group
            oh = np.arange(OH)[:, None]
            ow = np.arange(OW)[None, :]
            for f in range(F):
                g = f // fpg
                for c in range(Cg):
                    for kh in range(KH):
                        for kw in range(KW):
                            kk = (c * KH + kh) * KW + kw
                            ih = oh * strides[0] - pads[0] + kh * dil[0]
                            iw = ow * strides[1] - pads[1] + kw * dil[1]
                            ok = (ih >= 0) & (ih < H) & (iw >= 0) & (iw < Wd)
                            ihc, iwc = np.clip(ih, 0, H - 1), np.clip(iw, 0, Wd - 1)
                            src = xi[:, g * Cg + c][:, ihc, iwc]                 # [Nb, OH, OW]
                            idx[:, f, :, :, kk] = np.where(ok[None], src, -1)
                            w[:, f, :, :, kk] = W[f, c, kh, kw]
            b = None
            if bias is not None:
                b = np.broadcast_to(bias.reshape(1, F, 1, 1), out_shape).reshape(-1)
            return idx.reshape(-1, K), w.reshape(-1, K), b

        return self._gather(op, x, build)

    def _sum(self, op, v):               # node_converter.py:1094-1127
        x: Ct = v[0]
        axes = [int(t) for t in op.attrs.get("axes", [])]
        keep = bool(op.attrs.get("keep_dims", False))
        out_shape = op.result_type.dims

        def build():
            xi = np.arange(_numel(x.shape), dtype=np.int64).reshape(x.shape)
            ax = tuple(axes) if axes else tuple(range(len(x.shape)))
            moved = np.moveaxis(xi, ax, tuple(range(len(x.shape) - len(ax), len(x.shape))))
            K = int(np.prod([x.shape[t] for t in ax], dtype=np.int64)) if ax else 1
            idx = moved.reshape(-1, K)
            return idx, np.ones_like(idx), None

        out = self._gather(op, x, build, reduce_ok=True)
        _ = keep
        return Ct(out.data, out_shape, out.sharded)

    # ------------------------------------------------------------ data movement ---
    def _gather_rows(self, op: Op, srcs: List[Ct], build_index: Callable[[], np.ndarray], shape) -> Ct:
        """out rows = concat(srcs) rows picked by an index map computed with numpy on the host."""
        def mk():
            idx = build_index().reshape(-1).astype(np.int64)
            ident = len(srcs) == 1 and idx.size == _numel(srcs[0].shape) and bool(np.array_equal(idx, np.arange(idx.size)))
            return {"idx": self._dev(idx), "identity": ident}

        # index operands that arrive as clear function arguments can change between runs: cache constants only
        clear_names = [n for n, t in zip(op.operands, op.operand_types) if not t.encrypted]
        plan = self._plan(op, mk) if self._static_clear(clear_names) else mk()
        if plan["identity"]:                    # a pure re-interpretation of the same rows keeps its placement
            return self._reshaped(srcs[0], shape)
        data = self._full(srcs[0]) if len(srcs) == 1 else torch.cat([self._full(s) for s in srcs], dim=0)
        return Ct(torch.index_select(data, 0, plan["idx"]), shape)

    def _reshaped(self, x: Ct, shape) -> Ct:
        out = Ct(x.data, tuple(shape), x.sharded)
        if getattr(x, "_gathered", None) is not None:
            out._gathered = x._gathered
        return out

    @staticmethod
    def _ids(cts: List[Ct]) -> List[np.ndarray]:
        out, base = [], 0
        for c in cts:
            n = _numel(c.shape)
            out.append(np.arange(base, base + n, dtype=np.int64).reshape(c.shape))
            base += n
        return out

    def _clear_or_ct(self, op, vals, fn):
        """Apply a pure data-movement numpy function either on clear arrays or on row ids."""
        if all(not isinstance(t, Ct) for t in vals):
            return fn([np.asarray(t) for t in vals])
        cts = [t for t in vals if isinstance(t, Ct)]
        if len(cts) != len(vals):
            raise RuntimeError(f"{op.name}: mixing clear and encrypted operands is not supported")
        return self._gather_rows(op, cts, lambda: fn(self._ids(cts)), op.result_type.dims)

    def _concat(self, op, v):            # node_converter.py:341-394
        axis = int(op.attrs.get("axis", 0))
        return self._clear_or_ct(op, v, lambda a: np.concatenate(a, axis=axis))

    def _transpose(self, op, v):         # node_converter.py:1210-1230 (empty axes = reverse)
        axes = [int(t) for t in op.attrs.get("axes", [])]
        return self._clear_or_ct(op, v[:1], lambda a: np.transpose(a[0], axes if axes else None))

    def _reshape(self, op, v):           # tensor.collapse_shape / expand_shape (node_converter.py:725-842)
        shape = op.result_type.dims
        if isinstance(v[0], Ct):
            return self._reshaped(v[0], shape)
        return np.asarray(v[0]).reshape(shape)

    def _extract(self, op, v):           # tensor.extract with constant indices (node_converter.py:917-1039)
        index = tuple(int(np.asarray(t)) for t in v[1:])
        return self._clear_or_ct(op, v[:1], lambda a: a[0][index])

    def _extract_slice(self, op, v):
        sl = _static_slices(op.attrs["offsets"], op.attrs["sizes"], op.attrs["strides"])
        shape = op.result_type.dims
        return self._clear_or_ct(op, v[:1], lambda a: a[0][sl].reshape(shape))

    def _insert_slice(self, op, v):      # tensor.insert_slice (assign.static, node_converter.py:855-915)
        sl = _static_slices(op.attrs["offsets"], op.attrs["sizes"], op.attrs["strides"])

        def fn(a):
            out = a[1].copy()
            out[sl] = a[0].reshape(out[sl].shape)
            return out

        return self._clear_or_ct(op, v, fn)

    def _from_elements(self, op, v):     # tensor.from_elements (node_converter.py:266-306)
        shape = op.result_type.dims
        return self._clear_or_ct(op, v, lambda a: np.stack([t.reshape(()) for t in a]).reshape(shape))
