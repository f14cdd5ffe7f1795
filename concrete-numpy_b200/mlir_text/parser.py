"""
Parser for the FHE / FHELinalg MLIR text the reference hands to `JITSupport.compile`
(concrete/numpy/compilation/server.py:135-154).  Produces a flat `Program` (SSA op list) that
executor.py walks.  Accepts the generic op form for FHE / FHELinalg ops and the custom assembly
of `arith.constant`, `tensor.*`, `func.func` and `return`.
"""

import ast
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class MlirSyntaxError(RuntimeError):
    """Raised for text this backend cannot interpret (surfaced like the wheel's RuntimeError)."""


@dataclass(frozen=True)
class TType:
    kind: str                         # "eint" | "int" | "index"
    width: int                        # bit width (0 for index)
    shape: Optional[Tuple[int, ...]]  # None for scalars

    @property
    def encrypted(self) -> bool:
        return self.kind == "eint"

    @property
    def dims(self) -> Tuple[int, ...]:
        return () if self.shape is None else self.shape

    @property
    def size(self) -> int:
        n = 1
        for d in self.dims:
            n *= d
        return n


@dataclass
class Op:
    name: str
    result: Optional[str]
    operands: List[str]
    attrs: Dict[str, Any]
    result_type: Optional[TType]
    operand_types: List[TType] = field(default_factory=list)


@dataclass
class Program:
    name: str
    args: List[Tuple[str, TType]]
    ops: List[Op]
    returns: List[str]
    result_types: List[TType]


_SCALAR_RE = re.compile(r"^(?:!FHE\.eint<(\d+)>|i(\d+)|(index))$")


def parse_type(text: str) -> TType:
    text = text.strip()
    if text.startswith("tensor<") and text.endswith(">"):
        inner = text[len("tensor<"):-1]
        dims = []
        while True:
            m = re.match(r"^(\d+)x", inner)
            if not m:
                break
            dims.append(int(m.group(1)))
            inner = inner[m.end():]
        elt = parse_type(inner)
        if elt.shape is not None:
            raise MlirSyntaxError(f"nested tensor type: {text}")
        return TType(elt.kind, elt.width, tuple(dims))
    m = _SCALAR_RE.match(text)
    if not m:
        raise MlirSyntaxError(f"unsupported type: {text!r}")
    if m.group(1):
        return TType("eint", int(m.group(1)), None)
    if m.group(2):
        return TType("int", int(m.group(2)), None)
    return TType("index", 0, None)


def split_top(text: str, sep: str = ",") -> List[str]:
    """Split on `sep` at bracket depth 0 (brackets: () [] {} <>, '->' is not a bracket)."""
    out, depth, cur, i = [], 0, [], 0
    while i < len(text):
        ch = text[i]
        if ch == "-" and text[i:i + 2] == "->":
            cur.append("->")
            i += 2
            continue
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
        i += 1
    tail = "".join(cur).strip()
    if tail or out:
        out.append(tail)
    return [p for p in out if p != ""]


def _find_matching(text: str, start: int) -> int:
    """Index of the bracket closing the one at `start`."""
    pairs = {"(": ")", "[": "]", "{": "}", "<": ">"}
    open_ch = text[start]
    close_ch = pairs[open_ch]
    depth = 0
    i = start
    while i < len(text):
        if text[i:i + 2] == "->":
            i += 2
            continue
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise MlirSyntaxError(f"unbalanced brackets in: {text!r}")


def parse_attr(text: str):
    """Attribute text -> python value (int, bool, list, or (np.ndarray, TType) for dense)."""
    text = text.strip()
    if text.startswith("dense<"):
        end = _find_matching(text, len("dense"))
        body = text[len("dense<"):end]
        rest = text[end + 1:].strip()
        if not rest.startswith(":"):
            raise MlirSyntaxError(f"dense attribute without type: {text!r}")
        ttype = parse_type(rest[1:])
        body = body.strip()
        if body.startswith('"0x') or body.startswith('"0X'):
            # the MLIR printer switches to a hex blob above 64 elements (every lookup table of >= 7 bits, most weight
            # matrices): raw little-endian element storage, iN stored in ceil(N/8) bytes, i1 bit-packed; a single
            # element is a splat.  Values are sign-extended from N bits -- what the decimal form would have printed.
            if not body.endswith('"'):
                raise MlirSyntaxError(f"unterminated hex dense attribute: {text[:60]!r}...")
            try:
                raw = bytes.fromhex(body[3:-1])
            except ValueError as e:
                raise MlirSyntaxError(f"bad hex dense attribute: {e}") from None
            n = int(np.prod(ttype.dims, dtype=np.int64)) if ttype.dims else 1
            w = int(ttype.width)
            if w == 1:
                bits = np.unpackbits(np.frombuffer(raw, dtype=np.uint8), bitorder="little")
                vals = [int(b) for b in bits[:n]] if len(raw) == (n + 7) // 8 else [int(bits[0])]
            else:
                bw = (w + 7) // 8
                if len(raw) % bw:
                    raise MlirSyntaxError(f"hex dense attribute of {len(raw)} bytes for {bw}-byte elements")
                vals = []
                for i in range(len(raw) // bw):
                    v = int.from_bytes(raw[i * bw:(i + 1) * bw], "little") & ((1 << w) - 1)
                    vals.append(v - (1 << w) if v >> (w - 1) else v)
            if len(vals) == 1 and n != 1:
                vals = vals * n
            if len(vals) != n:
                raise MlirSyntaxError(f"hex dense attribute holds {len(vals)} elements, type wants {n}")
            data = np.array(vals, dtype=object)
        elif body.startswith("["):
            data = np.array(ast.literal_eval(body), dtype=object)
        elif body in ("true", "false"):
            data = np.full(ttype.dims, 1 if body == "true" else 0, dtype=object)
        else:
            data = np.full(ttype.dims, int(body), dtype=object)
        data = data.reshape(ttype.dims)
        # reduce mod 2^64 and reinterpret as int64 (values are used mod 2^width downstream)
        flat = np.array([int(v) & 0xFFFFFFFFFFFFFFFF for v in data.reshape(-1)], dtype=np.uint64)
        return flat.view(np.int64).reshape(ttype.dims), ttype
    if text in ("true", "false"):
        return text == "true"
    if text.startswith("["):
        inner = text[1:_find_matching(text, 0)]
        return [parse_attr(p) for p in split_top(inner)]
    m = re.match(r"^(-?\d+)(?:\s*:\s*(\S+))?$", text)
    if m:
        return int(m.group(1))
    raise MlirSyntaxError(f"unsupported attribute: {text!r}")


def _parse_attr_dict(text: str) -> Dict[str, Any]:
    out = {}
    for item in split_top(text):
        k, _, v = item.partition("=")
        out[k.strip()] = parse_attr(v)
    return out


def _int_list(text: str) -> List[int]:
    return [int(x) for x in ast.literal_eval(text)]


_GENERIC_RE = re.compile(r'^"([A-Za-z_.0-9]+)"\((.*?)\)\s*(.*)$', flags=re.S)


def _parse_op_rhs(rhs: str, result: Optional[str]) -> Op:
    rhs = rhs.strip()
    m = _GENERIC_RE.match(rhs)
    if m:
        name, operand_text, rest = m.group(1), m.group(2), m.group(3).strip()
        operands = [o.strip() for o in operand_text.split(",") if o.strip()]
        attrs = {}
        if rest.startswith("{"):
            end = _find_matching(rest, 0)
            attrs = _parse_attr_dict(rest[1:end])
            rest = rest[end + 1:].strip()
        if not rest.startswith(":"):
            raise MlirSyntaxError(f"missing type signature: {rhs!r}")
        sig = rest[1:].strip()
        if not sig.startswith("("):
            raise MlirSyntaxError(f"bad type signature: {rhs!r}")
        end = _find_matching(sig, 0)
        in_types = [parse_type(t) for t in split_top(sig[1:end])]
        out = sig[end + 1:].strip()
        if not out.startswith("->"):
            raise MlirSyntaxError(f"bad type signature: {rhs!r}")
        out = out[2:].strip()
        if out.startswith("("):
            out = out[1:_find_matching(out, 0)]
        return Op(name, result, operands, attrs, parse_type(out), in_types)

    if rhs.startswith("arith.constant"):
        value = parse_attr(rhs[len("arith.constant"):])
        if isinstance(value, tuple):
            data, ttype = value
            return Op("arith.constant", result, [], {"value": data}, ttype)
        m2 = re.match(r"^arith\.constant\s+(-?\d+|true|false)\s*:\s*(\S+)$", rhs)
        if not m2:
            raise MlirSyntaxError(f"bad constant: {rhs!r}")
        lit = m2.group(1)
        v = {"true": 1, "false": 0}.get(lit, None)
        v = int(lit) if v is None else v
        return Op("arith.constant", result, [], {"value": v}, parse_type(m2.group(2)))

    m = re.match(r"^tensor\.extract\s+(%[\w\-.]+)\[(.*?)\]\s*:\s*(.+)$", rhs)
    if m and not rhs.startswith("tensor.extract_slice"):
        src_t = parse_type(m.group(3))
        idx = [o.strip() for o in m.group(2).split(",") if o.strip()]
        return Op("tensor.extract", result, [m.group(1), *idx], {}, TType(src_t.kind, src_t.width, None), [src_t])

    m = re.match(r"^tensor\.extract_slice\s+(%[\w\-.]+)\s*(\[.*?\])\s*(\[.*?\])\s*(\[.*?\])\s*:\s*(.+?)\s+to\s+(.+)$", rhs)
    if m:
        return Op("tensor.extract_slice", result, [m.group(1)],
                  {"offsets": _int_list(m.group(2)), "sizes": _int_list(m.group(3)), "strides": _int_list(m.group(4))},
                  parse_type(m.group(6)), [parse_type(m.group(5))])

    m = re.match(r"^tensor\.insert_slice\s+(%[\w\-.]+)\s+into\s+(%[\w\-.]+)\s*(\[.*?\])\s*(\[.*?\])\s*(\[.*?\])\s*:\s*(.+?)\s+into\s+(.+)$", rhs)
    if m:
        return Op("tensor.insert_slice", result, [m.group(1), m.group(2)],
                  {"offsets": _int_list(m.group(3)), "sizes": _int_list(m.group(4)), "strides": _int_list(m.group(5))},
                  parse_type(m.group(7)), [parse_type(m.group(6)), parse_type(m.group(7))])

    m = re.match(r"^tensor\.(collapse_shape|expand_shape)\s+(%[\w\-.]+)\s*(\[.*\])\s*:\s*(.+?)\s+into\s+(.+)$", rhs)
    if m:
        return Op("tensor." + m.group(1), result, [m.group(2)],
                  {"reassociation": [list(g) for g in ast.literal_eval(m.group(3))]},
                  parse_type(m.group(5)), [parse_type(m.group(4))])

    m = re.match(r"^tensor\.from_elements\s+(.*?)\s*:\s*(.+)$", rhs)
    if m:
        operands = [o.strip() for o in m.group(1).split(",") if o.strip()]
        return Op("tensor.from_elements", result, operands, {}, parse_type(m.group(2)))

    raise MlirSyntaxError(f"unsupported operation: {rhs!r}")


def parse_module(text: str) -> Program:
    """Parse one `func.func @main` out of an MLIR module string."""
    m = re.search(r"func(?:\.func)?\s+@(\w+)\s*\(", text)
    if not m:
        raise MlirSyntaxError("no function found in MLIR module")
    name = m.group(1)
    open_paren = m.end() - 1
    close_paren = _find_matching(text, open_paren)
    args = []
    for item in split_top(text[open_paren + 1:close_paren]):
        an, _, at = item.partition(":")
        args.append((an.strip(), parse_type(at)))
    rest = text[close_paren + 1:]
    brace = rest.index("{")
    header = rest[:brace].strip()
    result_types: List[TType] = []
    if header.startswith("->"):
        h = header[2:].strip()
        if h.startswith("("):
            h = h[1:_find_matching(h, 0)]
        result_types = [parse_type(t) for t in split_top(h)]
    body = rest[brace + 1:]
    ops: List[Op] = []
    returns: List[str] = []
    # statements are one per line; a dense attribute may contain newlines only if hand-written,
    # which neither the reference nor our printer does
    for raw in body.split("\n"):
        line = raw.strip()
        if not line or line.startswith("//"):
            continue
        if line.startswith("}"):
            break
        if line.startswith("return") or line.startswith("func.return"):
            tail = line.split(None, 1)[1] if " " in line else ""
            names = tail.split(":")[0]
            returns = [o.strip() for o in names.split(",") if o.strip()]
            break
        lhs, eq, rhs = line.partition("=")
        if not eq or not lhs.strip().startswith("%"):
            raise MlirSyntaxError(f"cannot parse statement: {line!r}")
        ops.append(_parse_op_rhs(rhs, lhs.strip()))
    if not returns and result_types:
        raise MlirSyntaxError("function has results but no return statement")
    return Program(name, args, ops, returns, result_types)
