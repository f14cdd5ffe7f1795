"""
A small in-memory IR with the slice of the `mlir.ir` / `mlir.dialects` / `concrete.lang` Python
API that the reference's MLIR emission uses (concrete/numpy/mlir/graph_converter.py:662-739,
node_converter.py), printing textual MLIR in the FHE / FHELinalg / tensor / arith / func dialects.

The real bindings live in the absent concrete-compiler wheel; with these stand-ins installed
(`install_shims()` in mlir_text/__init__.py) the reference's `GraphConverter.convert` runs
unmodified and its output string is what `JITSupport.compile` of this backend receives -- exactly
the hand-off of the reference (`Server.create(mlir, ...)`, server.py:69-166).

Printing follows upstream conventions: FHE / FHELinalg ops in generic form
(`"FHE.add_eint"(%a, %b) : (T, T) -> T`), arith / tensor / func ops in their custom assembly,
`arith.constant` results named `%c<value>_<type>` / `%cst`, other results numbered `%0, %1, ...`
(golden: tests/mlir/test_graph_converter.py:473-485).
"""

import re
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

# ------------------------------------------------------------------ context stack --

_INSERTION_STACK: List["Block"] = []


class Context:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class Location:
    @staticmethod
    def unknown(context=None):
        return Location()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class InsertionPoint:
    def __init__(self, block: "Block"):
        self.block = block

    def __enter__(self):
        _INSERTION_STACK.append(self.block)
        return self

    def __exit__(self, *exc):
        _INSERTION_STACK.pop()
        return False


def _current_block() -> "Block":
    if not _INSERTION_STACK:
        raise RuntimeError("no insertion point is active")
    return _INSERTION_STACK[-1]


# ------------------------------------------------------------------------ types ---

class Type:
    def __init__(self, text: str):
        self.text = text

    def __str__(self):
        return self.text

    __repr__ = __str__

    def __eq__(self, other):
        return isinstance(other, Type) and other.text == self.text

    def __hash__(self):
        return hash(("type", self.text))

    @staticmethod
    def parse(text: str, context=None) -> "Type":
        return Type(text.strip())


class IntegerType(Type):
    @staticmethod
    def get_signless(width: int, context=None) -> Type:
        return Type(f"i{int(width)}")


class IndexType(Type):
    @staticmethod
    def parse(text: str, context=None) -> Type:
        return Type("index")

    @staticmethod
    def get(context=None) -> Type:
        return Type("index")


class RankedTensorType(Type):
    @staticmethod
    def get(shape: Sequence[int], element_type: Type) -> Type:
        dims = "".join(f"{int(d)}x" for d in shape)
        return Type(f"tensor<{dims}{element_type}>")


class EncryptedIntegerType(Type):
    @staticmethod
    def get(context, width: int) -> Type:
        return Type(f"!FHE.eint<{int(width)}>")


# ------------------------------------------------------------------- attributes ---

class Attribute:
    """Attribute = text (without the trailing ': type' when `type` is set) + optional type."""

    def __init__(self, text: str, type: Optional[Type] = None, value: Any = None):
        self.text = text
        self.type = type
        self.value = value

    def full(self) -> str:
        return f"{self.text} : {self.type}" if self.type is not None else self.text

    def in_array(self) -> str:
        # integer attrs of type i64 and index drop their type inside arrays
        if self.type is not None and str(self.type) in ("i64", "index") and not self.text.startswith("dense"):
            return self.text
        return self.full()

    def __str__(self):
        return self.full()

    __repr__ = __str__

    def __eq__(self, other):
        return isinstance(other, Attribute) and other.full() == self.full()

    def __hash__(self):
        return hash(("attr", self.full()))

    @staticmethod
    def parse(text: str, context=None) -> "Attribute":
        text = text.strip()
        m = re.match(r"^(dense<.*>)\s*:\s*(.+)$", text, flags=re.S)
        if m:
            return Attribute(re.sub(r"\s+", " ", m.group(1)), Type(m.group(2).strip()))
        return Attribute(text)


class IntegerAttr(Attribute):
    @staticmethod
    def get(type: Type, value) -> Attribute:
        return Attribute(str(int(value)), type, int(value))


class BoolAttr(Attribute):
    @staticmethod
    def get(value: bool, context=None) -> Attribute:
        return Attribute("true" if value else "false", None, bool(value))


class ArrayAttr(Attribute):
    @staticmethod
    def get(attributes: Sequence[Attribute], context=None) -> Attribute:
        return Attribute("[" + ", ".join(a.in_array() for a in attributes) + "]", None, list(attributes))


def _dense_body(arr: np.ndarray) -> str:
    flat = arr.reshape(-1)
    if flat.size > 0 and np.all(flat == flat[0]) and arr.ndim >= 1 and flat.size > 1:
        return str(int(flat[0]))   # splat form, as MLIR prints it
    return str(arr.tolist())


class DenseElementsAttr(Attribute):
    @staticmethod
    def get(array, type: Optional[Type] = None, context=None, signless=None, shape=None) -> Attribute:
        arr = np.asarray(array)
        if arr.dtype == np.uint64:
            arr = arr.view(np.int64)      # signless i64: printed as signed, like upstream
        if type is None:
            if arr.dtype.kind not in "iu":
                raise TypeError(f"unsupported dense element dtype {arr.dtype}")
            type = Type(f"i{arr.dtype.itemsize * 8}")
        tensor_type = RankedTensorType.get(arr.shape, type)
        return Attribute(f"dense<{_dense_body(arr)}>", tensor_type, arr)


# ----------------------------------------------------------------- values / ops ---

class Value:
    def __init__(self, type: Type, owner):
        self.type = type
        self.owner = owner

    def _root(self) -> Optional["Module"]:
        block = self.owner.block if isinstance(self.owner, Operation) else self.owner
        while block is not None and block.parent_op is not None:
            block = block.parent_op.block
        return block.module if block is not None else None

    def name(self) -> str:
        root = self._root()
        if root is not None:
            root.assign_names()
        return getattr(self, "_name", "%?")

    def __str__(self):
        if isinstance(self.owner, Operation):
            return f"Value({self.owner.to_text()})"
        return f"Value(<block argument> of type '{self.type}' at index: {self.arg_number})"

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class OpResult(Value):
    pass


class BlockArgument(Value):
    def __init__(self, type: Type, block: "Block", arg_number: int):
        super().__init__(type, block)
        self.arg_number = arg_number


class Block:
    def __init__(self, arg_types: Sequence[Type] = (), parent_op: Optional["Operation"] = None,
                 module: Optional["Module"] = None):
        self.operations: List[Operation] = []
        self.parent_op = parent_op
        self.module = module
        self.arguments = [BlockArgument(t, self, i) for i, t in enumerate(arg_types)]


def _as_value(x) -> Value:
    if isinstance(x, Value):
        return x
    if isinstance(x, Operation):
        return x.result
    raise TypeError(f"expected an MLIR value, got {type(x)}")


class Operation:
    """Generic operation; subclasses set OPNAME and may override `asm()` for custom assembly."""

    OPNAME = "builtin.unknown"

    def __init__(self, result_types: Sequence[Type], operands: Sequence[Any],
                 attributes: Optional[Dict[str, Attribute]] = None, insert: bool = True):
        self.operands = [_as_value(o) for o in operands]
        self.attributes = {k: v for k, v in (attributes or {}).items() if v is not None}
        self.results = [OpResult(t, self) for t in result_types]
        self.block: Optional[Block] = None
        if insert:
            block = _current_block()
            block.operations.append(self)
            self.block = block

    @property
    def result(self) -> OpResult:
        return self.results[0]

    # -- printing
    def _attr_dict(self) -> str:
        if not self.attributes:
            return ""
        return " {" + ", ".join(f"{k} = {v.full()}" for k, v in sorted(self.attributes.items())) + "}"

    def asm(self) -> str:
        ops = ", ".join(o.name() for o in self.operands)
        in_types = ", ".join(str(o.type) for o in self.operands)
        out_types = ", ".join(str(r.type) for r in self.results)
        if len(self.results) != 1:
            out_types = f"({out_types})"
        return f'"{self.OPNAME}"({ops}){self._attr_dict()} : ({in_types}) -> {out_types}'

    def to_text(self) -> str:
        if self.results:
            return f"{', '.join(r.name() for r in self.results)} = {self.asm()}"
        return self.asm()

    def suggested_name(self) -> Optional[str]:
        return None


def _simple_op(opname: str, n_operands: Optional[int] = None):
    class _Op(Operation):
        OPNAME = opname

        def __init__(self, result_type: Type, *operands, loc=None, ip=None):
            if len(operands) == 1 and isinstance(operands[0], (list, tuple)):
                operands = tuple(operands[0])
            if n_operands is not None and len(operands) != n_operands:
                raise TypeError(f"{opname} expects {n_operands} operands, got {len(operands)}")
            super().__init__([result_type], operands)

    _Op.__name__ = opname.replace(".", "_")
    return _Op


# ------------------------------------------------------------------ arith / func --

class ConstantOp(Operation):
    OPNAME = "arith.constant"

    def __init__(self, result_type: Type, value: Attribute, loc=None, ip=None):
        # the reference sometimes passes a mismatching result type (node_converter.py:1315);
        # like upstream's printer, the attribute's own type is what reaches the text
        actual = value.type if value.type is not None else result_type
        super().__init__([actual], [], {"value": value})
        self.value = value

    def asm(self) -> str:
        return f"arith.constant {self.value.full()}"

    def suggested_name(self) -> Optional[str]:
        v = self.value
        if v.text.startswith("dense") or v.type is None:
            return "cst"
        t = str(v.type)
        if t == "index":
            return f"c{v.text}"
        if t == "i1":
            return "true" if v.text != "0" else "false"
        return f"c{v.text}_{t}"


class ReturnOp(Operation):
    OPNAME = "func.return"

    def __init__(self, operands: Sequence[Value]):
        super().__init__([], operands)

    def asm(self) -> str:
        if not self.operands:
            return "return"
        return ("return " + ", ".join(o.name() for o in self.operands) + " : "
                + ", ".join(str(o.type) for o in self.operands))


class FuncOp(Operation):
    OPNAME = "func.func"

    def __init__(self, name: str, arg_types: Sequence[Type]):
        super().__init__([], [])
        self.sym_name = name
        self.body = Block(arg_types, parent_op=self)
        self.result_types: List[Type] = []

    @staticmethod
    def from_py_func(*arg_types: Type, results=None, name: Optional[str] = None):
        def decorator(fn):
            op = FuncOp(name or fn.__name__, arg_types)
            with InsertionPoint(op.body):
                ret = fn(*op.body.arguments)
                if ret is None:
                    values: List[Value] = []
                elif isinstance(ret, (Value, Operation)):
                    values = [_as_value(ret)]
                else:
                    values = [_as_value(v) for v in ret]
                ReturnOp(values)
            op.result_types = [v.type for v in values]
            return op

        return decorator

    def to_lines(self, indent: str) -> List[str]:
        args = ", ".join(f"%arg{a.arg_number}: {a.type}" for a in self.body.arguments)
        if len(self.result_types) == 1:
            res = f" -> {self.result_types[0]}"
        elif self.result_types:
            res = " -> (" + ", ".join(str(t) for t in self.result_types) + ")"
        else:
            res = ""
        lines = [f"{indent}func.func @{self.sym_name}({args}){res} {{"]
        for op in self.body.operations:
            lines.append(f"{indent}  {op.to_text()}")
        lines.append(f"{indent}}}")
        return lines


class Module:
    def __init__(self):
        self.body = Block(module=self)
        self._named_at = -1

    @staticmethod
    def create(loc=None) -> "Module":
        return Module()

    def _all_ops(self):
        for op in self.body.operations:
            yield op
            if isinstance(op, FuncOp):
                for inner in op.body.operations:
                    yield inner

    def assign_names(self):
        ops = list(self._all_ops())
        if self._named_at == len(ops):
            return
        for op in self.body.operations:
            if not isinstance(op, FuncOp):
                continue
            used: Dict[str, int] = {}
            counter = 0
            for a in op.body.arguments:
                a._name = f"%arg{a.arg_number}"
            for inner in op.body.operations:
                hint = inner.suggested_name()
                for r in inner.results:
                    if hint is None:
                        r._name = f"%{counter}"
                        counter += 1
                    else:
                        if hint in used:
                            # upstream uniquing: cst, cst_0, cst_1, ...
                            while True:
                                cand = f"{hint}_{used[hint]}"
                                used[hint] += 1
                                if cand not in used:
                                    break
                            used[cand] = 0
                            r._name = f"%{cand}"
                        else:
                            used[hint] = 0
                            r._name = f"%{hint}"
        self._named_at = len(ops)

    def __str__(self):
        self._named_at = -1
        self.assign_names()
        lines = ["module {"]
        for op in self.body.operations:
            if isinstance(op, FuncOp):
                lines.extend(op.to_lines("  "))
            else:
                lines.append("  " + op.to_text())
        lines.append("}")
        return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------- tensor ----

def _ints(attr: Attribute) -> str:
    return "[" + ", ".join(a.text for a in attr.value) + "]"


class ExtractOp(Operation):
    OPNAME = "tensor.extract"

    def __init__(self, result_type, tensor, indices, loc=None, ip=None):
        super().__init__([result_type], [tensor, *indices])

    def asm(self):
        idx = ", ".join(o.name() for o in self.operands[1:])
        return f"tensor.extract {self.operands[0].name()}[{idx}] : {self.operands[0].type}"


class ExtractSliceOp(Operation):
    OPNAME = "tensor.extract_slice"

    def __init__(self, result_type, source, offsets, sizes, strides, static_offsets, static_sizes,
                 static_strides, loc=None, ip=None):
        assert not offsets and not sizes and not strides, "dynamic slices are not emitted by the reference"
        super().__init__([result_type], [source])
        self.static = (static_offsets, static_sizes, static_strides)

    def asm(self):
        o, s, t = (_ints(a) for a in self.static)
        return (f"tensor.extract_slice {self.operands[0].name()}{o} {s} {t} : "
                f"{self.operands[0].type} to {self.result.type}")


class InsertSliceOp(Operation):
    OPNAME = "tensor.insert_slice"

    def __init__(self, result_type, source, dest, offsets, sizes, strides, static_offsets, static_sizes,
                 static_strides, loc=None, ip=None):
        assert not offsets and not sizes and not strides
        super().__init__([result_type], [source, dest])
        self.static = (static_offsets, static_sizes, static_strides)

    def asm(self):
        o, s, t = (_ints(a) for a in self.static)
        return (f"tensor.insert_slice {self.operands[0].name()} into {self.operands[1].name()}{o} {s} {t} : "
                f"{self.operands[0].type} into {self.operands[1].type}")


def _reassoc(attr: Attribute) -> str:
    return "[" + ", ".join("[" + ", ".join(a.text for a in grp.value) + "]" for grp in attr.value) + "]"


class CollapseShapeOp(Operation):
    OPNAME = "tensor.collapse_shape"

    def __init__(self, result_type, src, reassociation, loc=None, ip=None):
        super().__init__([result_type], [src])
        self.reassociation = reassociation

    def asm(self):
        return (f"tensor.collapse_shape {self.operands[0].name()} {_reassoc(self.reassociation)} : "
                f"{self.operands[0].type} into {self.result.type}")


class ExpandShapeOp(Operation):
    OPNAME = "tensor.expand_shape"

    def __init__(self, result_type, src, reassociation, loc=None, ip=None):
        super().__init__([result_type], [src])
        self.reassociation = reassociation

    def asm(self):
        return (f"tensor.expand_shape {self.operands[0].name()} {_reassoc(self.reassociation)} : "
                f"{self.operands[0].type} into {self.result.type}")


class FromElementsOp(Operation):
    OPNAME = "tensor.from_elements"

    def __init__(self, result_type, elements, loc=None, ip=None):
        super().__init__([result_type], list(elements))

    def asm(self):
        return ("tensor.from_elements " + ", ".join(o.name() for o in self.operands)
                + f" : {self.result.type}")


# ------------------------------------------------------------- FHE / FHELinalg ----

class _AttrOp(Operation):
    """Op with trailing keyword attributes."""

    ATTRS: Sequence[str] = ()
    N_OPERANDS = 1

    def __init__(self, result_type, *operands, loc=None, ip=None, **attrs):
        if len(operands) == 1 and isinstance(operands[0], (list, tuple)):
            operands = tuple(operands[0])
        unknown = set(attrs) - set(self.ATTRS)
        if unknown:
            raise TypeError(f"{self.OPNAME}: unexpected attributes {sorted(unknown)}")
        super().__init__([result_type], operands, attrs)


class SumOp(_AttrOp):
    OPNAME = "FHELinalg.sum"
    ATTRS = ("axes", "keep_dims")


class ConcatOp(_AttrOp):
    OPNAME = "FHELinalg.concat"
    ATTRS = ("axis",)


class TransposeOp(_AttrOp):
    OPNAME = "FHELinalg.transpose"
    ATTRS = ("axes",)


class Conv2dOp(Operation):
    OPNAME = "FHELinalg.conv2d"

    def __init__(self, result_type, input, weight, bias=None, padding=None, strides=None, dilations=None,
                 group=None, loc=None, ip=None):
        operands = [input, weight] + ([bias] if bias is not None else [])
        super().__init__([result_type], operands,
                         {"padding": padding, "strides": strides, "dilations": dilations, "group": group})
