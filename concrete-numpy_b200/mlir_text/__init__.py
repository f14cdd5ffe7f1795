"""
FHE / FHELinalg MLIR text: builder stand-ins for the `mlir` + `concrete.lang` Python bindings
(ir.py) and the parser that turns the text back into a flat program (parser.py).

`install_shims()` registers the stand-ins (and this backend's `concrete.compiler`) in
`sys.modules`, which is all the reference frontend needs to run unmodified on top of this backend.
"""

import sys
import types

from . import ir


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package so that `import a.b.c` works
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _fhe_module():
    names = {
        "AddEintIntOp": "FHE.add_eint_int", "AddEintOp": "FHE.add_eint",
        "SubIntEintOp": "FHE.sub_int_eint", "SubEintIntOp": "FHE.sub_eint_int", "SubEintOp": "FHE.sub_eint",
        "NegEintOp": "FHE.neg_eint", "MulEintIntOp": "FHE.mul_eint_int",
        "ApplyLookupTableEintOp": "FHE.apply_lookup_table",
        "ZeroEintOp": "FHE.zero", "ZeroTensorOp": "FHE.zero_tensor",
    }
    ops = {k: ir._simple_op(v) for k, v in names.items()}
    return _module("concrete.lang.dialects.fhe", EncryptedIntegerType=ir.EncryptedIntegerType, **ops)


def _fhelinalg_module():
    names = {
        "AddEintIntOp": "FHELinalg.add_eint_int", "AddEintOp": "FHELinalg.add_eint",
        "SubIntEintOp": "FHELinalg.sub_int_eint", "SubEintIntOp": "FHELinalg.sub_eint_int",
        "SubEintOp": "FHELinalg.sub_eint", "NegEintOp": "FHELinalg.neg_eint",
        "MulEintIntOp": "FHELinalg.mul_eint_int",
        "ApplyLookupTableEintOp": "FHELinalg.apply_lookup_table",
        "ApplyMappedLookupTableEintOp": "FHELinalg.apply_mapped_lookup_table",
        "Dot": "FHELinalg.dot_eint_int",
        "MatMulEintIntOp": "FHELinalg.matmul_eint_int", "MatMulIntEintOp": "FHELinalg.matmul_int_eint",
    }
    ops = {k: ir._simple_op(v) for k, v in names.items()}
    return _module("concrete.lang.dialects.fhelinalg", SumOp=ir.SumOp, ConcatOp=ir.ConcatOp,
                   TransposeOp=ir.TransposeOp, Conv2dOp=ir.Conv2dOp, **ops)


def install_shims(force: bool = True) -> None:
    """Make `import mlir...`, `import concrete.lang...` and `import concrete.compiler` resolve to
    this backend.  Call before importing `concrete.numpy`."""
    from .. import compiler as compiler_module

    mlir_ir = _module(
        "mlir.ir",
        **{n: getattr(ir, n) for n in (
            "ArrayAttr", "Attribute", "BlockArgument", "BoolAttr", "Context", "DenseElementsAttr",
            "IndexType", "IntegerAttr", "IntegerType", "OpResult", "RankedTensorType", "Type",
            "InsertionPoint", "Location", "Module", "Value", "Operation")},
    )
    arith = _module("mlir.dialects.arith", ConstantOp=ir.ConstantOp)
    func = _module("mlir.dialects.func", FuncOp=ir.FuncOp, ReturnOp=ir.ReturnOp)
    tensor = _module("mlir.dialects.tensor", ExtractOp=ir.ExtractOp, ExtractSliceOp=ir.ExtractSliceOp,
                     InsertSliceOp=ir.InsertSliceOp, CollapseShapeOp=ir.CollapseShapeOp,
                     ExpandShapeOp=ir.ExpandShapeOp, FromElementsOp=ir.FromElementsOp)
    dialects = _module("mlir.dialects", arith=arith, func=func, tensor=tensor)
    mlir = _module("mlir", ir=mlir_ir, dialects=dialects)
    fhe, fhelinalg = _fhe_module(), _fhelinalg_module()
    lang_dialects = _module("concrete.lang.dialects", fhe=fhe, fhelinalg=fhelinalg)
    lang = _module("concrete.lang", dialects=lang_dialects, register_dialects=lambda ctx: None)
    mods = {
        "mlir": mlir, "mlir.ir": mlir_ir, "mlir.dialects": dialects, "mlir.dialects.arith": arith,
        "mlir.dialects.func": func, "mlir.dialects.tensor": tensor,
        "concrete.lang": lang, "concrete.lang.dialects": lang_dialects,
        "concrete.lang.dialects.fhe": fhe, "concrete.lang.dialects.fhelinalg": fhelinalg,
        "concrete.compiler": compiler_module,
    }
    for name, mod in mods.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
    bind_to_parent_package()


def bind_to_parent_package() -> None:
    """`import concrete.compiler` found in sys.modules does not set the attribute on the parent
    `concrete` namespace package; do it by hand once `concrete` has been imported."""
    parent = sys.modules.get("concrete")
    if parent is not None:
        for leaf in ("compiler", "lang"):
            mod = sys.modules.get(f"concrete.{leaf}")
            if mod is not None:
                setattr(parent, leaf, mod)
