"""MLIR text accepted by the parser beyond what the in-repo builder prints: the real MLIR printer switches dense
attributes to hex blobs above 64 elements (every lookup table of >= 7 bits, most weight matrices), so `Server.create` /
`Server.load` must decode them (raw little-endian element storage, iN in ceil(N/8) bytes, i1 bit-packed, splats)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from concrete_numpy_b200.mlir_text.parser import MlirSyntaxError, parse_attr, parse_module


def _hex(values, width):
    bw = (width + 7) // 8
    return "0x" + b"".join((int(v) & ((1 << (8 * bw)) - 1)).to_bytes(bw, "little") for v in values).hex().upper()


def test_hex_dense_lookup_table_equals_decimal_form():
    rng = np.random.default_rng(0)
    table = rng.integers(0, 256, 256)
    dec = f'dense<{table.tolist()}> : tensor<256xi64>'
    hx = f'dense<"{_hex(table, 64)}"> : tensor<256xi64>'
    a, ta = parse_attr(dec)
    b, tb = parse_attr(hx)
    assert ta == tb and a.dtype == np.int64 and np.array_equal(a, b)


def test_hex_dense_signed_narrow_weights_bits_and_splats():
    w = np.array([[-3, 2, -1], [5, -8, 7]])
    got, t = parse_attr(f'dense<"{_hex(w.reshape(-1) & 0x1F, 5)}"> : tensor<2x3xi5>')         # i5: one byte per element
    assert t.dims == (2, 3) and np.array_equal(got, np.where(w >= 16, w - 32, w))
    got, _ = parse_attr(f'dense<"{_hex([300, -200, 1], 9)}"> : tensor<3xi9>')                    # i9: two bytes
    assert got.tolist() == [-212, -200, 1]                                                         # 300 wraps in 9 bits
    bits = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1])
    raw = np.packbits(bits, bitorder="little").tobytes().hex()
    got, _ = parse_attr(f'dense<"0x{raw}"> : tensor<10xi1>')
    assert got.tolist() == bits.tolist()
    got, _ = parse_attr('dense<"0x2A00000000000000"> : tensor<70xi64>')                          # one element = splat
    assert got.shape == (70,) and set(got.tolist()) == {42}
    with pytest.raises(MlirSyntaxError):
        parse_attr('dense<"0x0102"> : tensor<3xi64>')
    with pytest.raises(MlirSyntaxError):
        parse_attr('dense<"0xZZ"> : tensor<1xi8>')


def test_module_with_hex_table_parses_like_the_decimal_one():
    table = (np.arange(128) * 7) % 128
    body = '''module {
  func.func @main(%arg0: tensor<4x!FHE.eint<7>>) -> tensor<4x!FHE.eint<7>> {
    %cst = arith.constant dense<TABLE> : tensor<128xi64>
    %0 = "FHELinalg.apply_lookup_table"(%arg0, %cst) : (tensor<4x!FHE.eint<7>>, tensor<128xi64>) -> tensor<4x!FHE.eint<7>>
    return %0 : tensor<4x!FHE.eint<7>>
  }
}'''
    p_dec = parse_module(body.replace("TABLE", str(table.tolist())))
    p_hex = parse_module(body.replace("TABLE", '"' + _hex(table, 64) + '"'))
    c_dec = [op for op in p_dec.ops if op.name == "arith.constant"][0].attrs["value"]
    c_hex = [op for op in p_hex.ops if op.name == "arith.constant"][0].attrs["value"]
    assert np.array_equal(c_dec, c_hex) and [o.name for o in p_dec.ops] == [o.name for o in p_hex.ops]
