"""
Generates tests/golden/*.json(.gz) by running the REFERENCE's own frontend (imported from
/root/reference, see ref_frontend.py): trace -> fuse -> bounds -> GraphConverter passes -> MLIR text,
plus the reference's clear graph evaluation (`Graph.evaluate`, representation/graph.py:68-188) of
seeded samples as the expected outputs.  Run in the development container:

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py cfg1 ops   # a subset (prefix match)

The circuits restate the reference's behavioural tests (tests/execution/*.py, one or two
parameterisations per operator family) and the five BASELINE.json workloads.
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.path and os.path.abspath(sys.path[0]) == os.path.dirname(os.path.abspath(__file__)):
    sys.path.pop(0)      # `oracle` must resolve to the package, not oracle/oracle.py
sys.path.insert(0, ROOT)
from oracle import ref_frontend  # noqa: E402

cnp, connx = ref_frontend.load()
OUT = os.path.join(ROOT, "tests", "golden")
GLOBAL_P_ERROR = 1e-4     # the reference's own test configuration (tests/conftest.py:119)


def configuration():
    return cnp.Configuration(jit=True, enable_unsafe_features=True, use_insecure_key_cache=True,
                             insecure_key_cache_location="/tmp/cnp_golden_keycache",
                             global_p_error=GLOBAL_P_ERROR, dataflow_parallelize=False)


def _jsonable(v):
    if isinstance(v, tuple):
        return {"tuple": [_jsonable(t) for t in v]}
    a = np.asarray(v)
    return {"shape": list(a.shape), "data": [int(t) for t in a.reshape(-1)]}


def emit(name, function, statuses, inputset, samples):
    compiler = cnp.Compiler(function, statuses)
    circuit = compiler.compile(inputset, configuration())
    specs = circuit.client.specs
    cases = []
    for sample in samples:
        sample = sample if isinstance(sample, tuple) else (sample,)
        expected = circuit.graph(*sample)
        cases.append({"args": [_jsonable(a) for a in sample], "expected": _jsonable(expected)})
    doc = {
        "name": name,
        "mlir": circuit.mlir,
        "input_signs": [bool(s) for s in specs.input_signs],
        "output_signs": [bool(s) for s in specs.output_signs],
        "global_p_error": GLOBAL_P_ERROR,
        "graph": circuit.graph.format() if len(circuit.graph.graph) < 60 else None,
        "cases": cases,
    }
    text = json.dumps(doc)
    path = os.path.join(OUT, name + ".json")
    if len(text) > 200_000:
        with gzip.open(path + ".gz", "wt", encoding="utf-8") as f:
            f.write(text)
        path += ".gz"
    else:
        with open(path, "w", encoding="utf-8") as f:
            f.write(text)
    fb = circuit.server._compilation_feedback
    print(f"{name:36s} p={fb.n_pbs:6d} pbs depth={fb.tlu_depth:3d} norm2={fb.norm2:8.1f} "
          f"{os.path.getsize(path) / 1024:8.1f} KiB")


# ------------------------------------------------------------ BASELINE configs ----

def cfg1():
    rng = np.random.default_rng(1)
    table = cnp.LookupTable(rng.integers(0, 16, 16).tolist())
    data = np.random.default_rng(0)
    inputset = [data.integers(0, 16, 1024) for _ in range(8)] + [np.zeros(1024, dtype=np.int64), np.full(1024, 15)]
    emit("cfg1_lut4_1024", lambda x: table[x], {"x": "encrypted"}, inputset,
         [data.integers(0, 16, 1024) for _ in range(2)])


def cfg2():
    data = np.random.default_rng(0)
    shape = (64, 64)
    inputset = [data.integers(0, 13, shape) for _ in range(6)] + [np.zeros(shape, dtype=np.int64), np.full(shape, 12)]
    emit("cfg2_univariate8_64x64", lambda x: (x + 3) ** 2 // 5, {"x": "encrypted"}, inputset,
         [data.integers(0, 13, shape) for _ in range(2)])


def ternary_weights(rng, shape, nnz, axis_len):
    w = np.zeros(shape, dtype=np.int64).reshape(axis_len, -1)
    for col in range(w.shape[1]):
        rows = rng.choice(axis_len, nnz, replace=False)
        w[rows, col] = rng.choice([-1, 1], nnz)
    return w.reshape(shape)


def cfg3():
    wr = np.random.default_rng(1)
    W = ternary_weights(wr, (128, 64), 10, 128)
    data = np.random.default_rng(0)
    inputset = [data.integers(0, 4, (64, 128)) for _ in range(20)] + [np.zeros((64, 128), dtype=np.int64), np.full((64, 128), 3)]
    emit("cfg3_matmul6_relu", lambda x: np.maximum(x @ W, 0), {"x": "encrypted"}, inputset,
         [data.integers(0, 4, (64, 128)) for _ in range(2)])


def cfg4():
    wr = np.random.default_rng(1)
    W = wr.choice([-1, 0, 1], size=(8, 8, 3, 3), p=[0.25, 0.5, 0.25]).astype(np.int64)
    data = np.random.default_rng(0)
    shape = (1, 8, 28, 28)
    inputset = [data.integers(0, 2, shape) for _ in range(20)] + [np.zeros(shape, dtype=np.int64), np.ones(shape, dtype=np.int64)]
    emit("cfg4_conv2d_relu", lambda x: np.maximum(connx.conv(x, W, pads=(1, 1, 1, 1)), 0) // 4,
         {"x": "encrypted"}, inputset, [data.integers(0, 2, shape) for _ in range(2)])


def cfg5(entries=128):
    """Key-value database circuits of the reference's own example at `entries` rows: the example
    source is executed from /root/reference (up to its demo section) with NUMBER_OF_ENTRIES patched."""
    path = os.path.join(ref_frontend.REFERENCE, "examples", "key-value-database", "static-size.py")
    src = open(path, "r", encoding="utf-8").read().split("class KeyValueDatabase")[0]
    src = src.replace("NUMBER_OF_ENTRIES = 5", f"NUMBER_OF_ENTRIES = {entries}")
    ns = {}
    exec(compile(src, path, "exec"), ns)  # noqa: S102  (the reference's own example code)
    shape, kc, vc = ns["STATE_SHAPE"], ns["NUMBER_OF_KEY_CHUNKS"], ns["NUMBER_OF_VALUE_CHUNKS"]
    full = 2 ** ns["CHUNK_SIZE"] - 1
    in2 = [(np.zeros(shape, dtype=np.int64), np.full(kc, full))]
    in3 = [(np.zeros(shape, dtype=np.int64), np.full(kc, full), np.full(vc, full))]
    rng = np.random.default_rng(0)
    # a plausible database state: first 40 rows used, distinct keys
    state = np.zeros(shape, dtype=np.int64)
    used = 40 if entries >= 64 else max(entries // 2, 1)
    state[:used, 0] = 1
    state[:used, 1:] = rng.integers(0, 16, (used, kc + vc))
    key_hit = state[min(17, used - 1), 1:1 + kc].copy()
    key_miss = np.full(kc, 15)
    value = rng.integers(0, 16, vc)
    enc3 = {"state": "encrypted", "key": "encrypted", "value": "encrypted"}
    emit(f"cfg5_kvdb{entries}_query", ns["_query_impl"], {"state": "encrypted", "key": "encrypted"}, in2,
         [(state, key_hit), (state, key_miss)])
    emit(f"cfg5_kvdb{entries}_replace", ns["_replace_impl"], enc3, in3, [(state, key_hit, value)])
    emit(f"cfg5_kvdb{entries}_insert", ns["_insert_impl"], enc3, in3, [(state, key_miss, value)])


# ------------------------------------------------------ operator coverage ---------

def _rand(rng, lo, hi, shape):
    return rng.integers(lo, hi + 1, shape) if shape != () else int(rng.integers(lo, hi + 1))


def ops():
    rng = np.random.default_rng(7)

    def case(name, fn, params, n_samples=3):
        """params: {name: (status, lo, hi, shape)}"""
        names = list(params)
        statuses = {k: v[0] for k, v in params.items()}

        def sample(mode=None):
            vals = []
            for k in names:
                _, lo, hi, shape = params[k]
                if mode == "min":
                    v = np.full(shape, lo) if shape != () else lo
                elif mode == "max":
                    v = np.full(shape, hi) if shape != () else hi
                else:
                    v = _rand(rng, lo, hi, shape)
                vals.append(v)
            return tuple(vals) if len(vals) > 1 else vals[0]

        inputset = [sample() for _ in range(40)] + [sample("min"), sample("max")]
        samples = [sample() for _ in range(n_samples)]
        only = os.environ.get("GOLDEN_ONLY")
        if only is None or only in name:
            emit("ops_" + name, fn, statuses, inputset, samples)

    E, C = "encrypted", "clear"
    # add / sub / mul / neg (tests/execution/test_add.py, test_sub.py, test_mul.py, test_neg.py)
    case("add_const", lambda x: x + 42, {"x": (E, 0, 85, ())})
    case("add_tensor_const", lambda x: x + np.array([[1, 2, 3], [4, 5, 6]]), {"x": (E, 0, 20, (2, 3))})
    case("add_enc_enc", lambda x, y: x + y, {"x": (E, 0, 60, ()), "y": (E, 0, 60, ())})
    case("add_enc_clear_arg", lambda x, y: x + y, {"x": (E, 0, 40, (3,)), "y": (C, 0, 20, (3,))})
    case("add_broadcast", lambda x, y: x + y, {"x": (E, 0, 30, (3, 2)), "y": (E, 0, 30, (2,))})
    case("sub_const_left", lambda x: 100 - x, {"x": (E, 0, 60, (3,))})
    case("sub_const_right", lambda x: x - 7, {"x": (E, 7, 60, ())})
    case("sub_enc_enc", lambda x, y: x - y, {"x": (E, 30, 60, (2, 2)), "y": (E, 0, 30, (2, 2))})
    case("sub_signed_result", lambda x, y: x - y, {"x": (E, 0, 15, (4,)), "y": (E, 0, 15, (4,))})
    case("mul_const", lambda x: x * 3, {"x": (E, 0, 20, (2, 3))})
    case("mul_tensor_const", lambda x: x * np.array([1, -2, 3]), {"x": (E, 0, 10, (3,))})
    case("mul_clear_arg", lambda x, y: x * y, {"x": (E, 0, 7, ()), "y": (C, 0, 7, ())})
    case("neg", lambda x: -x, {"x": (E, 0, 20, (3,))})
    case("signed_input", lambda x: x + 5, {"x": (E, -8, 7, (4,))})
    case("signed_input_tlu", lambda x: np.abs(x) * 2, {"x": (E, -7, 7, (2, 2))})
    # dot / matmul (test_dot.py, test_matmul.py)
    w5 = rng.integers(-3, 4, 5)
    case("dot", lambda x: np.dot(x, w5), {"x": (E, 0, 7, (5,))})
    case("dot_left_const", lambda x: np.dot(w5, x), {"x": (E, 0, 7, (5,))})
    w32 = rng.integers(0, 4, (3, 2))
    case("matmul_2d", lambda x: x @ w32, {"x": (E, 0, 7, (4, 3))})
    w23 = rng.integers(0, 4, (2, 3))
    case("matmul_const_left", lambda x: w23 @ x, {"x": (E, 0, 7, (3, 4))})
    case("matmul_vec_mat", lambda x: x @ w32, {"x": (E, 0, 7, (3,))})
    case("matmul_mat_vec", lambda x: w23 @ x, {"x": (E, 0, 7, (3,))})
    wb = rng.integers(0, 3, (2, 3, 2))
    case("matmul_batched", lambda x: x @ wb, {"x": (E, 0, 5, (2, 4, 3))})
    case("matmul_batched_bcast", lambda x: x @ w32, {"x": (E, 0, 5, (2, 2, 4, 3))})
    wk = rng.integers(-2, 3, (40, 20))
    case("matmul_wide", lambda x: x @ wk, {"x": (E, 0, 1, (5, 40))}, n_samples=2)
    # conv2d (test_convolution.py)
    wc = rng.integers(0, 3, (2, 3, 2, 2))
    case("conv2d", lambda x: connx.conv(x, wc), {"x": (E, 0, 3, (1, 3, 4, 4))})
    bc = rng.integers(0, 4, 2)
    case("conv2d_bias_stride", lambda x: connx.conv(x, wc, bc, strides=(2, 2)), {"x": (E, 0, 3, (1, 3, 6, 6))})
    wg = rng.integers(0, 3, (4, 2, 2, 2))
    case("conv2d_group_pad", lambda x: connx.conv(x, wg, pads=(1, 1, 1, 1), group=2), {"x": (E, 0, 2, (1, 4, 3, 3))})
    case("conv2d_dilation", lambda x: connx.conv(x, wc, dilations=(2, 2)), {"x": (E, 0, 3, (1, 3, 5, 5))})
    # direct table lookups (test_direct_table_lookup.py)
    t3 = cnp.LookupTable([3, 1, 2, 0, 7, 5, 6, 4])
    case("tlu_scalar", lambda x: t3[x], {"x": (E, 0, 7, ())})
    case("tlu_tensor", lambda x: t3[x], {"x": (E, 0, 7, (3, 2))})
    t_neg = cnp.LookupTable([-1, 2, -3, 4])
    case("tlu_negative_outputs", lambda x: t_neg[x], {"x": (E, 0, 3, (3,))})
    ts = cnp.LookupTable([t3, cnp.LookupTable([0, 1, 2, 3, 4, 5, 6, 7]), t3])
    case("tlu_multi_table", lambda x: ts[x], {"x": (E, 0, 7, (3,))})
    tsg = cnp.LookupTable([5, 6, 7, 8, 1, 2, 3, 4])
    case("tlu_signed_index", lambda x: tsg[x], {"x": (E, -4, 3, (2,))})
    # univariate / fused (test_others.py)
    case("uni_square", lambda x: x ** 2, {"x": (E, 0, 11, (2, 2))})
    case("uni_sqrt_fused", lambda x: np.floor(np.sqrt(x + 1) * 3).astype(np.int64), {"x": (E, 0, 50, (3,))})
    case("uni_sin_fused", lambda x: np.round(np.sin(x) * 10).astype(np.int64) + 10, {"x": (E, 0, 31, ())})
    case("uni_where", lambda x: np.where(x < 5, x * 3, x), {"x": (E, 0, 15, (4,))})
    case("uni_minimum_const", lambda x: np.minimum(x, 9), {"x": (E, 0, 31, (2, 3))})
    case("uni_chain", lambda x: ((x + 1) ** 2 // 3) % 7, {"x": (E, 0, 14, (5,))})
    case("uni_multi_input_fused", lambda x: np.round(np.sqrt(x) * np.sin(x)).astype(np.int64) + 6, {"x": (E, 0, 31, (3,))})
    case("round_bit_pattern", lambda x: cnp.round_bit_pattern(x, lsbs_to_remove=2) // 4, {"x": (E, 0, 59, (4,))})
    case("astype_via_univariate", cnp.univariate(lambda v: (v * 1.5).astype(np.int64)) if hasattr(cnp, "univariate") else (lambda x: x), {"x": (E, 0, 20, (3,))})
    # comparisons / equality (test_comparison.py)
    case("less_const", lambda x: x < 7, {"x": (E, 0, 15, (4,))})
    case("less_enc_small", lambda x, y: x < y, {"x": (E, 0, 3, (3,)), "y": (E, 0, 3, (3,))})
    case("less_enc_chunked", lambda x, y: x < y, {"x": (E, 0, 63, (3,)), "y": (E, 0, 63, (3,))})
    case("greater_equal_enc", lambda x, y: x >= y, {"x": (E, 0, 31, ()), "y": (E, 0, 31, ())})
    case("less_equal_signed_mixed", lambda x, y: x <= y, {"x": (E, -16, 15, (2,)), "y": (E, 0, 31, (2,))})
    case("equal_enc", lambda x, y: x == y, {"x": (E, 0, 15, (5,)), "y": (E, 0, 15, (5,))})
    case("not_equal_enc_chunked", lambda x, y: x != y, {"x": (E, 0, 63, (4,)), "y": (E, 0, 63, (4,))})
    # bitwise / shifts (test_bitwise.py, test_shift.py)
    case("bitwise_and_enc", lambda x, y: x & y, {"x": (E, 0, 15, (3,)), "y": (E, 0, 15, (3,))})
    case("bitwise_or_enc", lambda x, y: x | y, {"x": (E, 0, 31, ()), "y": (E, 0, 31, ())})
    case("bitwise_xor_enc", lambda x, y: x ^ y, {"x": (E, 0, 7, (2, 2)), "y": (E, 0, 7, (2, 2))})
    case("bitwise_and_const", lambda x: x & 5, {"x": (E, 0, 15, (3,))})
    case("left_shift_enc", lambda x, y: x << y, {"x": (E, 0, 3, (3,)), "y": (E, 0, 3, (3,))})
    case("right_shift_enc", lambda x, y: x >> y, {"x": (E, 0, 31, (3,)), "y": (E, 0, 3, (3,))})
    case("right_shift_const", lambda x: x >> 2, {"x": (E, 0, 63, ())})
    # reductions / shape ops (test_sum.py, test_concat.py, test_reshape.py, test_transpose.py, ...)
    case("sum_all", lambda x: np.sum(x), {"x": (E, 0, 5, (3, 4))})
    case("sum_axis0", lambda x: np.sum(x, axis=0), {"x": (E, 0, 5, (3, 4))})
    case("sum_axis_neg_keepdims", lambda x: np.sum(x, axis=-1, keepdims=True), {"x": (E, 0, 5, (2, 3, 4))})
    case("sum_axes_tuple", lambda x: np.sum(x, axis=(0, 2)), {"x": (E, 0, 3, (2, 3, 4))})
    case("concat_axis0", lambda x, y: np.concatenate((x, y)), {"x": (E, 0, 9, (2, 3)), "y": (E, 0, 9, (4, 3))})
    case("concat_axis1", lambda x, y: np.concatenate((x, y), axis=1), {"x": (E, 0, 9, (2, 3)), "y": (E, 0, 9, (2, 1))})
    case("concat_axis_none", lambda x, y: np.concatenate((x, y), axis=None), {"x": (E, 0, 9, (2, 3)), "y": (E, 0, 9, (4,))})
    case("reshape_collapse", lambda x: x.reshape((6,)) + 1, {"x": (E, 0, 9, (2, 3))})
    case("reshape_expand", lambda x: x.reshape((2, 3, 2)) + 1, {"x": (E, 0, 9, (6, 2))})
    case("reshape_general", lambda x: x.reshape((3, 4)) + 1, {"x": (E, 0, 9, (2, 6))})
    case("flatten", lambda x: x.flatten() * 2, {"x": (E, 0, 9, (2, 2, 2))})
    case("transpose", lambda x: np.transpose(x) + 1, {"x": (E, 0, 9, (2, 3))})
    case("transpose_axes", lambda x: np.transpose(x, (1, 2, 0)) + 1, {"x": (E, 0, 9, (2, 3, 4))})
    case("expand_dims_squeeze", lambda x: np.squeeze(np.expand_dims(x, axis=1)) + 1, {"x": (E, 0, 9, (3,))})
    case("broadcast_to", lambda x: np.broadcast_to(x, (2, 3)) + 1, {"x": (E, 0, 9, (3,))})
    # static indexing / assignment (test_static_indexing.py, test_static_assignment.py)
    case("index_scalar", lambda x: x[1, 2] + 1, {"x": (E, 0, 9, (2, 3))})
    case("index_negative", lambda x: x[-1] + 1, {"x": (E, 0, 9, (4,))})
    case("index_slice", lambda x: x[1:3] + 1, {"x": (E, 0, 9, (5,))})
    case("index_reverse", lambda x: x[::-1] + 1, {"x": (E, 0, 9, (5,))})
    case("index_mixed", lambda x: x[1, ::2] + 1, {"x": (E, 0, 9, (3, 5))})
    case("index_column", lambda x: x[:, 1] + 1, {"x": (E, 0, 9, (3, 4))})

    def assign_scalar(x):
        x[0] = 3
        return x

    def assign_slice(x, y):
        x[1:3] = y
        return x

    def assign_column(x):
        x[:, 1] = np.array([1, 2, 3])
        return x

    case("assign_scalar_clear", assign_scalar, {"x": (E, 0, 9, (4,))})
    case("assign_slice_enc", assign_slice, {"x": (E, 0, 9, (5,)), "y": (E, 0, 9, (2,))})
    case("assign_column_clear", assign_column, {"x": (E, 0, 9, (3, 3))})
    # array / zeros / ones (test_array.py, test_zeros.py, test_ones.py)
    case("array_from_scalars", lambda x, y: cnp.array([x, y, 3]) + 1, {"x": (E, 0, 9, ()), "y": (E, 0, 9, ())})
    case("array_2d", lambda x, y: cnp.array([[x, y], [y, x]]), {"x": (E, 0, 9, ()), "y": (E, 0, 9, ())})
    case("zeros_plus", lambda x: cnp.zeros((2, 2)) + x, {"x": (E, 0, 9, (2, 2))})
    case("ones_plus", lambda x: cnp.ones((3,)) + x, {"x": (E, 0, 9, (3,))})
    case("zero_one_scalar", lambda x: cnp.zero() + cnp.one() + x, {"x": (E, 0, 9, ())})
    case("two_outputs_style", lambda x, y: (x + y) * 2 - y, {"x": (E, 0, 9, (2,)), "y": (E, 0, 9, (2,))})
    case("unused_arg", lambda x, y: x + 10, {"x": (E, 0, 10, ()), "y": (E, 0, 1, ())})

    # iteration over an encrypted tensor (test_iter.py)
    def iter_sum(x):
        result = cnp.zeros(x.shape[1:])
        for value in x:
            result += value
        return result

    case("iter_sum", iter_sum, {"x": (E, 0, 3, (3, 2, 4))})

    # more univariate / constant-operand cases of test_others.py:125-470
    case("others_const_floordiv", lambda x: 127 // x, {"x": (E, 1, 127, (3,))})
    case("others_pow2", lambda x: 2 ** x, {"x": (E, 0, 6, (2, 2))})
    case("others_invert_and", lambda x: (~x) & 10, {"x": (E, 0, 15, (4,))})
    case("others_clip", lambda x: x.clip(5, 10), {"x": (E, 0, 15, (3, 2))})
    case("others_maximum_tensor", lambda x: np.maximum(x, [[10, 20], [30, 40], [50, 60]]), {"x": (E, 0, 63, (3, 2))})
    case("others_shape_attrs", lambda x: x + x.shape[0] + x.ndim + x.size, {"x": (E, 0, 15, (3, 2))})
    case("others_ones_zeros_like", lambda x: x + np.ones_like(x) + np.zeros_like(x), {"x": (E, 0, 15, (2, 3))})
    table32 = cnp.LookupTable(list(range(32)))
    case("others_tlu_of_sum", lambda x, y: table32[x + y], {"x": (E, 0, 15, (3,)), "y": (E, 0, 15, (3,))})
    case("others_cube", lambda x: x ** 3, {"x": (E, 0, 4, ())})


def known_answers():
    """Known-answer circuits of the reference (tests/compilation/test_circuit.py:132-192, 307-345;
    docs/tutorial/table_lookups.md:39-42)."""
    emit("ka_client_server", lambda x: x + 42, {"x": "encrypted"}, [np.random.default_rng(3).integers(0, 2 ** 4, 3) for _ in range(98)] + [np.zeros(3, dtype=np.int64), np.full(3, 15)],
         [np.array([3, 8, 1])])
    table = cnp.LookupTable([2, -1, 3, 0])
    emit("ka_doc_table", lambda x: table[x], {"x": "encrypted"}, range(4), [0, 1, 2, 3])


def wide():
    """Widths 9..16 (reference: MAXIMUM_TLU_BIT_WIDTH = 16, concrete/numpy/mlir/utils.py:16; the compiler switches
    to a CRT encoding there, client.py:213-235).  Shapes follow tests/execution/test_add.py / test_mul.py /
    test_matmul.py / test_others.py with larger value ranges."""
    rng = np.random.default_rng(7)

    def ins(lo, hi, shape, n=12):
        return [rng.integers(lo, hi + 1, shape) for _ in range(n)] + [np.full(shape, lo, dtype=np.int64), np.full(shape, hi, dtype=np.int64)]

    def pairs(a, b):
        return list(zip(a, b))

    emit("wide_square_div_10bit", lambda x: (x ** 2) // 3, {"x": "encrypted"}, ins(0, 31, (4,)),
         [np.array([0, 31, 17, 5]), rng.integers(0, 32, (4,))])
    emit("wide_linear_tlu_16bit", lambda x, y: (x * 3 + y) % 1000, {"x": "encrypted", "y": "encrypted"},
         pairs(ins(0, 15000, (3,)), ins(0, 15000, (3,))),
         [(np.array([15000, 0, 1234]), np.array([15000, 7, 4321])), (rng.integers(0, 15001, (3,)), rng.integers(0, 15001, (3,)))])
    emit("wide_signed_abs_12bit", lambda x, y: np.abs(x - y) + 1, {"x": "encrypted", "y": "encrypted"},
         pairs(ins(0, 2000, (3,)), ins(0, 2000, (3,))),
         [(np.array([0, 2000, 1000]), np.array([2000, 0, 999])), (rng.integers(0, 2001, (3,)), rng.integers(0, 2001, (3,)))])
    W = np.array([[100, -7], [3, 250], [-41, 9]], dtype=np.int64)
    emit("wide_matmul_relu_14bit", lambda x: np.maximum(x @ W, 0) // 2, {"x": "encrypted"}, ins(0, 20, (2, 3)),
         [np.array([[20, 20, 0], [0, 20, 20]]), rng.integers(0, 21, (2, 3))])
    tables = cnp.LookupTable([cnp.LookupTable(((np.arange(512) * 5) % 301).tolist()), cnp.LookupTable((np.arange(512) // 2).tolist())])
    emit("wide_multi_table_9bit", lambda x: tables[x], {"x": "encrypted"}, ins(0, 511, (2,)),
         [np.array([511, 0]), rng.integers(0, 512, (2,))])
    emit("wide_signed_output_11bit", lambda x: x * 7 - 3000, {"x": "encrypted"}, ins(0, 600, (3,)),
         [np.array([0, 600, 429]), rng.integers(0, 601, (3,))])


def cfg3b():
    """BASELINE config 3, variant (B) of SURVEY 8d: dense 6-bit x, ternary W -> 13-bit accumulators, i.e. the circuit
    the reference's compiler would run on its CRT / without-padding path."""
    wr = np.random.default_rng(1)
    W = ternary_weights(wr, (128, 64), 40, 128)
    data = np.random.default_rng(0)
    inputset = [data.integers(0, 64, (64, 128)) for _ in range(20)] + [np.zeros((64, 128), dtype=np.int64), np.full((64, 128), 63)]
    emit("cfg3b_matmul6_dense_relu", lambda x: np.maximum(x @ W, 0), {"x": "encrypted"}, inputset,
         [data.integers(0, 64, (64, 128))])


def linear_wide():
    """Circuits without table lookups at widths beyond 16 bits (graph_converter.py:288-306 only limits circuits WITH
    lookups); value ranges of tests/execution/test_add.py:62-68 ([0, 1000000])."""
    rng = np.random.default_rng(11)

    def ins(lo, hi, shape, n=10):
        return [rng.integers(lo, hi + 1, shape) for _ in range(n)] + [np.full(shape, lo, dtype=np.int64), np.full(shape, hi, dtype=np.int64)]

    emit("lin_add_const_20bit", lambda x: x + 42, {"x": "encrypted"}, ins(0, 1000000, (3,)),
         [np.array([0, 1000000, 123456]), rng.integers(0, 1000001, (3,))])
    emit("lin_mul_sub_31bit", lambda x, y: x * 1000 - y, {"x": "encrypted", "y": "encrypted"},
         list(zip(ins(0, 1000000, (2, 2)), ins(0, 1000000, (2, 2))))
         + [(np.zeros((2, 2), dtype=np.int64), np.full((2, 2), 1000000)), (np.full((2, 2), 1000000), np.zeros((2, 2), dtype=np.int64))],
         [(np.array([[0, 1000000], [1, 999999]]), np.array([[1000000, 0], [1000000, 7]])),
          (rng.integers(0, 1000001, (2, 2)), rng.integers(0, 1000001, (2, 2)))])
    W = np.array([[5000, -7], [3, 2500], [-4100, 9]], dtype=np.int64)
    corner = [np.array([[1000, 1000, 0], [0, 0, 1000]]), np.array([[0, 1000, 0], [1000, 0, 1000]])]   # extreme sums
    emit("lin_matmul_24bit", lambda x: x @ W, {"x": "encrypted"}, ins(0, 1000, (2, 3)) + corner,
         [corner[0], rng.integers(0, 1001, (2, 3))])
    emit("lin_sum_neg_31bit", lambda x: -np.sum(x * 1003), {"x": "encrypted"}, ins(0, 100000, (8,)),
         [np.full((8,), 100000), rng.integers(0, 100001, (8,))])


def sweep():
    """Full parameterisations of the reference's operator tests, read from the reference itself: the test modules under
    /root/reference/tests/execution are imported and the argument lists of their `pytest.mark.parametrize` decorators are
    iterated, so the functions, shapes, ranges and encryption statuses below are the reference's own objects (nothing is
    restated here).  Covers test_comparison.py, test_shift.py, test_bitwise.py, test_matmul.py, test_sum.py,
    test_static_indexing.py and test_static_assignment.py."""
    import importlib.util
    import itertools

    rng = np.random.default_rng(2024)
    np.random.seed(2024)                      # the assignment cases draw their constants at import time

    def load(name):
        path = os.path.join(ref_frontend.REFERENCE, "tests", "execution", name + ".py")
        spec = importlib.util.spec_from_file_location("ref_tests_" + name, path)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        return module

    def grid(test_function):
        """[{argname: value}] for the cartesian product of the test's parametrize marks"""
        axes = []
        for mark in reversed(test_function.pytestmark):
            if mark.name != "parametrize":
                continue
            names = [n.strip() for n in mark.args[0].split(",")]
            rows = []
            for item in mark.args[1]:
                if type(item).__name__ == "ParameterSet":        # pytest.param(...)
                    values = item.values
                else:
                    values = (item,) if len(names) == 1 else item
                rows.append(dict(zip(names, values)))
            axes.append(rows)
        for combo in itertools.product(*axes):
            merged = {}
            for part in combo:
                merged.update(part)
            yield merged

    def draw(spec, mode=None):
        lo, hi = spec["range"]
        shape = spec.get("shape", ())
        if mode == "min":
            return np.full(shape, lo) if shape != () else lo
        if mode == "max":
            return np.full(shape, hi) if shape != () else hi
        return _rand(rng, lo, hi, shape)

    def from_parameters(name, function, parameters, extra_samples=()):
        names = list(parameters)
        statuses = {k: parameters[k]["status"] for k in names}

        def sample(mode=None):
            vals = tuple(draw(parameters[k], mode) for k in names)
            return vals if len(vals) > 1 else vals[0]

        inputset = [sample() for _ in range(60)] + [sample("min"), sample("max")]
        emit(name, function, statuses, inputset, [sample(), sample()] + list(extra_samples))

    only = os.environ.get("GOLDEN_ONLY")

    def wanted(name):
        return only is None or only in name

    comparison = load("test_comparison")
    for i, row in enumerate(grid(comparison.test_comparison)):
        if wanted("cmp"):
            from_parameters(f"sw_cmp_{i:02d}", row["function"], row["parameters"])
    for i, row in enumerate(grid(comparison.test_optimized_comparison)):
        if wanted("cmpopt"):
            from_parameters(f"sw_cmpopt_{i:02d}", row["function"], row["parameters"])
    shift = load("test_shift")
    for i, row in enumerate(grid(shift.test_left_shift)):
        if wanted("shl"):
            extra = [(a, b) for a in range(2) for b in range(8)] if i == 0 else []     # test_left_shift_coverage
            from_parameters(f"sw_shl_{i:02d}", row["function"], row["parameters"], extra)
    for i, row in enumerate(grid(shift.test_right_shift)):
        if wanted("shr"):
            extra = [(0b11, 0), (0b11, 1), (0b110, 2), (0b1100, 3), (0b11000, 4), (0b110000, 5)] if i == 0 else []
            from_parameters(f"sw_shr_{i:02d}", row["function"], row["parameters"], extra)
    bitwise = load("test_bitwise")
    for i, row in enumerate(grid(bitwise.test_bitwise)):
        if wanted("bit"):
            from_parameters(f"sw_bit_{i:02d}", row["function"], row["parameters"])
    matmul = load("test_matmul")
    for i, row in enumerate(grid(matmul.test_matmul)):
        if not wanted("mm"):
            continue
        lhs_shape, rhs_shape, (lo, hi) = row["lhs_shape"], row["rhs_shape"], row["bounds"]
        lhs_cst, rhs_cst = rng.integers(lo, hi, lhs_shape), rng.integers(lo, hi, rhs_shape)
        emit(f"sw_mm_{i:02d}_enc_cst", lambda x: x @ rhs_cst, {"x": "encrypted"},
             [rng.integers(lo, hi, lhs_shape) for _ in range(40)] + [np.full(lhs_shape, hi - 1), np.full(lhs_shape, lo)],
             [rng.integers(lo, hi, lhs_shape)])
        emit(f"sw_mm_{i:02d}_cst_enc", lambda x: lhs_cst @ x, {"x": "encrypted"},
             [rng.integers(lo, hi, rhs_shape) for _ in range(40)] + [np.full(rhs_shape, hi - 1), np.full(rhs_shape, lo)],
             [rng.integers(lo, hi, rhs_shape)])
    summation = load("test_sum")
    for i, row in enumerate(grid(summation.test_sum)):
        if wanted("sum"):
            from_parameters(f"sw_sum_{i:02d}", row["function"], row["parameters"])
    indexing = load("test_static_indexing")
    for i, row in enumerate(grid(indexing.test_static_indexing)):
        if wanted("idx"):
            shape = row["shape"]
            emit(f"sw_idx_{i:02d}", row["function"], {"x": "encrypted"},
                 [rng.integers(0, 2 ** 5, shape) for _ in range(40)] + [np.zeros(shape, dtype=np.int64), np.full(shape, 31)],
                 [rng.integers(0, 2 ** 5, shape)])
    assignment = load("test_static_assignment")
    for i, row in enumerate(grid(assignment.test_static_assignment)):
        if wanted("asg"):
            shape = row["shape"]
            emit(f"sw_asg_{i:02d}", row["function"], {"x": "encrypted"},
                 [rng.integers(0, 2 ** 7, shape) for _ in range(40)] + [np.zeros(shape, dtype=np.int64), np.full(shape, 127)],
                 [rng.integers(0, 2 ** 7, shape)])


GROUPS = {"lin": linear_wide, "cfg3b": cfg3b, "wide": wide, "cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "ops": ops, "ka": known_answers,
          "cfg5small": lambda: cfg5(8), "sweep": sweep}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    wanted = sys.argv[1:] or [g for g in GROUPS if g not in ("cfg5small", "sweep")] + ["cfg5small", "sweep"]
    for g in wanted:
        GROUPS[g]()
