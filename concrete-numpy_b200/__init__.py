"""
B200-native encrypted-execution backend for the hot path behind concrete-numpy's
`Circuit.run` / `Server.run` (reference: concrete/numpy/compilation/server.py:270-286).

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/fhe_b200.h)
  _lib.py          ctypes binding of libfhe_b200.so (fails loudly when the library is missing)
  params.py        crypto-parameter selection + noise model (stands in for concrete-optimizer)
  engine.py        device-resident key sets, encrypt / decrypt, keyswitch / bootstrap / linear-op launches
  mlir_text/       FHE / FHELinalg MLIR text: builder shim (`mlir`, `concrete.lang`) and parser
  wide.py          9..16-bit values: CRT residues, bit extraction, circuit bootstrapping, vertical packing
  executor.py      walks the parsed program and launches the kernels
  compiler.py      `concrete.compiler`-compatible classes (JITSupport, ClientSupport, ...)
  api.py           Circuit / Client / Server mirror of the reference's compilation layer
"""

__version__ = "0.1.0"
