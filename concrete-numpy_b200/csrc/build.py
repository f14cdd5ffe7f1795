"""Builds the C-ABI shared library (sm_100a only) in-tree: concrete-numpy_b200/libfhe_b200.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libfhe_b200.so")
SOURCES = ["api.cu", "keys.cu", "keyswitch.cu", "keyswitch_tc.cu", "linear.cu", "pbs.cu", "pbs_inst_a.cu", "pbs_inst_b.cu", "pbs_inst_c.cu", "pbs_inst_d.cu", "pbs_inst_e.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SOURCES + ["common.cuh", "pbs_kernels.cuh", "pbs_pipe.cuh", "tmem.cuh", "build.py"]
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{out.decode()}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
