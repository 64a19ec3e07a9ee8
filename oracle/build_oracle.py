"""Compile the plain-C part of the oracle (gcc, -O2, no fast-math) into oracle/libcbn_oracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcbn_oracle.so")


def build_oracle() -> str:
    src = os.path.join(HERE, "count_oracle.c")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, src], check=True)
    return LIB


def load_oracle():
    import ctypes as C

    lib = C.CDLL(build_oracle())
    P = C.c_void_p
    lib.oracle_count_families.argtypes = [P, C.c_int64, C.c_int64, P, P, P, C.c_int32, P, C.c_int32, P]
    lib.oracle_count_families.restype = None
    lib.oracle_cpt_from_counts.argtypes = [P, C.c_int64, C.c_int32, C.c_int64, P, P]
    lib.oracle_cpt_from_counts.restype = None
    return lib


def count_families(codes, n, fam_vars, cards):
    """codes uint8 [n_cols, ld] numpy; fam_vars: list of var-id lists (node last).  Returns list of int64 tables."""
    import numpy as np

    lib = load_oracle()
    mv = max(len(f) for f in fam_vars)
    nv = np.array([len(f) for f in fam_vars], dtype=np.int32)
    vs = np.zeros((len(fam_vars), mv), dtype=np.int32)
    cs = np.ones((len(fam_vars), mv), dtype=np.int32)
    sizes = []
    for i, f in enumerate(fam_vars):
        vs[i, : len(f)] = f
        cs[i, : len(f)] = [cards[v] for v in f]
        sizes.append(int(np.prod([cards[v] for v in f])))
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    counts = np.zeros(int(off[-1]), dtype=np.int64)
    codes = np.ascontiguousarray(codes)
    lib.oracle_count_families(codes.ctypes.data, codes.strides[0], int(n), nv.ctypes.data, vs.ctypes.data, cs.ctypes.data,
                              mv, off.ctypes.data, len(fam_vars), counts.ctypes.data)
    return [counts[off[i]: off[i + 1]].reshape([cards[v] for v in f]) for i, f in enumerate(fam_vars)]


if __name__ == "__main__":
    print(build_oracle())
