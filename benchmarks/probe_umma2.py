#!/usr/bin/env python
"""Layout discovery for MN-major tf32 operands (runs on a B200).

The operand under test is a raw shared-memory image whose float at byte offset 4*i holds the value i (exact in tf32
for i < 2048).  Multiplying it with a known-good K-major one-hot operand makes the accumulator spell out WHICH smem
element the tensor core reads as (row, k): D[m][n] = index of B(n, k = m)  resp.  index of A(m, k = n).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glint_word2vec_b200 import _C  # noqa: E402
from benchmarks.probe_umma import sw128, desc, idesc, opword, run  # noqa: E402

np.set_printoptions(linewidth=250, threshold=100000)


def desc_raw(start, lbo, sbo, layout=2, base_offset=0, lbo_mode=0):
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (base_offset << 49) | (lbo_mode << 52) | (layout << 61)


def main():
    # ---- B under test (MN-major), A = one-hot K-major: A[m][k] = 1 if m == k (m < 8)
    A = np.zeros((128, 32), np.float32)
    for k in range(8):
        A[k, k] = 1
    nB = 2048
    Bimg = np.arange(nB, dtype=np.float32).view(np.uint8)            # 8 KB: values = float index
    image = np.concatenate([sw128(A), Bimg])
    boff = 16384
    for name, lbo, sbo, layout in (("sw128 lbo16 sbo1024", 16, 1024, 2), ("sw128 lbo1024 sbo1024", 1024, 1024, 2),
                                   ("sw128 lbo4096 sbo128", 4096, 128, 2), ("none lbo128 sbo1024", 128, 1024, 0),
                                   ("none lbo1024 sbo128", 1024, 128, 0)):
        for N in (32, 64):
            ops = [[desc(0, 16, 1024), desc_raw(boff, lbo, sbo, layout), opword(idesc(128, N, 0, 1), 0)]]
            got = run(image, ops, N)
            print(f"B MN-major {name} N={N}: D[k][n] = float index of B(n,k); rows k=0..7")
            print(got[:8].astype(np.int64))

    # ---- A under test (MN-major), B = one-hot K-major [16 x 32]: B[n][k] = 1 if n == k (n < 8)
    B = np.zeros((16, 32), np.float32)
    for k in range(8):
        B[k, k] = 1
    Aimg = np.arange(nB, dtype=np.float32).view(np.uint8)
    image = np.concatenate([sw128(B), Aimg])
    aoff = 16 * 128
    for name, lbo, sbo, layout in (("sw128 lbo1024 sbo1024", 1024, 1024, 2), ("sw128 lbo2048 sbo1024", 2048, 1024, 2),
                                   ("sw128 lbo1024 sbo4096", 1024, 4096, 2), ("none lbo128 sbo1024", 128, 1024, 0)):
        ops = [[desc_raw(aoff, lbo, sbo, layout), desc(0, 16, 1024), opword(idesc(128, 16, 1, 0), 0)]]
        got = run(image, ops, 16)
        print(f"A MN-major {name}: D[m][n<8] = float index of A(m,k=n); all 128 rows, 8 cols")
        print(got[:, :8].astype(np.int64).T)
    return 0


if __name__ == "__main__":
    sys.exit(main())
