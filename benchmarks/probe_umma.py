#!/usr/bin/env python
"""Descriptor probes for sgns_tile.cu (runs on a B200; ``csrc/umma_probe.cu``).

Checks, against numpy, every shared-memory layout assumption of the tensor-core training kernel:

1. K-major SWIZZLE_128B A and B (the layout nn_tc.cu already uses) -- sanity;
2. ``G . V``   : A = G K-major, B = a gathered [rows x 32] tile read MN-major, N = 32;
3. ``G^T . U`` : A = G read MN-major (LBO = distance of the 32-column blocks of G), B = gathered tile MN-major,
                 incl. the overlapping second M block;
4. TMA tile::gather4 : row order, swizzle and transaction bytes.

Inputs are small integers so tf32 products and fp32 sums are exact: every comparison is ``==``.
Prints one PASS/FAIL line per hypothesis and a JSON summary; exit code 0 iff the layouts sgns_tile.cu assumes pass.
"""
import json
import sys
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glint_word2vec_b200 import _C  # noqa: E402


def sw128(mat):
    """[rows, 32] fp32 -> bytes of the SWIZZLE_128B tile (row pitch 128 B, chunk c of row r at c ^ (r & 7))."""
    rows = mat.shape[0]
    assert mat.shape[1] == 32 and rows % 8 == 0
    out = np.zeros((rows, 8, 4), dtype=np.float32)
    src = mat.reshape(rows, 8, 4)
    r = np.arange(rows)
    for c in range(8):
        out[r, c ^ (r & 7)] = src[r, c]
    return out.reshape(-1).view(np.uint8)


def desc(start, lbo, sbo):
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (2 << 61)


def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def opword(idc, acc, col=0):
    return idc | (acc << 32) | (col << 40)


def run(image, ops, ncols):
    img = torch.from_numpy(np.ascontiguousarray(image)).cuda()
    o = torch.tensor(ops, dtype=torch.int64).cuda()
    out = _C.umma_probe(img, o, ncols)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def ints(rng, shape, lo=-4, hi=5):
    return rng.integers(lo, hi, size=shape).astype(np.float32)


def main():
    rng = np.random.default_rng(0)
    res = {}

    # ---- 1. K-major A [128 x 32] . B [64 x 32]^T
    A = ints(rng, (128, 32)); B = ints(rng, (64, 32))
    image = np.concatenate([sw128(A), sw128(B)])
    ops = [[desc(32 * k, 16, 1024), desc(16384 + 32 * k, 16, 1024), opword(idesc(128, 64), 1 if k else 0)] for k in range(4)]
    got = run(image, ops, 64)
    res["kmajor"] = bool(np.array_equal(got, A @ B.T))
    print("1. K-major A,B             :", "PASS" if res["kmajor"] else "FAIL")

    # ---- 2. dU = G . V : G [128 x R] as R/32 K-major blocks, V tile [R x 32] read MN-major
    R = 192
    G = ints(rng, (128, R)); V = ints(rng, (R, 32))
    gimg = np.concatenate([sw128(G[:, 32 * b:32 * b + 32]) for b in range(R // 32)])
    voff = gimg.size
    image = np.concatenate([gimg, sw128(V)])
    want = G @ V
    for name, lbo, sbo in (("lbo=16,sbo=1024", 16, 1024), ("lbo=1024,sbo=16", 1024, 16), ("lbo=4096,sbo=1024", 4096, 1024)):
        ops = []
        for kk in range(R // 8):
            a = desc((kk // 4) * 16384 + (kk % 4) * 32, 16, 1024)
            b = desc(voff + kk * 1024, lbo, sbo)
            ops.append([a, b, opword(idesc(128, 32, 0, 1), 1 if kk else 0)])
        got = run(image, ops, 32)
        ok = bool(np.array_equal(got, want))
        res["gv_" + name] = ok
        print(f"2. G.V   B MN-major {name:18s}:", "PASS" if ok else "FAIL")

    # ---- 3. dV = G^T . U : A = G^T read MN-major out of the same G image, B = U tile [128 x 32] MN-major
    U = ints(rng, (128, 32))
    uoff = gimg.size
    image = np.concatenate([gimg, sw128(U)])
    for j0 in (0, 64):
        want = G[:, j0:j0 + 128].T @ U
        for name, lbo, sbo in (("lbo=16384,sbo=1024", 16384, 1024), ("lbo=1024,sbo=16384", 1024, 16384)):
            ops = []
            for kk in range(16):
                a = desc((j0 // 32) * 16384 + kk * 1024, lbo, sbo)
                b = desc(uoff + kk * 1024, 16, 1024)
                ops.append([a, b, opword(idesc(128, 32, 1, 1), 1 if kk else 0)])
            got = run(image, ops, 32)
            ok = bool(np.array_equal(got, want))
            res[f"gtu_j{j0}_" + name] = ok
            print(f"3. G^T.U j0={j0:3d} A MN-major {name:20s}:", "PASS" if ok else "FAIL")

    # ---- 3b. N = 64 through two 32-column tiles (LBO of B = tile stride): wider phase-B MMAs
    U2 = ints(rng, (128, 64))
    image = np.concatenate([gimg, sw128(U2[:, :32]), sw128(U2[:, 32:])])
    want = G[:, :128].T @ U2
    ops = []
    for kk in range(16):
        a = desc(kk * 1024, 16384, 1024)
        b = desc(uoff + kk * 1024, 16384, 1024)
        ops.append([a, b, opword(idesc(128, 64, 1, 1), 1 if kk else 0)])
    got = run(image, ops, 64)
    res["gtu_n64"] = bool(np.array_equal(got, want))
    print("3b. G^T.U N=64 (LBO(B) = 16 KB)  :", "PASS" if res["gtu_n64"] else "FAIL")

    # ---- 4. gather4
    table = ints(rng, (1000, 64), -100, 100)
    rows = rng.integers(0, 1000, size=32).astype(np.int32)
    tdev = torch.from_numpy(table).cuda()
    rdev = torch.from_numpy(rows).cuda()
    for col in (0, 32):
        out = _C.gather4_probe(tdev, rdev, col, 32, 512, 0)
        torch.cuda.synchronize()
        want = sw128(table[rows, col:col + 32])
        ok = bool(np.array_equal(out.cpu().numpy(), want))
        res[f"gather4_col{col}"] = ok
        print(f"4. gather4 col={col:2d}            :", "PASS" if ok else "FAIL")
        if not ok:
            g = out.cpu().numpy().view(np.float32).reshape(-1, 32)
            print("   first rows got :", g[:2, :8], "\n   want (unswizzled):", table[rows[:2], col:col + 8])
    # partially out-of-bounds columns (K = 40: second 32-column block has 8 valid columns) must zero-fill
    t40 = np.ascontiguousarray(table[:, :40])
    out = _C.gather4_probe(torch.from_numpy(t40).cuda(), rdev, 32, 32, 512, 0)
    torch.cuda.synchronize()
    w = np.zeros((32, 32), np.float32); w[:, :8] = t40[rows, 32:40]
    res["gather4_oob"] = bool(np.array_equal(out.cpu().numpy(), sw128(w)))
    print("4b. gather4 OOB zero fill (K=40):", "PASS" if res["gather4_oob"] else "FAIL")

    need = ["kmajor", "gv_lbo=16,sbo=1024", "gtu_j0_lbo=16384,sbo=1024", "gtu_j64_lbo=16384,sbo=1024",
            "gather4_col0", "gather4_col32", "gather4_oob"]
    ok = all(res.get(k, False) for k in need)
    print(json.dumps({"probe": res, "assumed_ok": ok}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
