#!/usr/bin/env python
"""Probe 3 (B200): MN-major tf32 operands need the SWIZZLE_128B_BASE32B layout (32-byte chunks, 4-row period; probes
1-2 showed that tcgen05.mma silently does nothing with MN-major SWIZZLE_128B tf32 descriptors).  Checks:

1. TMA gather4 with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces that layout;
2. B MN-major (N = 32 / 64), A MN-major (M = 128 through four 32-wide atoms), LBO / SBO roles;
3. the exact phase-B sequences of sgns_tile.cu: dU = G . V (K = NN) and dV = G^T . U (K = 128 centres, M padded);
4. where the rows of an M = 64 accumulator live in TMEM.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glint_word2vec_b200 import _C  # noqa: E402
from benchmarks.probe_umma import sw128, desc, idesc, opword, run, ints  # noqa: E402

np.set_printoptions(linewidth=250, threshold=100000)


def sw32(mat):
    """[rows, 32] fp32 -> bytes of the SWIZZLE_128B_ATOM_32B tile: 32-byte chunk c of row r stored at c ^ (r & 3)."""
    rows = mat.shape[0]
    assert mat.shape[1] == 32 and rows % 4 == 0
    out = np.zeros((rows, 4, 8), dtype=np.float32)
    src = mat.reshape(rows, 4, 8)
    r = np.arange(rows)
    for c in range(4):
        out[r, c ^ (r & 3)] = src[r, c]
    return out.reshape(-1).view(np.uint8)


def desc1(start, lbo, sbo):
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (1 << 61)


def cases():
    rng = np.random.default_rng(1)
    A = ints(rng, (128, 32))
    Bt = ints(rng, (8, 32))
    out = {}

    def add(name, image, ops, ncols, want, rows=None):
        out[name] = (image, ops, ncols, want, rows)

    # 2a. B MN-major, one MMA: D[m][n] = sum_{k<8} A[m][k] Bt[k][n]
    image = np.concatenate([sw128(A), sw32(Bt)])
    for lbo, sbo in ((16, 512), (512, 16), (1024, 512), (512, 1024), (4096, 512), (512, 512), (1024, 1024)):
        add(f"b_mn_l{lbo}_s{sbo}", image, [[desc(0, 16, 1024), desc1(16384, lbo, sbo), opword(idesc(128, 32, 0, 1), 0)]], 32,
            A[:, :8] @ Bt)
    # 2b. B MN-major N = 64: two [8 x 32] tiles `stride` bytes apart
    Bt2 = ints(rng, (8, 64))
    for stride in (1024, 4096):
        pad = np.zeros(stride - 1024, np.uint8)
        image = np.concatenate([sw128(A), sw32(Bt2[:, :32]), pad, sw32(Bt2[:, 32:])])
        for lbo, sbo in ((stride, 512), (512, stride)):
            add(f"b_mn64_st{stride}_l{lbo}_s{sbo}", image,
                [[desc(0, 16, 1024), desc1(16384, lbo, sbo), opword(idesc(128, 64, 0, 1), 0)]], 64, A[:, :8] @ Bt2)
    # 2c. A MN-major: At [8 k-rows x 128 m] as four [8 x 32] tiles `stride` apart; B K-major [16 x 32]
    At = ints(rng, (8, 128))
    B = ints(rng, (16, 32))
    for stride in (1024, 16384):
        parts = []
        for t in range(4):
            parts.append(sw32(At[:, 32 * t:32 * t + 32]))
            if t < 3:
                parts.append(np.zeros(stride - 1024, np.uint8))
        image = np.concatenate([sw128(B), np.concatenate(parts)])
        for lbo, sbo in ((stride, 512), (512, stride)):
            add(f"a_mn_st{stride}_l{lbo}_s{sbo}", image,
                [[desc1(2048, lbo, sbo), desc(0, 16, 1024), opword(idesc(128, 16, 1, 0), 0)]], 16, At.T @ B[:, :8].T)
    # 3a. dU = Gneg . Vneg
    NN = 64
    G = ints(rng, (128, NN))
    V = ints(rng, (NN, 32))
    gk = np.concatenate([sw128(G[:, 32 * b:32 * b + 32]) for b in range(NN // 32)])
    image = np.concatenate([gk, sw32(V)])
    for sbo in (512, 1024):
        ops = []
        for kk in range(NN // 8):
            ops.append([desc((kk // 4) * 16384 + (kk % 4) * 32, 16, 1024), desc1(gk.size + kk * 1024, 1024, sbo),
                        opword(idesc(128, 32, 0, 1), 1 if kk else 0)])
        add(f"dU_seq_s{sbo}", image, ops, 32, G @ V)
    # 3b. dVneg = Gneg^T . U
    U = ints(rng, (128, 32))
    g32 = np.concatenate([sw32(G[:, 32 * b:32 * b + 32]) for b in range(NN // 32)])
    filler = np.zeros(2 * 16384, np.uint8)
    image = np.concatenate([g32, filler, sw32(U)])
    uoff = g32.size + filler.size
    ops = []
    for kk in range(16):
        ops.append([desc1(kk * 1024, 16384, 512), desc1(uoff + kk * 1024, 1024, 512), opword(idesc(128, 32, 1, 1), 1 if kk else 0)])
    add("dV_seq", image, ops, 32, G.T @ U, NN)
    # 4. M = 64 placement
    A64 = ints(rng, (64, 32))
    image = np.concatenate([sw128(A64), sw128(B)])
    add("m64", image, [[desc(0, 16, 1024), desc(8192, 16, 1024), opword(idesc(64, 16), 0)]], 16, A64[:, :8] @ B[:, :8].T)
    # discovery
    Aoh = np.zeros((128, 32), np.float32)
    for k in range(8):
        Aoh[k, k] = 1
    image = np.concatenate([sw128(Aoh), np.arange(2048, dtype=np.float32).view(np.uint8)])
    for lbo, sbo in ((512, 512), (1024, 512)):
        add(f"disc_l{lbo}_s{sbo}", image, [[desc(0, 16, 1024), desc1(16384, lbo, sbo), opword(idesc(128, 32, 0, 1), 0)]], 32, None)
    return out


def run_case(name):
    image, ops, ncols, want, rows = cases()[name]
    got = run(image, ops, ncols)
    if name == "m64":
        lanes = []
        for r in range(64):
            hit = np.where((got == want[r]).all(1))[0]
            lanes.append(int(hit[0]) if len(hit) else -1)
        print("RESULT", name, json.dumps(lanes))
    elif want is None:
        print("RESULT", name, json.dumps(got[:8].astype(np.int64).tolist()))
    else:
        g = got if rows is None else got[:rows]
        ok = bool(np.array_equal(g, want))
        extra = ""
        if not ok:
            extra = " nonzero=%d match_frac=%.3f" % (int((g != 0).sum()), float((g == want).mean()))
        print("RESULT", name, "PASS" if ok else "FAIL" + extra)


def main():
    import subprocess
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        return 0
    rng = np.random.default_rng(1)
    table = ints(rng, (1000, 64), -100, 100)
    rows = rng.integers(0, 1000, size=32).astype(np.int32)
    out = _C.gather4_probe(torch.from_numpy(table).cuda(), torch.from_numpy(rows).cuda(), 32, 32, 512, 1)
    torch.cuda.synchronize()
    print("RESULT gather4_sw32", "PASS" if np.array_equal(out.cpu().numpy(), sw32(table[rows, 32:64])) else "FAIL", flush=True)
    for name in cases():
        r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=120)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        if lines:
            print(lines[0], flush=True)
        else:
            err = [l for l in r.stderr.splitlines() if "error" in l.lower()]
            print("RESULT", name, "CRASH", err[-1] if err else r.stderr[-200:], flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
