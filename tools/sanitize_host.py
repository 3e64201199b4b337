#!/usr/bin/env python
"""AddressSanitizer + UndefinedBehaviorSanitizer run of the two native HOST modules
(csrc/match_replay.cpp, csrc/cluster_graph.cpp) on random inputs; no GPU needed.

    python tools/sanitize_host.py [cases]

Builds both files with `g++ -fsanitize=address,undefined` into a scratch library, re-executes
itself with the sanitizer runtimes preloaded and drives the entry points through ctypes with
random component / overlap tables (ties, merges, empty slices) and random instance graphs. Any
finding aborts the run with the sanitizer's report. Last run: clean (round 2)."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "empanada-napari_b200", "csrc")
P, I, LL, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double


def build(workdir):
    stub = os.path.join(workdir, "stub.cpp")
    with open(stub, "w") as f:
        f.write('#include <cstdio>\nextern "C" int be_set_error(const char* m) { std::fprintf(stderr, "ERR %s\\n", m); return -1; }\n')
    lib = os.path.join(workdir, "host_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", os.path.join(CSRC, "match_replay.cpp"),
                           os.path.join(CSRC, "cluster_graph.cpp"), stub, "-o", lib])
    return lib


def random_tables(rng, n, maxc, density, tie):
    n_cc = rng.integers(0, maxc + 1, n).astype(np.int32)
    if rng.random() < 0.3:
        n_cc[rng.integers(0, n)] = 0
    cap = max(1, int(n_cc.max()))
    table = np.zeros((n, cap, 5), np.int32)
    for s in range(n):
        c = n_cc[s]
        table[s, :c, 0] = rng.integers(2, 5, c) if tie else rng.integers(1, 200, c)
        y0, x0 = rng.integers(0, 50, c), rng.integers(0, 50, c)
        table[s, :c, 1], table[s, :c, 2] = y0, x0
        table[s, :c, 3], table[s, :c, 4] = y0 + rng.integers(1, 20, c), x0 + rng.integers(1, 20, c)
    keys, vals = [], []
    for s in range(1, n):
        a, b = int(n_cc[s - 1]), int(n_cc[s])
        if a == 0 or b == 0:
            continue
        npair = min(a * b, rng.binomial(a * b, min(1.0, density / max(a, b))))
        if npair == 0:
            continue
        flat = rng.choice(a * b, size=npair, replace=False)
        q, c = flat // b + 1, flat % b + 1
        lim = np.minimum(table[s - 1, q - 1, 0], table[s, c - 1, 0])
        v = np.ones(npair, np.int64) if tie else np.maximum(1, (lim * rng.uniform(0.05, 1.0, npair) / 3).astype(np.int64))
        keys.append((np.uint64(s) << np.uint64(40)) | (q.astype(np.uint64) << np.uint64(20)) | c.astype(np.uint64))
        vals.append(v.astype(np.int32))
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.int32)
    perm = rng.permutation(len(keys))
    return n_cc, table, np.ascontiguousarray(keys[perm]), np.ascontiguousarray(vals[perm])


def drive(lib_path, cases):
    L = ctypes.CDLL(lib_path)
    L.be_match_replay.argtypes = [I, P, P, I, P, P, LL, I, I, D, D, I, P, I, P, P, P, I, P]
    L.be_components_clusters.argtypes = [I, P, P, P, P, P, P, P, LL, D, D, D, I, P, P]
    L.be_components_clusters_fetch.argtypes = [P, P]
    rng = np.random.default_rng(2026)
    for _ in range(cases):
        n = int(rng.integers(1, 40))
        n_cc, table, keys, vals = random_tables(rng, n, int(rng.choice([1, 3, 8, 30, 80])),
                                                float(rng.choice([0.3, 0.8, 1.2, 2.0, 4.0])), bool(rng.random() < 0.4))
        cap = table.shape[1]
        lut = np.zeros((n, cap + 1), np.int32)
        mi = int(n_cc.sum()) + 1
        labels, sizes, boxes, k = np.zeros(mi, np.int32), np.zeros(mi, np.int64), np.zeros((mi, 6), np.int32), np.zeros(1, np.int32)
        rc = L.be_match_replay(n, n_cc.ctypes.data, table.ctypes.data, cap, keys.ctypes.data, vals.ctypes.data, len(keys), 1,
                               1000, 0.25, 0.25, int(rng.integers(0, 3)), lut.ctypes.data, cap + 1, labels.ctypes.data,
                               sizes.ctypes.data, boxes.ctypes.data, mi, k.ctypes.data)
        assert rc == 0
    comps = 0
    for _ in range(max(1, cases // 5)):
        n, m = int(rng.integers(2, 200)), int(rng.integers(1, 300))
        base = int(rng.choice([0, 1000, 40000, 3000000]))
        ids = np.sort(rng.choice(np.arange(base, base + 5 * n), size=n, replace=False))
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
        key = np.unique(a[a < b] * 100000 + b[a < b])
        a, b = key // 100000, key % 100000
        parent = list(range(n))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for u, v in zip(a.tolist(), b.tolist()):
            parent[find(u)] = find(v)
        roots = np.array([find(x) for x in range(n)])
        iou = np.where(rng.random(len(a)) < 0.5, rng.uniform(0.7, 1.0, len(a)), rng.uniform(0.0, 0.05, len(a)))
        ov = np.where(rng.random(len(a)) < 0.5, rng.integers(1, 90, len(a)), rng.integers(90, 400, len(a)))
        for r in np.unique(roots):
            members = np.flatnonzero(roots == r)
            if len(members) < 2:
                continue
            sel = np.flatnonzero(roots[a] == r)
            mem = np.ascontiguousarray(ids[members], dtype=np.int32)
            ea, eb = np.ascontiguousarray(ids[a[sel]], dtype=np.int32), np.ascontiguousarray(ids[b[sel]], dtype=np.int32)
            ei, eo = np.ascontiguousarray(iou[sel], dtype=np.float64), np.ascontiguousarray(ov[sel], dtype=np.int64)
            noff, eoff = np.array([0, len(mem)], np.int32), np.array([0, len(ea)], np.int32)
            ncl, tot = np.zeros(1, np.int32), np.zeros(2, np.int64)
            rc = L.be_components_clusters(1, noff.ctypes.data, mem.ctypes.data, eoff.ctypes.data, ea.ctypes.data, eb.ctypes.data,
                                          ei.ctypes.data, eo.ctypes.data, int(ids.max()) + 1 + int(rng.integers(0, 50)),
                                          float(rng.choice([0.0, 0.5, 0.75, 0.9])), 0.01, 100.0, 1, ncl.ctypes.data, tot.ctypes.data)
            assert rc == 0
            sz, out = np.zeros(max(1, int(tot[0])), np.int32), np.zeros(max(1, int(tot[1])), np.int32)
            L.be_components_clusters_fetch(sz.ctypes.data, out.ctypes.data)
            comps += 1
    print(f"sanitizer run finished without findings: {cases} tracker replays, {comps} instance-graph components")


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    if os.environ.get("B200_SANITIZE_CHILD"):
        drive(os.environ["B200_SANITIZE_CHILD"], cases)
        return
    with tempfile.TemporaryDirectory() as work:
        lib = build(work)
        runtimes = " ".join(subprocess.check_output(["gcc", f"-print-file-name={n}"], text=True).strip()
                            for n in ("libasan.so", "libubsan.so"))
        env = dict(os.environ, LD_PRELOAD=runtimes, ASAN_OPTIONS="detect_leaks=0", B200_SANITIZE_CHILD=lib)
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__), str(cases)], env=env))


if __name__ == "__main__":
    main()
