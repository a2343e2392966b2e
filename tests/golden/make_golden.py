#!/usr/bin/env python
"""Regenerate tests/golden/* from the UNMODIFIED reference (oracle/_ref/birch_ref).

Run in the authoring container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py [--slow]

Outputs (all committed):
  genus_<level>[_s<seed>].npz   genus tables (layout: ternary_birch_b200/table.py)
  hecke_small.npz               full CSR triples for the small genera (C1, even level)
  hecke_digests.json            sha256 + nnz per conductor for the large ones (C2..C5 samples)
  records_small.npz             per-neighbor records and isometry sequences, small cases
  eigen_85085.json              rational eigenvectors of level 85085 (found here by exact
                                linear algebra on the reference's T_2, T_3, ...) and the
                                reference's eigenvalues for every good p <= 1000 (C3)
  errors.json                   which exception the reference raises in int64 mode at large p
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import RefBinary, RefError  # noqa: E402

OUT = Path(__file__).resolve().parent


def primes_upto(n):
    sieve = np.ones(n + 1, dtype=bool)
    sieve[:2] = False
    for i in range(2, int(n ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = False
    return [int(p) for p in np.nonzero(sieve)[0]]


def digest(m) -> dict:
    h = hashlib.sha256()
    for key in ("data", "indices", "indptr"):
        h.update(np.ascontiguousarray(m[key], dtype="<i4").tobytes())
    return dict(conductor=int(m["conductor"]), dim=int(m["dim"]), nnz=int(m["data"].size), sha256=h.hexdigest())


def save_genus(level, seed):
    ref = RefBinary(level, seed)
    stream = ref._stream(["genus"])
    name = f"genus_{level}.npz" if seed == 1 else f"genus_{level}_s{seed}.npz"
    np.savez_compressed(OUT / name, stream=stream)
    return ref


# ---------------------------------------------------------------------------
# exact linear algebra for the eigenvector fixture
# ---------------------------------------------------------------------------
Q = (1 << 61) - 1


def nullspace_mod(rows, ncols):
    """Right nullspace basis of an integer matrix (list of int rows) modulo the prime Q."""
    m = [[x % Q for x in r] for r in rows]
    pivots = []
    r = 0
    for c in range(ncols):
        piv = next((i for i in range(r, len(m)) if m[i][c]), None)
        if piv is None:
            continue
        m[r], m[piv] = m[piv], m[r]
        inv = pow(m[r][c], Q - 2, Q)
        m[r] = [(x * inv) % Q for x in m[r]]
        for i in range(len(m)):
            if i != r and m[i][c]:
                f = m[i][c]
                m[i] = [(x - f * y) % Q for x, y in zip(m[i], m[r])]
        pivots.append(c)
        r += 1
        if r == len(m):
            break
    free = [c for c in range(ncols) if c not in pivots]
    basis = []
    for fc in free:
        v = [0] * ncols
        v[fc] = 1
        for i, pc in enumerate(pivots):
            v[pc] = (-m[i][fc]) % Q
        basis.append(v)
    return basis


def rational_reconstruct(a, bound=1 << 28):
    """a mod Q -> (num, den) with |num|, den <= bound, or None."""
    r0, r1, s0, s1 = Q, a % Q, 0, 1
    while r1 > bound:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        s0, s1 = s1, s0 - q * s1
    if s1 == 0 or abs(s1) > bound:
        return None
    if s1 < 0:
        r1, s1 = -r1, -s1
    return r1, s1


def integral_vector(v):
    from math import gcd
    nums, dens = [], []
    for a in v:
        rr = rational_reconstruct(a)
        if rr is None:
            return None
        nums.append(rr[0])
        dens.append(rr[1])
    lcm = 1
    for d in dens:
        lcm = lcm * d // gcd(lcm, d)
    out = [n * (lcm // d) for n, d in zip(nums, dens)]
    g = 0
    for x in out:
        g = gcd(g, abs(x))
    out = [x // g for x in out]
    first = next(x for x in out if x)
    if first < 0:
        out = [-x for x in out]
    return out


def dense_from(m):
    dim = m["dim"]
    d = np.zeros((dim, dim), dtype=np.int64)
    for r in range(dim):
        lo, hi = m["indptr"][r], m["indptr"][r + 1]
        d[r, m["indices"][lo:hi]] = m["data"][lo:hi]
    return d


def find_rational_eigenvectors(ref, small_primes=(2, 3, 19, 23)):
    """Simultaneous integer-eigenvalue eigenvectors (multiplicity one) of T_p, p in small_primes."""
    hecke = {p: ref.hecke(p, True, True) for p in small_primes}
    C = len(hecke[small_primes[0]])
    found = []
    for k in range(C):
        dim = hecke[small_primes[0]][k]["dim"]
        if dim == 0:
            continue
        mats = [dense_from(hecke[p][k]) for p in small_primes]

        def search(level, stack):
            p = small_primes[level]
            T = mats[level]
            bound = int(2 * p ** 0.5)
            for e in list(range(-bound, bound + 1)) + [p + 1]:
                rows = stack + [[int(x) - (e if i == j else 0) for j, x in enumerate(row)] for i, row in enumerate(T)]
                ns = nullspace_mod(rows, dim)
                if len(ns) == 1:
                    v = integral_vector(ns[0])
                    if v is not None and all(max(abs(x) for x in v) < 2 ** 20 for _ in [0]):
                        # confirm over Z on every small prime
                        ok = True
                        vv = np.array(v, dtype=object)
                        for M in mats:
                            w = M.astype(object).dot(vv)
                            nz = next(i for i, x in enumerate(v) if x)
                            lam = w[nz] // v[nz]
                            if any(w[i] != lam * v[i] for i in range(dim)):
                                ok = False
                        if ok:
                            found.append(dict(conductor=int(hecke[p][k]["conductor"]), vector=[int(x) for x in v]))
                elif len(ns) > 1 and level + 1 < len(small_primes):
                    search(level + 1, rows)

        search(0, [])
    # de-duplicate
    uniq = []
    for f in found:
        if f not in uniq:
            uniq.append(f)
    return uniq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slow", action="store_true", help="also regenerate the minutes-long samples (C5, error cases)")
    ap.add_argument("--bench", action="store_true",
                    help="only add the digest of the bench workload itself (level 1009091, T_9973, int64: ~1 min, 8 GB of scratch) "
                         "to hecke_digests_slow.json")
    args = ap.parse_args()
    if not RefBinary.available():
        sys.exit("oracle/_ref/birch_ref missing: make -C oracle ref")
    if args.bench:
        ref = RefBinary(1009091, 1)
        slow = json.load(open(OUT / "hecke_digests_slow.json")) if (OUT / "hecke_digests_slow.json").exists() else {}
        slow["L1009091_p9973_z64"] = [digest(m) for m in ref.hecke(9973, False, True)]
        with open(OUT / "hecke_digests_slow.json", "w") as f:
            json.dump(slow, f)
        print("bench workload digest written")
        return

    small = {}
    digests = {}
    records = {}

    # ---- C1: BirchGenus(11), all good p < 100, every entry ----------------
    ref = save_genus(11, 1)
    for p in primes_upto(100):
        if p == 11:
            continue
        for precise in (False, True):
            for k, m in enumerate(ref.hecke(p, precise, True)):
                tag = f"L11_p{p}_{'z' if precise else 'z64'}_k{k}"
                small[tag + "_data"], small[tag + "_indices"], small[tag + "_indptr"] = m["data"], m["indices"], m["indptr"]
    for p in (2, 3, 31):
        records[f"L11_p{p}_z64_rec"] = ref.neighbor_records(p, False)
        records[f"L11_p{p}_z_rec"] = ref.neighbor_records(p, True)
        records[f"L11_p{p}_z64_iso"] = ref.isometry_sequence(p, False)
        records[f"L11_p{p}_z_iso"] = ref.isometry_sequence(p, True)

    # ---- even level, seed-dependent representative order ------------------
    ref = save_genus(210, 7)
    for p in (11, 13, 59):
        for k, m in enumerate(ref.hecke(p, False, True)):
            tag = f"L210_p{p}_z64_k{k}"
            small[tag + "_data"], small[tag + "_indices"], small[tag + "_indptr"] = m["data"], m["indices"], m["indptr"]
    records["L210_p11_z64_iso"] = ref.isometry_sequence(11, False)

    # ---- C3 genus: level 5*7*11*13*17 ---------------------------------------
    ref = save_genus(85085, 1)
    for p in (2, 3, 19, 101, 997):
        for precise in (False, True):
            digests[f"L85085_p{p}_{'z' if precise else 'z64'}"] = [digest(m) for m in ref.hecke(p, precise, True)]
    records["L85085_p3_z64_rec"] = ref.neighbor_records(3, False)
    records["L85085_p3_z64_iso"] = ref.isometry_sequence(3, False)
    records["L85085_p2_z_iso"] = ref.isometry_sequence(2, True)
    records["L85085_p19_z_rec"] = ref.neighbor_records(19, True)

    vecs = find_rational_eigenvectors(ref)
    good = [p for p in primes_upto(1000) if 85085 % p]
    pairs = [(v["conductor"], v["vector"]) for v in vecs]
    aps64 = ref.eigenvalues(pairs, good, False)
    apsz = ref.eigenvalues(pairs, good, True)
    big = [65521, 65537, 99991]
    apsbig = ref.eigenvalues(pairs, big, True)
    with open(OUT / "eigen_85085.json", "w") as f:
        json.dump(dict(level=85085, seed=1, vectors=vecs, primes=good,
                       aps_z64={str(p): aps64[p] for p in good},
                       aps_z={str(p): apsz[p] for p in good},
                       aps_z_big={str(p): apsbig[p] for p in big}), f)
    print(f"eigen_85085: {len(vecs)} rational eigenvectors, {len(good)} primes")

    # ---- C2 / C5 genus: level 11*13*17*19*23 --------------------------------
    ref = save_genus(1062347, 1)
    for p in (2, 3, 101):
        for precise in (False, True):
            digests[f"L1062347_p{p}_{'z' if precise else 'z64'}"] = [digest(m) for m in ref.hecke(p, precise, True)]
    digests["L1062347_p997_z64"] = [digest(m) for m in ref.hecke(997, False, True)]

    # ---- C4 genus: level 97*101*103 -----------------------------------------
    ref = save_genus(1009091, 1)
    for p, modes in ((2, (False, True)), (107, (False, True)), (1009, (False,))):
        for precise in modes:
            digests[f"L1009091_p{p}_{'z' if precise else 'z64'}"] = [digest(m) for m in ref.hecke(p, precise, True)]

    # ---- L=11 eigenvalues at large p (a_p of the elliptic curve 11a) --------
    ref11 = RefBinary(11, 1)
    plist = [2, 3, 5, 7, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 97, 101, 997, 7919, 65521, 65537, 99991, 1000003]
    ap11 = ref11.eigenvalues([(1, [2, -3]), (1, [1, 1])], plist, True)
    with open(OUT / "eigen_11.json", "w") as f:
        json.dump(dict(level=11, seed=1, vectors=[dict(conductor=1, vector=[2, -3]), dict(conductor=1, vector=[1, 1])],
                       aps_z={str(p): ap11[p] for p in plist}), f)

    if args.slow:
        errors = {}
        ref = RefBinary(1062347, 1)
        digests["L1062347_p10007_z64"] = [digest(m) for m in ref.hecke(10007, False, True)]
        for p in (20011, 40009):
            try:
                ref.hecke(p, False, True)
                errors[f"L1062347_p{p}_z64"] = None
            except RefError as e:
                errors[f"L1062347_p{p}_z64"] = dict(kind=e.kind, what=e.what)
        with open(OUT / "errors.json", "w") as f:
            json.dump(errors, f, indent=1)
        old = {}
        if (OUT / "hecke_digests_slow.json").exists():
            old = json.load(open(OUT / "hecke_digests_slow.json"))
        old.update({k: v for k, v in digests.items() if k in ("L1062347_p10007_z64",)})
        with open(OUT / "hecke_digests_slow.json", "w") as f:
            json.dump(old, f)
        digests.pop("L1062347_p10007_z64", None)

    np.savez_compressed(OUT / "hecke_small.npz", **small)
    np.savez_compressed(OUT / "records_small.npz", **records)
    with open(OUT / "hecke_digests.json", "w") as f:
        json.dump(digests, f)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
