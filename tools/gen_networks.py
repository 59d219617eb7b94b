#!/usr/bin/env python
"""Generate hdp_b200/csrc/thr_net_gen.cuh: straight-line compare-exchange networks for k_thr_net (thr_net.cu).

    python tools/gen_networks.py            # rewrites the header
    python tools/gen_networks.py --check    # verifies every network (0/1 principle for the merges, random + 0/1 samples for the sorts)

Every network works on values kept in registers, one grid cell per lane, and is data-oblivious: the same instruction
sequence for all 32 cells of a warp.  Two kinds:

  sort<N>            N values -> descending order (Knuth's merge exchange, TAOCP 5.2.2 Algorithm M: any N)
  merge<K, NB>       two descending lists of K and NB values -> the K largest of their union, descending
                     (Batcher's odd-even merge or the bitonic merge, whichever is smaller after pruning)

Networks are built on power-of-two wire counts with virtual -inf padding and then pruned symbolically: a comparator
whose lower input is the virtual -inf disappears, one whose upper input is -inf becomes a renaming, and operations that
cannot reach one of the K wanted outputs are dropped (half a comparator = one FMNMX is kept where only its max or only
its min is live).  The header holds the surviving max/min operations in SSA form.
"""
import argparse
import itertools
import os
import random
import sys

NEG = None                      # the virtual -inf


def merge_exchange(n):
    """Knuth Algorithm M: sorting network for any n; comparator (i, j), i < j."""
    if n < 2:
        return []
    t = (n - 1).bit_length()
    ces = []
    p = 1 << (t - 1)
    while p > 0:
        q, r, d = 1 << (t - 1), 0, p
        while True:
            for i in range(n - d):
                if (i & p) == r:
                    ces.append((i, i + d))
            if q == p:
                break
            d, q, r = q - p, q >> 1, p
        p >>= 1
    return ces


def oddeven_merge(lo, n, r, out):
    m = r * 2
    if m < n:
        oddeven_merge(lo, n, m, out)
        oddeven_merge(lo + r, n, m, out)
        for i in range(lo + r, lo + n - r, m):
            out.append((i, i + r))
    else:
        out.append((lo, lo + r))


def bitonic_merge(n):
    ces = []
    m = n // 2
    while m >= 1:
        for i in range(n):
            if (i & m) == 0:
                ces.append((i, i + m))
        m //= 2
    return ces


class Sym:
    """Symbolic run of a comparator list over wires holding input names, SSA temporaries or the virtual -inf."""

    def __init__(self, wires):
        self.w = list(wires)
        self.ops = []               # (dst, 'max'|'min', a, b)
        self.n_tmp = 0

    def ce(self, i, j):             # wire i <- max, wire j <- min   (descending by wire index)
        a, b = self.w[i], self.w[j]
        if b is NEG:
            return
        if a is NEG:
            self.w[i], self.w[j] = b, NEG
            return
        hi, lo = f"t{self.n_tmp}", f"t{self.n_tmp + 1}"
        self.n_tmp += 2
        self.ops.append((hi, "max", a, b))
        self.ops.append((lo, "min", a, b))
        self.w[i], self.w[j] = hi, lo

    def prune(self, outputs):
        live = set(o for o in outputs if o is not NEG)
        kept = []
        for dst, op, a, b in reversed(self.ops):
            if dst in live:
                kept.append((dst, op, a, b))
                live.add(a); live.add(b)
        kept.reverse()
        return kept


def build_sort(n):
    s = Sym([f"v[{i}]" for i in range(n)])
    for i, j in merge_exchange(n):
        s.ce(i, j)
    outs = s.w[:n]
    return s.prune(outs), outs


def build_merge(k, nb, kind):
    kp = 1
    while kp < max(k, nb):
        kp *= 2
    if kind == "oddeven":
        wires = [f"a[{i}]" if i < k else NEG for i in range(kp)] + [f"b[{i}]" if i < nb else NEG for i in range(kp)]
        ces = []
        oddeven_merge(0, 2 * kp, 1, ces)
    else:                       # bitonic: a descending, -inf valley, b ascending
        wires = [f"a[{i}]" if i < k else NEG for i in range(kp)] + [f"b[{kp - 1 - i}]" if kp - 1 - i < nb else NEG for i in range(kp)]
        ces = bitonic_merge(2 * kp)
    s = Sym(wires)
    for i, j in ces:
        s.ce(i, j)
    outs = s.w[:k]
    return s.prune(outs), outs


def best_merge(k, nb):
    cands = [(build_merge(k, nb, kind), kind) for kind in ("oddeven", "bitonic")]
    (ops, outs), kind = min(cands, key=lambda c: len(c[0][0]))
    return ops, outs, kind


def evaluate(ops, outs, env):
    """Runs a network on many input vectors at once: env maps input names to NumPy arrays of one value per test case."""
    import numpy as np
    env = dict(env)
    for dst, op, a, b in ops:
        env[dst] = np.maximum(env[a], env[b]) if op == "max" else np.minimum(env[a], env[b])
    n = len(next(iter(env.values())))
    return np.stack([env[o] if o is not NEG else np.full(n, -np.inf) for o in outs], axis=1)


def check_sort(n, ops, outs, rng):
    import numpy as np
    gen = np.random.default_rng(rng.randrange(1 << 30))
    vals = np.concatenate([gen.integers(0, 2, (4000, n)), gen.integers(0, 40, (2000, n)), gen.permuted(np.tile(np.arange(n), (2000, 1)), axis=1)])
    got = evaluate(ops, outs, {f"v[{i}]": vals[:, i].astype(float) for i in range(n)})
    assert np.array_equal(got, -np.sort(-vals.astype(float), axis=1)), n


def check_merge(k, nb, ops, outs):
    # 0/1 principle restricted to sorted inputs: every (number of ones in a, number of ones in b) pair, all at once
    import numpy as np
    oa, ob = np.meshgrid(np.arange(k + 1), np.arange(nb + 1), indexing="ij")
    oa, ob = oa.ravel(), ob.ravel()
    env = {f"a[{i}]": (oa > i).astype(float) for i in range(k)}
    env.update({f"b[{i}]": (ob > i).astype(float) for i in range(nb)})
    got = evaluate(ops, outs, env)
    want = (np.arange(k)[None, :] < (oa + ob)[:, None]).astype(float)       # the k largest of oa + ob ones and the rest zeros
    assert np.array_equal(got, want), (k, nb)


SORTS = [8, 16, 24, 30, 32]
MERGES = [(k, nb) for k in (16, 32, 48, 64) for nb in sorted({8, 16, 24, 30, 32, k}) if nb <= k or nb <= 32]


def emit(path):
    lines = ["// thr_net_gen.cuh - GENERATED by tools/gen_networks.py; do not edit.",
             "// Compare-exchange networks for k_thr_net: sort<N> (descending) and merge<K, NB> (K largest of two descending lists).",
             "#pragma once", "", "namespace hdp { namespace net {", "",
             "template <int N> struct Sort;", "template <int K, int NB> struct Merge;", ""]
    summary = []
    for n in SORTS:
        ops, outs = build_sort(n)
        summary.append(f"sort<{n}>: {len(ops)} min/max")
        lines.append(f"template <> struct Sort<{n}> {{   // {len(ops)} min/max operations")
        lines.append(f"    static __device__ __forceinline__ void run(float (&v)[{n}]) {{")
        for dst, op, a, b in ops:
            lines.append(f"        const float {dst} = f{op}f({a}, {b});")
        for i, o in enumerate(outs):
            lines.append(f"        v[{i}] = {o};")
        lines.append("    }")
        lines.append("};")
        lines.append("")
    for k, nb in MERGES:
        ops, outs, kind = best_merge(k, nb)
        summary.append(f"merge<{k},{nb}>: {len(ops)} min/max ({kind})")
        lines.append(f"template <> struct Merge<{k}, {nb}> {{   // {len(ops)} min/max operations ({kind}, pruned)")
        lines.append(f"    // c = the {k} largest of a (descending, {k}) and b (descending, {nb}), descending; c may alias a")
        lines.append(f"    static __device__ __forceinline__ void run(float (&a)[{k}], const float (&b)[{nb}]) {{")
        for dst, op, x, y in ops:
            lines.append(f"        const float {dst} = f{op}f({x}, {y});")
        for i, o in enumerate(outs):
            if o != f"a[{i}]":
                lines.append(f"        a[{i}] = {o};")
        lines.append("    }")
        lines.append("};")
        lines.append("")
    lines.append("}}  // namespace hdp::net")
    lines.insert(2, "// " + "; ".join(summary))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return summary


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if args.check:
        rng = random.Random(1)
        for n in SORTS:
            check_sort(n, *build_sort(n), rng)
        for k, nb in MERGES:
            ops, outs, _ = best_merge(k, nb)
            check_merge(k, nb, ops, outs)
        print("all networks verified")
        return
    for s in emit(os.path.join(root, "hdp_b200", "csrc", "thr_net_gen.cuh")):
        print(s)


if __name__ == "__main__":
    main()
