#!/usr/bin/env python3
"""genome-kmer-count's device path alone (for ncu): the bench genome resident, one k = 8 and one k = 12 spectrum.
   --skew F: F of every contig overwritten with low-complexity runs (poly-A, (CA)n, (CAG)n) -- same-bin atomics."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def skew_genome(g, frac, seed=5):
    rng = np.random.default_rng(seed)
    motifs = [b"A", b"CA", b"CAG", b"T", b"GT"]
    for s in g.seqs:
        budget = int(len(s) * frac)
        while budget > 0 and len(s) > 20000:
            L = int(rng.integers(200, 20000))
            L = min(L, budget)
            at = int(rng.integers(0, len(s) - L))
            m = motifs[int(rng.integers(0, len(motifs)))]
            s[at:at + L] = np.frombuffer((m * (L // len(m) + 1))[:L], dtype=np.uint8)
            budget -= L


def main():
    import bench
    import torch
    from pss_testlib import Synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-scale", type=float, default=1.0)
    ap.add_argument("--skew", type=float, default=0.0)
    ap.add_argument("--ks", default="8,12")
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    Synth.set_threads(os.cpu_count() or 1)
    plan = bench.contig_plan(a.genome_scale)
    g = Synth.genome(bench.GENOME_SEED, [l for _, l in plan], names=[n for n, _ in plan], n_frac=0.01, lower_frac=0.03)
    if a.skew > 0:
        skew_genome(g, a.skew)
    pkg = importlib.import_module("pss-bam_b200")
    ctx = pkg.Context(0)
    ctx.upload_genome(list(zip(g.names, g.seqs)))
    counts = torch.zeros(1 << 24, dtype=torch.int64, device="cuda")
    out = {"genome_bases": int(sum(g.lens)), "skew": a.skew}
    for k in [int(x) for x in a.ks.split(",")]:
        best = None
        for _ in range(a.reps):
            ctx.timing_reset(True)
            ctx.kmer_spectrum_device(k, counts.data_ptr())
            ms = ctx.timing()["kernel_ms"]
            best = ms if best is None else min(best, ms)
        out[f"k{k}_ms"] = best
        out[f"k{k}_total"] = int(counts[: 1 << (2 * k)].sum().item())
        out[f"k{k}_max_bin"] = int(counts[: 1 << (2 * k)].max().item())
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
