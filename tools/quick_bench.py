"""Developer probe (not the contract bench): resident-SAM tally throughput on one GPU."""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pss_testlib import Synth, reads_cfg_config1, reads_cfg_config2  # noqa: E402

pkg = importlib.import_module("pss-bam_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mb", type=int, default=200)
ap.add_argument("--reads", type=int, default=4_000_000)
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mode", default="pss")
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--tally-only", action="store_true")
ap.add_argument("--min-len", type=int, default=30)
ap.add_argument("--max-len", type=int, default=150)
ap.add_argument("--sorted", action="store_true", help="coordinate-sorted reads (a real BAM): genome gathers hit L2")
a = ap.parse_args()

t = time.time()
nc = 8
g = Synth.genome(1, [a.genome_mb * 1_000_000 // nc] * nc, n_frac=0.01, lower_frac=0.03)
cfg = reads_cfg_config1(3) if a.config == 1 else reads_cfg_config2(4, min_len=a.min_len, max_len=a.max_len)
if a.sorted:
    cfg.sorted_total = a.reads
cap = Synth.sam_bound(cfg, 0, a.reads)
host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
n = Synth.sam_into(cfg, g, 0, a.reads, host.data_ptr(), cap)
print(f"synth: {time.time() - t:.1f}s, {n / 1e6:.1f} MB SAM, {n / a.reads:.1f} B/read", flush=True)

ctx = pkg.Context(0)
t = time.time()
ctx.upload_genome(list(zip(g.names, g.seqs)))
print(f"genome upload+pack: {time.time() - t:.2f}s", ctx.genome_info(), flush=True)
dev = host[:n].cuda()
torch.cuda.synchronize()

def begin():
    if a.mode == "pss":
        ctx.pss_begin(pkg.PssOptions())
    else:
        ctx.fragkon_begin(pkg.FragkonOptions(klen=a.k))

for it in range(a.iters):
    begin()
    ctx.timing_reset(True)
    ctx.feed_device(dev.data_ptr(), n)
    ctx.sync()
    tm = ctx.timing()
    ms = tm["kernel_ms"]
    print(f"iter {it}: kernel {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  {a.reads / ms / 1e6:.3f} G reads/s  launches {tm['launches']}", flush=True)
print(ctx.stats())
if a.tally_only:
    sys.exit(0)

# host-fed (e2e) path
for rep in range(2):
    begin()
    ctx.timing_reset(True)
    torch.cuda.synchronize()
    t = time.time()
    ctx.feed_ptr(host.data_ptr(), n, last=True)
    ctx.sync()
    dt = time.time() - t
    tm = ctx.timing()
    print(f"host feed (pinned): {dt * 1e3:.1f} ms  {n / dt / 1e9:.2f} GB/s  {a.reads / dt / 1e6:.2f} M reads/s  kernels {tm['kernel_ms']:.3f} ms in {tm['launches']} launches")
    print(ctx.stats())

for k in (7, 8, 9, 10, 12):
    ctx.timing_reset(True)
    c = ctx.kmer_spectrum(k)
    tm = ctx.timing()
    print(f"spectrum k={k}: kernel {tm['kernel_ms']:.3f} ms total {int(c.sum())}")
