"""Parity at the scale BASELINE.json names: the 3.1 Gb / 24-contig synthetic genome of bench.py (base offsets beyond
2^31, contigs of up to 249 Mb) with a slice of the 200 M-read set, against the oracle; plus the size-independent
properties the bench relies on (shard invariance, host-fed == device-resident)."""
import importlib
import os
import sys

import numpy as np
import pytest

from pss_testlib import FkParams, Oracle, PssParams, Synth, reads_cfg_config2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("pss-bam_b200")


def test_full_size_genome_slice_of_reads():
    import bench
    import torch
    Synth.set_threads(os.cpu_count() or 1)
    plan = bench.contig_plan(1.0)
    g = Synth.genome(bench.GENOME_SEED, [l for _, l in plan], names=[n for n, _ in plan], n_frac=0.01, lower_frac=0.03)
    assert sum(g.lens) > 3_000_000_000
    ctx = pkg.Context(0)
    ctx.upload_genome(list(zip(g.names, g.seqs)))
    cfg = reads_cfg_config2(seed=bench.READS_SEED)
    lo, n = 5 * 25_000_000, 1_500_000                       # a slice of rank 5's shard of the 200 M reads
    sam = Synth.sam(cfg, g, lo, lo + n)

    ora = Oracle(contigs=list(zip(g.names, g.seqs)))
    f, r, st = ora.pss(sam, PssParams())
    fp, tp, fst = ora.fragkon(sam, FkParams(klen=8))
    ora.close()

    gf, gr = ctx.pss(sam)
    assert ctx.stats() == st and st["counted"] > 1_000_000
    assert np.array_equal(gf, f) and np.array_equal(gr, r)
    gfp, gtp = ctx.fragkon(sam, pkg.FragkonOptions(klen=8))
    assert ctx.stats() == fst
    assert np.array_equal(gfp, fp) and np.array_equal(gtp, tp)

    # device-resident input gives the same tables as host-fed input
    dev = torch.frombuffer(bytearray(sam), dtype=torch.uint8).cuda()
    ctx.pss_begin(pkg.PssOptions())
    ctx.feed_device(dev.data_ptr(), len(sam))
    df, dr = ctx.pss_finish()
    assert np.array_equal(df, f) and np.array_equal(dr, r)

    # shard invariance: two halves cut at a newline sum to the whole
    d = importlib.import_module("pss-bam_b200.dist")
    parts = []
    for rank in range(2):
        a, b = d.shard_sam_bytes(sam, rank, 2)
        parts.append(ctx.pss(sam[a:b]))
    assert np.array_equal(parts[0][0] + parts[1][0], f) and np.array_equal(parts[0][1] + parts[1][1], r)

    # spectrum: total = number of all-ACGT k-mer starts; shards add up
    k8 = ctx.kmer_spectrum(8)
    assert np.array_equal(sum(ctx.kmer_spectrum(8, s, 3) for s in range(3)), k8)
    assert 0.95 * sum(g.lens) < int(k8.sum()) <= sum(g.lens)
    ctx.close()
