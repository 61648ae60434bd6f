"""GPU parity: the CUDA path, called through the C ABI (include/pssgpu.h), against the CPU oracle on the same
seeded inputs.  Bit-exact: every counter of every table, every per-line outcome."""
import importlib

import numpy as np
import pytest

from pss_testlib import FkParams, Oracle, PssParams, Synth, reads_cfg_config1, reads_cfg_config2

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("pss-bam_b200")


@pytest.fixture(scope="module")
def world():
    g = Synth.genome(11, [400000, 250000, 1000, 90], n_frac=0.01, lower_frac=0.05)
    ora = Oracle(fasta=g.fasta_bytes())
    ctx = pkg.Context(0)
    ctx.upload_genome(ora.contigs())
    yield g, ora, ctx
    ctx.close()


def _opts(p: PssParams):
    return pkg.PssOptions(p.region_len, p.min_len, p.max_len, p.min_mq, p.up_ctx, p.down_ctx, p.merged_only)


def _line_offsets(sam: bytes):
    a = np.frombuffer(sam, dtype=np.uint8)
    nl = np.flatnonzero(a == 10)
    starts = np.concatenate([[0], nl + 1])
    return starts[starts < len(sam)]


def _check_pss(ctx, ora, sam, p=PssParams(), chunk=None):
    f, r, st, status = ora.pss(sam, p, want_status=True)
    ctx.debug_status(True)
    ctx.pss_begin(_opts(p))
    if chunk is None:
        ctx.feed(sam, last=True)
    else:
        for i in range(0, len(sam), chunk):
            ctx.feed(sam[i:i + chunk], last=(i + chunk >= len(sam)))
    gf, gr = ctx.pss_finish()
    gst = ctx.stats()
    off, code = ctx.debug_fetch()
    ctx.debug_status(False)
    assert np.array_equal(off, _line_offsets(sam).astype(np.uint64))
    assert np.array_equal(code, status)
    assert gst == st
    assert np.array_equal(gf, f)
    assert np.array_equal(gr, r)


@pytest.mark.parametrize("cfgname,n", [("c1", 20000), ("c2", 60000)])
def test_pss_matches_oracle(world, cfgname, n):
    g, ora, ctx = world
    cfg = reads_cfg_config1(seed=5) if cfgname == "c1" else reads_cfg_config2(seed=6)
    _check_pss(ctx, ora, Synth.sam(cfg, g, 0, n))


def test_pss_options(world):
    g, ora, ctx = world
    sam = Synth.sam(reads_cfg_config2(seed=7), g, 0, 30000)
    for p in (PssParams(region_len=5), PssParams(region_len=30, min_len=40, max_len=120, min_mq=20),
              PssParams(up_ctx=b"CT", down_ctx=b"AGN"), PssParams(merged_only=1), PssParams(region_len=0),
              PssParams(region_len=31), PssParams(region_len=64, up_ctx=b"ACGTN"), PssParams(region_len=150)):   # > 30: pss_record_wide

        _check_pss(ctx, ora, sam, p)


def test_pss_chunked_feed(world):
    g, ora, ctx = world
    sam = Synth.sam(reads_cfg_config2(seed=8), g, 0, 5000)
    _check_pss(ctx, ora, sam, chunk=977)
    _check_pss(ctx, ora, sam[:-1], chunk=100003)      # last line without '\n'


@pytest.mark.parametrize("K", [1, 2, 5, 7, 8, 9, 12])
def test_fragkon_matches_oracle(world, K):
    g, ora, ctx = world
    sam = Synth.sam(reads_cfg_config2(seed=9), g, 0, 40000)
    p = FkParams(klen=K, min_mq=10 if K == 5 else 0)
    fp, tp, st, status = ora.fragkon(sam, p, want_status=True)
    ctx.debug_status(True)
    gfp, gtp = ctx.fragkon(sam, pkg.FragkonOptions(p.klen, p.min_len, p.max_len, p.min_mq, p.merged_only))
    gst = ctx.stats()
    off, code = ctx.debug_fetch()
    ctx.debug_status(False)
    assert np.array_equal(code, status)
    assert gst == st
    assert np.array_equal(gfp, fp)
    assert np.array_equal(gtp, tp)


@pytest.mark.parametrize("k", [1, 3, 6, 7, 8, 9, 10, 11, 12, 13])
def test_kmer_spectrum_matches_oracle(world, k):
    g, ora, ctx = world
    want = ora.kmer_spectrum(k)
    assert np.array_equal(ctx.kmer_spectrum(k), want)
    parts = sum(ctx.kmer_spectrum(k, s, 3) for s in range(3))
    assert np.array_equal(parts, want)


def test_malformed_lines(world):
    g, ora, ctx = world
    good = Synth.sam(reads_cfg_config1(seed=10), g, 0, 50).split(b"\n")[:-1]
    lines = []
    for i, ln in enumerate(good):
        f = ln.split(b"\t")
        k = i % 10
        if k == 0: ln = ln.replace(b"\t", b" ")                       # space separated
        elif k == 1: ln = b"  " + ln                                   # leading white space
        elif k == 2: f[1] = f[1] + b"x"; ln = b"\t".join(f)            # flag with trailing garbage
        elif k == 3: f[8] = b"0x32"; f[1] = b"99"; ln = b"\t".join(f)  # hex TLEN on a paired record
        elif k == 4: ln = ln + b"\r"                                   # CRLF
        elif k == 5: ln = b"@SQ\tSN:chr1\tLN:400000"                   # header line
        elif k == 6: ln = b""                                          # empty line
        elif k == 7: f[3] = b"+" + f[3]; ln = b"\t".join(f)            # signed POS
        elif k == 8: ln = ln.replace(b"\t", b"\t\t", 1)                # doubled tab
        lines.append(ln)
    sam = b"\n".join(lines) + b"\n"
    _check_pss(ctx, ora, sam)
