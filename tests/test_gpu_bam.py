"""BGZF / BAM ingest on the GPU (pssgpu_feed_bam) against the path it replaces: `samtools view` of the same
file (an independent Python reader here) -> oracle tally (pss-bam.c:148-162 + :764-783, fragkon.c:84-93 + :342-363).
Every count and every outcome counter must be equal; malformed files must be reported, not crash."""
import importlib
import os

import numpy as np
import pytest

from pss_testlib import (FkParams, Oracle, PssParams, Synth, bam_to_sam, bgzf_blocks, oracle_parallel, reads_cfg_config1,
                         reads_cfg_config2)
from test_bam_logic import _edge_sam, _refs

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("pss-bam_b200")


@pytest.fixture(scope="module")
def small():
    g = Synth.genome(31, [300000, 200000, 5000], n_frac=0.01, lower_frac=0.03)
    ora = Oracle(fasta=g.fasta_bytes())
    ctx = pkg.Context(0)
    ctx.upload_genome(ora.contigs())
    yield g, ora, ctx
    ctx.close()


def _feed_chunks(ctx, bam, chunk):
    if chunk is None:
        ctx.feed_bam(bam, last=True)
        return
    for off in range(0, len(bam), chunk):
        ctx.feed_bam(bam[off:off + chunk], last=False)
    ctx.feed_bam(b"", last=True)


def _check_pss(ctx, ora, bam, view, p=PssParams(), chunk=None):
    f, r, st = ora.pss(view, p)
    ctx.pss_begin(pkg.PssOptions(p.region_len, p.min_len, p.max_len, p.min_mq, p.up_ctx, p.down_ctx, p.merged_only))
    _feed_chunks(ctx, bam, chunk)
    gf, gr = ctx.pss_finish()
    assert ctx.stats() == st
    assert np.array_equal(gf, f) and np.array_equal(gr, r)


@pytest.mark.parametrize("level,block,chunk", [(6, 0, None), (0, 0, 1 << 20), (1, 700, 65536), (9, 90, 1000), (6, 0, 1), (6, 3000, 17)])
def test_bam_equals_samtools_view_path(small, level, block, chunk):
    g, ora, ctx = small
    n = 3000 if chunk in (1, 17) else 30000
    sam = Synth.sam(reads_cfg_config2(seed=7), g, 0, n)
    bam = Synth.bam(sam, _refs(g), level=level, block_payload=block, qual_mode=1)
    view = bam_to_sam(bam) if n <= 3000 else sam          # (the reader reproduces the generator's lines: test_bam_logic)
    _check_pss(ctx, ora, bam, view, chunk=chunk)
    info = ctx.bam_info()
    assert info["records"] == n and info["references"] == 4 and info["dropped_by_read_group"] == 0
    # fragkon and the fused pass over the same bytes
    fp, tp, fst = ora.fragkon(view, FkParams(klen=8))
    ctx.fragkon_begin(pkg.FragkonOptions(klen=8))
    _feed_chunks(ctx, bam, chunk if chunk not in (1, 17) else 4096)
    gfp, gtp = ctx.fragkon_finish()
    assert ctx.stats() == fst and np.array_equal(gfp, fp) and np.array_equal(gtp, tp)
    pp = PssParams(min_len=30, max_len=150, min_mq=30)
    f, r, st = ora.pss(view, pp)
    fp, tp, fst = ora.fragkon(view, FkParams(klen=6, min_len=30, max_len=150, min_mq=30))
    ctx.both_begin(pkg.PssOptions(min_len=30, max_len=150, min_mq=30), pkg.FragkonOptions(klen=6, min_len=30, max_len=150, min_mq=30))
    ctx.feed_bam(bam, last=True)
    gf, gr = ctx.pss_finish()
    gfp, gtp = ctx.fragkon_finish()
    assert ctx.stats() == st and ctx.fragkon_stats() == fst
    assert np.array_equal(gf, f) and np.array_equal(gr, r) and np.array_equal(gfp, fp) and np.array_equal(gtp, tp)


def test_bam_read_group_filter(small):
    """`samtools view -r RG` natively (pss-bam.c:153-155)."""
    g, ora, ctx = small
    sam = Synth.sam(reads_cfg_config2(seed=8), g, 0, 9000)
    bam = Synth.bam(sam, _refs(g), rg_mode=1)
    try:
        for rg, kept in (("rgA", 3000), ("rgB", 3000), ("nope", 0)):
            ctx.bam_read_group(rg)
            view = bam_to_sam(bam, read_group=rg)
            assert view.count(b"\n") == kept
            _check_pss(ctx, ora, bam, view, chunk=50000)
            info = ctx.bam_info()
            assert info["records"] == 9000 and info["dropped_by_read_group"] == 9000 - kept
    finally:
        ctx.bam_read_group(None)
    _check_pss(ctx, ora, bam, bam_to_sam(bam))


def test_bam_edge_records(small):
    g, ora, ctx = small
    sam = _edge_sam(g)
    for block in (0, 64, 5000):
        bam = Synth.bam(sam, _refs(g), block_payload=block)
        view = bam_to_sam(bam)
        for p in (PssParams(), PssParams(region_len=1), PssParams(region_len=0)):
            _check_pss(ctx, ora, bam, view, p)
            _check_pss(ctx, ora, bam, view, p, chunk=100)


def test_bam_many_references_and_long_header(small):
    """3000 contigs (global contig table, a 150 KB BAM header that spans several BGZF blocks) fed in small pieces: the
    header is completed across batches on the device."""
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(300, 900, size=3000)]
    names = [f"scaffold_{i:05d}_{'x' * int(rng.integers(0, 30))}" for i in range(3000)]
    g = Synth.genome(33, lens, names=names)
    ora = Oracle(fasta=g.fasta_bytes())
    ctx = pkg.Context(0)
    ctx.upload_genome(ora.contigs())
    sam = Synth.sam(reads_cfg_config2(seed=9, min_len=20, max_len=60), g, 0, 20000)
    bam = Synth.bam(sam, _refs(g), block_payload=20000)
    _check_pss(ctx, ora, bam, sam)
    _check_pss(ctx, ora, bam, sam, chunk=30000)
    assert ctx.bam_info()["references"] == 3001
    ctx.close()


def test_bam_multiple_batches_equal_text_path():
    """Enough data for several device batches ($PSSGPU_BAM_BATCH_MB = 16 MiB of compressed bytes each here): tables equal
    those of the SAM text path and of the oracle; first-record guesses are (almost) never wrong on real-sized blocks."""
    os.environ["PSSGPU_BAM_BATCH_MB"] = "16"              # read when the context first ingests BAM
    Synth.set_threads(os.cpu_count() or 1)
    g = Synth.genome(35, [40_000_000, 25_000_000, 3_000_000], n_frac=0.01, lower_frac=0.03)
    n = 1_600_000
    sam = Synth.sam(reads_cfg_config2(seed=10), g, 0, n)
    bam = Synth.bam(sam, _refs(g), level=1, qual_mode=1)
    assert len(bam) > (70 << 20)
    ora = Oracle(contigs=list(zip(g.names, g.seqs)))
    f, r, st = oracle_parallel(ora, sam, "pss", PssParams(), os.cpu_count() or 1)
    ctx = pkg.Context(0)
    ctx.upload_genome(list(zip(g.names, g.seqs)))
    tf, tr = ctx.pss(sam)
    assert np.array_equal(tf, f) and np.array_equal(tr, r)
    for chunk in (None, 10_000_001):
        ctx.pss_begin(pkg.PssOptions())
        _feed_chunks(ctx, bam, chunk)
        gf, gr = ctx.pss_finish()
        assert ctx.stats() == st
        assert np.array_equal(gf, f) and np.array_equal(gr, r)
        info = ctx.bam_info()
        assert info["records"] == n and info["batches"] >= 4
        assert info["blocks_rewalked"] <= 4 * info["batches"]
    ctx.close()
    del os.environ["PSSGPU_BAM_BATCH_MB"]


def test_bam_malformed_input_is_reported(small):
    g, ora, ctx = small
    sam = Synth.sam(reads_cfg_config1(seed=11), g, 0, 5000)
    bam = Synth.bam(sam, _refs(g))
    blocks = bgzf_blocks(bam)

    def run(data, expect_ok=False):
        ctx.pss_begin(pkg.PssOptions())
        try:
            ctx.feed_bam(data, last=True)
            ctx.pss_finish()
        except pkg.PssGpuError as ex:
            assert not expect_ok, ex
            return str(ex)
        assert expect_ok, "malformed input went unnoticed"
        return ""

    run(bam, expect_ok=True)
    assert "BGZF" in run(bam[:len(bam) // 2])                              # ends inside a block
    cut = blocks[len(blocks) // 2][0]
    assert "record" in run(bam[:cut]) or "ends inside" in run(bam[:cut])   # ends between blocks, inside a record
    assert "BGZF" in run(b"@HD\tVN:1.6\n" + sam[:1000])                    # SAM text is not BGZF
    bad = bytearray(bam)
    at, total, po, pl, isize = blocks[1]
    for k in range(po + 10, po + 40):
        bad[k] ^= 0x5a                                                     # garbage in a deflate stream
    run(bytes(bad))
    import zlib
    raw = b"XAM\x01" + b"\0" * 100                                         # a BGZF file that is not BAM
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(raw) + co.flush()
    blk = (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + (len(comp) + 25).to_bytes(2, "little") + comp
           + zlib.crc32(raw).to_bytes(4, "little") + len(raw).to_bytes(4, "little"))
    assert "magic" in run(blk)
    # the CRC32 of every block is checked on the device (htslib does, so the reference never tallies a damaged block):
    # a wrong trailer, and a stored block whose payload changed under an intact deflate structure
    bad = bytearray(bam)
    at, total, po, pl, isize = blocks[len(blocks) // 2]
    bad[at + total - 8] ^= 0x01
    assert "CRC32" in run(bytes(bad))
    bam0 = Synth.bam(sam, _refs(g), level=0)
    b0 = bgzf_blocks(bam0)
    at, total, po, pl, isize = b0[len(b0) // 2]
    bad = bytearray(bam0)
    bad[po + pl - 1] ^= 0x40                                               # last payload byte of a stored block: a QUAL or tag byte
    assert "CRC32" in run(bytes(bad))
    run(bam0, expect_ok=True)
    # and the context still works afterwards
    _check_pss(ctx, ora, bam, sam)


def test_bam_crc_check_can_be_switched_off(small):
    """$PSSGPU_BAM_CRC=0 (read when a context first ingests BAM): a wrong CRC32 trailer is then ignored, as before."""
    g, ora, _ = small
    sam = Synth.sam(reads_cfg_config1(seed=12), g, 0, 3000)
    bam = Synth.bam(sam, _refs(g))
    blocks = bgzf_blocks(bam)
    bad = bytearray(bam)
    at, total, po, pl, isize = blocks[1]
    bad[at + total - 7] ^= 0x80
    os.environ["PSSGPU_BAM_CRC"] = "0"
    try:
        ctx = pkg.Context(0)
        ctx.upload_genome(ora.contigs())
        _check_pss(ctx, ora, bytes(bad), sam)
        ctx.close()
    finally:
        del os.environ["PSSGPU_BAM_CRC"]
