"""CPU tests of the BGZF / BAM ingest logic the device runs (pss_inflate.h, pss_bamrec.h compiled for the host in
tests/host_emul) and of the test infrastructure around it (the C BAM writer in tests/synth, the independent Python
BAM reader in pss_testlib).  Oracle for the ingest = the same alignments as SAM text through the existing path:
`samtools view` of the file (here: the Python reader) -> oracle tally (pss-bam.c:148-162 + :764-783)."""
import ctypes as C
import os
import random
import zlib

import numpy as np
import pytest

from pss_testlib import (BamEmul, FkParams, Oracle, PssParams, Synth, bam_to_sam, bgzf_blocks, bgzf_inflate,
                         reads_cfg_config1, reads_cfg_config2)


def _refs(g, extra=(("chrUn_synthetic_decoy", 1000),)):
    return list(zip(g.names, g.lens)) + list(extra)


@pytest.fixture(scope="module")
def small():
    g = Synth.genome(31, [300000, 200000, 5000], n_frac=0.01, lower_frac=0.03)
    sam = Synth.sam(reads_cfg_config2(seed=7), g, 0, 12000)
    return g, sam, Oracle(fasta=g.fasta_bytes())


@pytest.mark.parametrize("variant", ["", "walk", "lit9"])
def test_inflate_logic_matches_zlib(variant):
    """Raw DEFLATE streams of every block type (stored / fixed / dynamic, several blocks per stream, all strategies,
    misaligned starts) through the device decoder's host build -- as shipped, with a second-level table area so small
    that long codes fall back to the canonical walk, and with a 9-bit first level."""
    lib = BamEmul.lib() if not variant else BamEmul.variant_lib(variant)
    rng = random.Random(1)

    def mk(kind, n):
        if kind == 0:
            return os.urandom(n)
        if kind == 1:
            return bytes(rng.choice(b"ACGT") for _ in range(n))
        if kind == 2:
            return ((b"r123\t0\tchr1\t1000\t60\t50M\t*\t0\t0\t" + bytes(rng.choice(b"ACGT") for _ in range(50)) + b"\tIIIIIIIII\n") * (n // 50 + 1))[:n]
        if kind == 3:
            return b"A" * n
        return bytes(rng.randrange(256) if rng.random() < 0.1 else 65 for _ in range(n))

    for trial in range(1200 if not variant else 500):
        data = mk(rng.randrange(5), rng.choice([0, 1, 2, 3, 10, 100, 1000, 5000, 65280, rng.randrange(66000)]))
        n = len(data)
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, rng.choice([1, 8, 9]),
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED]))
        parts = []
        if rng.random() < 0.3 and n > 10:
            k = rng.randrange(n)
            parts += [co.compress(data[:k]), co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])), co.compress(data[k:])]
        else:
            parts.append(co.compress(data))
        parts.append(co.flush())
        comp = b"".join(parts)
        pad = rng.randrange(4)
        buf = C.create_string_buffer(b"\x55" * pad + comp + b"\xaa" * 8)
        out = C.create_string_buffer(n + 16)
        rc = lib.emul_inflate(C.addressof(buf) + pad, len(comp), C.addressof(out), n)
        assert rc == 0 and out.raw[:n] == data, (trial, n, rc)
        if n:                                            # a wrong ISIZE is an error, not an overrun
            assert lib.emul_inflate(C.addressof(buf) + pad, len(comp), C.addressof(out), n - 1) != 0
            assert lib.emul_inflate(C.addressof(buf) + pad, len(comp), C.addressof(out), n + 1) != 0


def test_crc32_lane_logic_matches_zlib():
    """pss_crc32.h: interleaved words per lane, advance tables, x^(32 k) fix-up -- every length class and alignment."""
    rng = np.random.default_rng(77)
    buf = rng.integers(0, 256, size=70000 + 16, dtype=np.uint8)
    base = (-buf.ctypes.data) % 4                      # offset of a 4-byte aligned address
    lengths = list(range(0, 300)) + [511, 512, 513, 1023, 4095, 4096, 4097, 65279, 65280, 65535, 65536]
    lengths += [int(x) for x in rng.integers(300, 65536, size=60)]
    for n in lengths:
        for mis in range(4):
            want = zlib.crc32(buf[base + mis: base + mis + n].tobytes())
            got = BamEmul.crc32(buf, base + mis, n)
            assert got == want, (n, mis, hex(got), hex(want))
    z = np.zeros(65536 + 8, dtype=np.uint8)            # all-zero and all-ones payloads
    assert BamEmul.crc32(z, 0, 65536) == zlib.crc32(bytes(65536))
    z[:] = 255
    assert BamEmul.crc32(z, 1, 65535) == zlib.crc32(b"\xff" * 65535)


def test_inflate_logic_survives_corrupt_streams():
    lib = BamEmul.lib()
    rng = random.Random(2)
    for _ in range(1500):
        data = bytes(rng.choice(b"ACGTN\t0123") for _ in range(rng.randrange(1, 4000)))
        comp = bytearray(zlib.compress(data, rng.choice([1, 6]))[2:-4])
        for _ in range(rng.randrange(1, 4)):
            comp[rng.randrange(len(comp))] ^= 1 << rng.randrange(8)
        buf = C.create_string_buffer(bytes(comp) + b"\0" * 8)
        out = C.create_string_buffer(len(data) + 16)
        lib.emul_inflate(C.addressof(buf), len(comp), C.addressof(out), len(data))     # any code, no crash, no overrun
        assert out.raw[len(data):] == b"\0" * 16


@pytest.mark.parametrize("level,block", [(6, 0), (0, 0), (1, 700), (9, 90)])
def test_writer_reader_roundtrip_and_device_inflate(small, level, block):
    g, sam, _ = small
    bam = Synth.bam(sam, _refs(g), level=level, block_payload=block)
    assert bam_to_sam(bam) == sam                       # the C writer and the Python reader agree on every field and tag
    assert BamEmul.inflate(bam) == bgzf_inflate(bam)    # device inflate logic == zlib on every block
    assert bgzf_blocks(bam)[-1][4] == 0                 # EOF block


def test_rendered_lines_tally_like_samtools_view_output(small):
    g, sam, ora = small
    bam = Synth.bam(sam, _refs(g), qual_mode=1)
    view = bam_to_sam(bam)
    text, n_rec, n_drop = BamEmul.render(bgzf_inflate(bam))
    assert n_rec == 12000 and n_drop == 0 and text.count(b"\n") == 12000
    for p in (PssParams(), PssParams(region_len=7, min_mq=20, min_len=35, max_len=120), PssParams(merged_only=1, up_ctx=b"AG")):
        want, got = ora.pss(view, p, want_status=True), ora.pss(text, p, want_status=True)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and want[2] == got[2]
        assert np.array_equal(want[3], got[3])          # line by line the same outcome
    for k in (1, 8, 11):
        want, got = ora.fragkon(view, FkParams(klen=k), want_status=True), ora.fragkon(text, FkParams(klen=k), want_status=True)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and want[2] == got[2]
        assert np.array_equal(want[3], got[3])


def test_read_group_filter(small):
    g, sam, ora = small
    bam = Synth.bam(sam, _refs(g), rg_mode=1)
    u = bgzf_inflate(bam)
    for rg in ("rgA", "rgB", "rgC", "rg"):
        view = bam_to_sam(bam, read_group=rg)
        text, n_rec, n_drop = BamEmul.render(u, read_group=rg)
        assert n_rec == 12000 and n_rec - n_drop == view.count(b"\n") == text.count(b"\n")
        want, got = ora.pss(view), ora.pss(text)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and want[2] == got[2]
    assert bam_to_sam(bam, read_group="rgA").count(b"\n") == 4000


def _edge_sam(g):
    name = g.names[0]
    seq60 = g.seqs[0][100:160].tobytes().upper().replace(b"N", b"A")
    long_seq = g.seqs[0][1000:7000].tobytes().upper().replace(b"N", b"A")
    rows = [
        b"unmapped\t4\t*\t0\t0\t*\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
        b"noseq\t0\t%s\t101\t30\t60M\t*\t0\t0\t*\t*" % name.encode(),
        b"ok\t0\t%s\t101\t30\t60M\t*\t0\t0\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"qualstar\t0\t%s\t101\t30\t60M\t*\t0\t0\t%s\t*" % (name.encode(), seq60),
        b"one\t0\t%s\t101\t30\t1M\t*\t0\t0\tA\t*" % name.encode(),
        b"hard\t16\t%s\t101\t30\t5H60M3H\t*\t0\t0\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"eqx\t0\t%s\t101\t30\t30=1X29=\t*\t0\t0\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"pair1\t99\t%s\t101\t30\t60M\t=\t101\t60\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"pair2\t147\t%s\t101\t30\t60M\t=\t101\t-60\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"pairfar\t163\t%s\t101\t30\t60M\t%s\t9000\t8960\t%s\t%s" % (name.encode(), g.names[1].encode(), seq60, b"I" * 60),
        b"iupac\t0\t%s\t101\t30\t16M\t*\t0\t0\t=ACMGRSVTWYHKDBN\tIIIIIIIIIIIIIIII" % name.encode(),
        b"n" * 250 + b"\t0\t%s\t101\t255\t60M\t*\t0\t0\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"long\t0\t%s\t1001\t60\t6000M\t*\t0\t0\t%s\t%s\tNM:i:-7\tXL:i:70000\tXN:i:-40000\tZZ:Z:a b c" % (name.encode(), long_seq, b"5" * 6000),
        b"manyops\t0\t%s\t1001\t60\t%s\t*\t0\t0\t%s\t%s" % (name.encode(), b"1M1I" * 100 + b"1M", b"A" * 201, b"I" * 201),
        b"decoy\t0\tchrUn_synthetic_decoy\t5\t30\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
        b"mapq\t0\t%s\t101\t0\t60M\t*\t0\t2147483647\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
        b"negt\t1\t%s\t101\t7\t60M\t*\t0\t-2147483648\t%s\t%s" % (name.encode(), seq60, b"I" * 60),
    ]
    return b"\n".join(rows) + b"\n"


def test_edge_records(small):
    g, _, ora = small
    sam = _edge_sam(g)
    for block in (0, 64, 5000):
        bam = Synth.bam(sam, _refs(g), block_payload=block)
        view = bam_to_sam(bam)
        assert view == sam                                  # writer and reader agree on every edge record too
        assert BamEmul.inflate(bam) == bgzf_inflate(bam)
        text, n_rec, _ = BamEmul.render(bgzf_inflate(bam))
        assert n_rec == sam.count(b"\n")
        for p in (PssParams(), PssParams(region_len=1), PssParams(region_len=0)):
            want, got = ora.pss(view, p, want_status=True), ora.pss(text, p, want_status=True)
            assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and want[2] == got[2]
            assert np.array_equal(want[3], got[3])
        want, got = ora.fragkon(view, FkParams(klen=5), want_status=True), ora.fragkon(text, FkParams(klen=5), want_status=True)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[3], got[3])
    # field for field what the text holds: FLAG RNAME POS MAPQ CIGAR TLEN SEQ as samtools prints them
    for a, b in zip(view.split(b"\n")[:-1], text.split(b"\n")[:-1]):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        assert [fa[i] for i in (1, 2, 3, 4, 5, 8, 9)] == [fb[i] for i in (1, 2, 3, 4, 5, 8, 9)]
        assert len(fb) == 11 and len(fb[10]) == len(fa[10])


def test_first_record_guess_rarely_wrong(small):
    """The guess is only a starting point (every guess is verified on the device), but it should be right nearly always:
    at BGZF's real block size, and at a block size that puts a boundary into almost every record."""
    g, sam, _ = small
    u = bgzf_inflate(Synth.bam(sam, _refs(g), qual_mode=1))
    tested, wrong, missing = BamEmul.guess_stats(u, 65280)
    assert tested > 30 and wrong == 0
    tested, wrong, missing = BamEmul.guess_stats(u, 1500)
    assert tested > 1500 and wrong <= tested // 50 and missing == 0
