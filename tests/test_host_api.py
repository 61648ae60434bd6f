"""The host-side C API (pss-bam_b200/host: sam-parse.h / fasta-genome-io.h / kmer.h surface and the table writers),
built WITHOUT the GPU library and checked against the oracle and the reference's golden outputs."""
import ctypes as C
import gzip
import json
import os
import random
import subprocess

import numpy as np
import pytest

from pss_testlib import Oracle, Synth, parse_counts_file, reads_cfg_config2, tmpdir
from test_record_logic import _mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pss-bam_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden", "v1")
MAN = json.load(open(os.path.join(GOLD, "manifest.json")))
SO = os.path.join(ROOT, "tests", "host_emul", "libpsshostapi.so")


@pytest.fixture(scope="module")
def lib():
    srcs = [os.path.join(HOST, f) for f in ("pss_sam.c", "pss_fasta.c", "pss_kmer.c", "pss_tables.c")]
    subprocess.run(["gcc", "-O2", "-g", "-std=gnu11", "-Wall", "-fPIC", "-shared", "-o", SO, *srcs, "-lz"], check=True)
    h = C.CDLL(SO)
    h.init_genome.restype = C.c_void_p
    h.init_genome.argtypes = [C.c_char_p]
    h.find_seq.restype = C.c_void_p
    h.find_seq.argtypes = [C.c_void_p, C.c_char_p]
    h.destroy_genome.argtypes = [C.c_void_p]
    h.init_KSP.restype = C.c_void_p
    h.add_to_ksp.argtypes = [C.c_char_p, C.c_void_p]
    h.kmer2count.argtypes = [C.c_char_p, C.c_void_p]
    h.kmer2count.restype = C.c_uint
    h.destroy_KSP.argtypes = [C.c_void_p]
    h.line2saml.argtypes = [C.c_char_p, C.c_void_p]
    h.pss_sub_rates.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    h.pss_write_counts.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
    h.pss_write_rates.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
    return h


class _Seq(C.Structure):
    _fields_ = [("id", C.c_char * 512), ("seq", C.c_void_p), ("len", C.c_size_t)]


class _Genome(C.Structure):
    _fields_ = [("seqs", C.POINTER(C.POINTER(_Seq))), ("dummy", C.c_void_p), ("n_seqs", C.c_size_t)]


class _Saml(C.Structure):      # sam-parse.h:20-56 layout
    _fields_ = [("qname", C.c_char * 2048), ("flag", C.c_uint), ("bits", C.c_uint16), ("rname", C.c_char * 2048),
                ("pos", C.c_ulong), ("mapq", C.c_uint), ("cigar", C.c_char * 2048), ("mrnm", C.c_char * 2048),
                ("mpos", C.c_uint), ("isize", C.c_int), ("seq_len", C.c_int), ("seq", C.c_char * 2048),
                ("qual", C.c_char * 2048), ("tags", C.c_char * 2048), ("BC", C.c_char * 2048), ("RG", C.c_char * 2048),
                ("opt_tags", C.c_char * 2048), ("aln_seq_len", C.c_int), ("NM", C.c_int), ("AS", C.c_int),
                ("XM", C.c_int), ("XO", C.c_int), ("XG", C.c_int)]


def test_saml_layout_matches_reference():
    assert C.sizeof(_Saml) == 20536          # sizeof(Saml) of the reference
    assert C.sizeof(_Seq) == 528


@pytest.mark.parametrize("gz", [False, True])
def test_init_genome_matches_oracle(lib, gz):
    g = Synth.genome(3, [5000, 70, 1234, 1], names=["chrB", "chrA", "z", "a"], n_frac=0.05, lower_frac=0.2)
    fa = g.fasta_bytes(width=61).replace(b">chrA\n", b">chrA some description here\n")
    fa += b">last\nACGT acgt\tNN\n\n>empty\n"
    d = tmpdir()
    fn = os.path.join(d, "g.fa.gz" if gz else "g.fa")
    with (gzip.open if gz else open)(fn, "wb") as f:
        f.write(fa)
    ora = Oracle(fasta=fa)
    want = ora.contigs()
    G = lib.init_genome(fn.encode())
    assert G
    gg = C.cast(G, C.POINTER(_Genome)).contents
    assert gg.n_seqs == len(want)
    for i, (cid, seq) in enumerate(want):
        s = gg.seqs[i].contents
        assert s.id.decode() == cid and s.len == len(seq)
        assert C.string_at(s.seq, s.len) == seq
    hit = lib.find_seq(G, b"chrA")
    assert hit and C.cast(hit, C.POINTER(_Seq)).contents.len == 70
    assert not lib.find_seq(G, b"chrZ")
    lib.destroy_genome(G)
    assert not lib.init_genome(os.path.join(d, "missing.fa").encode())


def test_line2saml_matches_glibc_sscanf(lib):
    olib = Oracle.lib()
    olib.ora_parse_line.argtypes = [C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_ulong), C.POINTER(C.c_uint),
                                    C.POINTER(C.c_int), C.c_char_p, C.c_char_p, C.c_char_p]
    g = Synth.genome(5, [20000, 9000])
    rng = random.Random(4)
    lines = Synth.sam(reads_cfg_config2(seed=6, min_len=20, max_len=80), g, 0, 3000).split(b"\n")[:-1]
    sp = _Saml()
    n_ok = 0
    for ln in lines:
        if rng.random() < 0.6:
            ln = _mutate(rng, ln)
        if b"\x00" in ln:
            continue
        fl, ps, mq, isz = C.c_uint(), C.c_ulong(), C.c_uint(), C.c_int()
        rn, cg, sq = C.create_string_buffer(2048), C.create_string_buffer(2048), C.create_string_buffer(2048)
        want = olib.ora_parse_line(ln + b"\n", C.byref(fl), C.byref(ps), C.byref(mq), C.byref(isz), rn, cg, sq)
        got = lib.line2saml(ln + b"\n", C.byref(sp))
        assert got == (0 if want == 0 else 1), ln[:200]
        if want == 0:
            n_ok += 1
            paired = fl.value & 1
            assert (sp.flag, sp.pos, sp.mapq) == (fl.value, ps.value, mq.value)
            assert sp.isize == (isz.value if paired else len(sq.value))
            assert (sp.rname, sp.cigar, sp.seq, sp.seq_len) == (rn.value, cg.value, sq.value, len(sq.value))
            assert (sp.bits & 0xfff) == (fl.value & 0xfff)          # the twelve unpacked flag bits
    assert n_ok > 1000


def test_kmer_table_api(lib):
    for k in (1, 3, 8, 10):
        ks = lib.init_KSP(k)
        seq = Synth.genome(k, [5000], n_frac=0.02).seqs[0].tobytes()
        want = {}
        for i in range(len(seq) - k + 1):
            km = seq[i:i + k]
            rc = lib.add_to_ksp(seq[i:], ks)
            ok = all(c in b"ACGT" for c in km.upper())
            assert rc == (0 if ok else -1)
            if ok:
                want[km.upper()] = want.get(km.upper(), 0) + 1
        for km, c in list(want.items())[:200]:
            assert lib.kmer2count(km, ks) == c
            assert lib.kmer2count(km.lower(), ks) == c
        assert lib.kmer2count(b"N" * k, ks) == 0
        lib.destroy_KSP(ks)
    assert not lib.init_KSP(0)


@pytest.mark.parametrize("case", MAN["pss"][:6], ids=lambda c: c["sam"] + "".join(c["args"]))
def test_writers_reproduce_reference_files(lib, case):
    R = 15
    if "-r" in case["args"]:
        R = int(case["args"][case["args"].index("-r") + 1])
    gold_counts = open(os.path.join(GOLD, case["counts"]), "rb").read()
    gold_rates = open(os.path.join(GOLD, case["rates"]), "rb").read()
    fwd, rev = parse_counts_file(gold_counts, R)
    fr = np.zeros((R, 12)); rr = np.zeros((R, 12))
    lib.pss_sub_rates(fwd.ctypes.data, R, fr.ctypes.data)
    lib.pss_sub_rates(rev.ctypes.data, R, rr.ctypes.data)
    d = tmpdir()
    cwd = os.getcwd()
    os.chdir(d)
    try:
        sam = (case["sam"] + ".sam").encode()
        assert lib.pss_write_counts(b"genome.fa", sam, b"out", fwd.ctypes.data, rev.ctypes.data, R) == 0
        assert lib.pss_write_rates(b"genome.fa", sam, b"out", fr.ctypes.data, rr.ctypes.data, R) == 0
        assert open("out.pss.counts.txt", "rb").read() == gold_counts
        assert open("out.pss.rates.txt", "rb").read() == gold_rates
    finally:
        os.chdir(cwd)


def test_parallel_pread_reads_what_the_file_holds(tmp_path):
    """pss_io.c: the BAM file reader of the host programs (several pread()s at once) against plain reads: every
    thread count, offsets inside and at the end of the file, requests that run past the end."""
    so = os.path.join(ROOT, "tests", "host_emul", "libpssio.so")
    subprocess.run(["gcc", "-O2", "-g", "-std=gnu11", "-Wall", "-fPIC", "-shared", "-o", so, os.path.join(HOST, "pss_io.c"), "-lpthread"], check=True)
    h = C.CDLL(so)
    h.pss_pread_parallel.restype = C.c_size_t
    h.pss_pread_parallel.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int64, C.c_int]
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, size=5 * (1 << 20) + 12345, dtype=np.uint8).tobytes()
    fn = tmp_path / "blob"
    fn.write_bytes(data)
    fd = os.open(fn, os.O_RDONLY)
    try:
        buf = (C.c_char * (len(data) + 4096))()
        for threads in (0, 1, 2, 3, 4, 7, 16, 99):
            for off, n in ((0, len(data)), (0, len(data) + 4000), (1, 3 << 20), (len(data) - 5, 100), (len(data), 10), (777, 0),
                           ((1 << 20) - 1, (2 << 20) + 2)):
                got = h.pss_pread_parallel(fd, buf, n, off, threads)
                assert got == max(0, min(n, len(data) - off)), (threads, off, n, got)
                assert buf.raw[:got] == data[off:off + got], (threads, off, n)
    finally:
        os.close(fd)
