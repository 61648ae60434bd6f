"""The per-record logic the CUDA kernels run (pss-bam_b200/csrc/pss_record.h), compiled for the host by
tests/host_emul and fuzzed against the CPU oracle -- which itself is pinned to the reference binaries by
test_oracle_golden.py / test_oracle_vs_ref.py.  Covers what a GPU run of the parity tests would be slow to explore:
glibc sscanf corner cases, malformed lines, option combinations."""
import ctypes as C
import random

import numpy as np
import pytest

from pss_testlib import Emul, FkParams, Oracle, PssParams, Synth, reads_cfg_config1, reads_cfg_config2


@pytest.fixture(scope="module")
def world():
    g = Synth.genome(21, [60000, 45000, 700, 60, 33], names=["chr1", "chr2", "chrM", "tiny", "x" * 40],
                     n_frac=0.02, lower_frac=0.05)
    g.seqs[1][500:520] = np.frombuffer(b"RYKMSWBDHVNrykmswbdh", dtype=np.uint8)
    g.seqs[1][600:604] = np.frombuffer(b"*-.X", dtype=np.uint8)          # "other" symbols (exception list)
    ora = Oracle(fasta=g.fasta_bytes())
    em = Emul(ora.contigs())
    yield g, ora, em
    em.close()
    ora.close()


def _same_pss(ora, em, sam, p=PssParams()):
    f, r, st, status = ora.pss(sam, p, want_status=True)
    for slow in (0, 1):
        ef, er, est, fast = em.pss(sam, p, force_slow=slow)
        bad = np.flatnonzero(status != est)
        assert bad.size == 0, (slow, bad[:5], status[bad[:5]], est[bad[:5]], sam.split(b"\n")[int(bad[0])][:200])
        assert np.array_equal(f, ef) and np.array_equal(r, er), slow
    return st


def _same_fk(ora, em, sam, p):
    fp, tp, st, status = ora.fragkon(sam, p, want_status=True)
    for slow in (0, 1):
        efp, etp, est, _ = em.fragkon(sam, p, force_slow=slow)
        assert np.array_equal(status, est), slow
        assert np.array_equal(fp, efp) and np.array_equal(tp, etp), slow


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_reads_default_options(world, seed):
    g, ora, em = world
    st = _same_pss(ora, em, Synth.sam(reads_cfg_config2(seed=seed, min_len=20, max_len=160), g, 0, 20000))
    assert st["counted"] > 5000 and st["filtered"] > 1000 and st["parse_fail"] > 50 and st["no_contig"] > 20
    _same_pss(ora, em, Synth.sam(reads_cfg_config1(seed=seed), g, 0, 5000))


@pytest.mark.parametrize("p", [
    PssParams(region_len=0), PssParams(region_len=1), PssParams(region_len=4), PssParams(region_len=7), PssParams(region_len=16),
    PssParams(region_len=29), PssParams(region_len=30), PssParams(region_len=31), PssParams(region_len=47),
    PssParams(region_len=150), PssParams(min_len=35, max_len=90), PssParams(min_mq=37),
    PssParams(min_mq=-1), PssParams(up_ctx=b"C", down_ctx=b"G"), PssParams(up_ctx=b"ACGTN", down_ctx=b"acgt"),
    PssParams(up_ctx=b"RYKM*", down_ctx=b"ACGT-X."), PssParams(merged_only=1), PssParams(up_ctx=b"", down_ctx=b"ACGT"),
], ids=repr)
def test_options(world, p):
    g, ora, em = world
    _same_pss(ora, em, Synth.sam(reads_cfg_config2(seed=9, min_len=20, max_len=120), g, 0, 8000), p)


@pytest.mark.parametrize("K", list(range(1, 15)))
def test_fragkon_every_k(world, K):
    g, ora, em = world
    sam = Synth.sam(reads_cfg_config2(seed=30 + K, min_len=5, max_len=80), g, 0, 6000)
    _same_fk(ora, em, sam, FkParams(klen=K))
    _same_fk(ora, em, sam, FkParams(klen=K, min_len=20, max_len=60, min_mq=15, merged_only=K & 1))


def _mutate(rng, ln: bytes) -> bytes:
    f = ln.split(b"\t")
    k = rng.randrange(28)
    if k == 0: return ln.replace(b"\t", b" ")
    if k == 1: return ln.replace(b"\t", b" \t ", rng.randrange(1, 4))
    if k == 2: return b" \t" + ln
    if k == 3: f[1] += rng.choice([b"x", b"abc", b".5", b"e3"])
    elif k == 4: f[3] = rng.choice([b"+", b"-", b"0", b"00"]) + f[3]
    elif k == 5: f[8] = rng.choice([b"0x1e", b"0X1E", b"036", b"-036", b"0x", b"08", b"+30", b"-0", b"0x7fffffff", b"-2147483648"]); f[1] = b"99"
    elif k == 6: return ln + b"\r"
    elif k == 7: return b"@SQ\tSN:chr1\tLN:60000"
    elif k == 8: return b""
    elif k == 9: return b"\t".join(f[:rng.randrange(1, 11)])
    elif k == 10: f[4] = rng.choice([b"4294967296", b"99999999999999999999999", b"-1", b"255x"])
    elif k == 11: f[3] = rng.choice([b"18446744073709551615", b"18446744073709551616", b"-1", b"999999999999999999", b"1e3"])
    elif k == 12: f[7] = rng.choice([b"*", b"-", b"12345678901234567890123", b"0"])
    elif k == 13: f[5] = rng.choice([b"*", b"M", b"30", b"030M", b"+30M", b"30M1", b"0M", b"3" + f[5]])
    elif k == 14: f[9] = f[9][:-1]
    elif k == 15: f[10] = b"*"
    elif k == 16: f[2] = rng.choice([b"chr", b"chr11", b"*", b"tiny", b"x" * 40, b"x" * 41, b"CHR1"])
    elif k == 17: f[1] = str(rng.randrange(0, 4096)).encode()
    elif k == 18: f[3] = str(rng.choice([0, 1, 2, 3, 4, 5, 59960, 59990, 60000, 60001])).encode()
    elif k == 19: f[9] = f[9].lower()
    elif k == 20: f[9] = f[9].replace(b"A", b"N", 2)
    elif k == 21: f[0] = b"q" * rng.choice([2047, 2048, 3000])
    elif k == 22: f[9] = b"A" * 2048; f[10] = b"I" * 2048
    elif k == 23: return ln.replace(b"\t", b"\x0b", 1)
    elif k == 24: return ln.replace(b"\t", b"\x01", 1)
    elif k == 25: return ln[:len(ln) // 2] + b"\x00" + ln[len(ln) // 2:]
    elif k == 26: f[1] = b"99"; f[8] = str(rng.choice([len(f[9]), -len(f[9]), len(f[9]) + 1, 5, 10 ** 6, 10 ** 6 + 1, 2 * 10 ** 9])).encode()
    elif k == 27: f[2] = b"tiny"; f[3] = str(rng.randrange(1, 40)).encode(); f[9] = f[9][:20]; f[10] = f[10][:20]; f[5] = b"20M"
    return b"\t".join(f)


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_malformed_lines(world, seed):
    g, ora, em = world
    rng = random.Random(seed)
    good = Synth.sam(reads_cfg_config2(seed=seed, min_len=20, max_len=70), g, 0, 3000).split(b"\n")[:-1]
    lines = [_mutate(rng, ln) if rng.random() < 0.7 else ln for ln in good]
    sam = b"\n".join(lines) + (b"\n" if seed & 1 else b"")          # with and without a final newline
    st = _same_pss(ora, em, sam)
    assert st["parse_fail"] > 100 and st["undefined"] > 5
    _same_fk(ora, em, sam, FkParams(klen=7))
    _same_fk(ora, em, sam, FkParams(klen=4))


def test_overlong_lines_split_like_fgets(world):
    g, ora, em = world
    good = Synth.sam(reads_cfg_config1(seed=5), g, 0, 20).split(b"\n")[:-1]
    # a 200000-byte line, a 200001-byte line and a 450000-byte line whose tail is a good record
    lines = [good[0], b"x" * 199999, good[1], b"y" * 200000, good[2], b"z" * 199990 + b" " + good[3] + b"\t" + b"t" * 250000, good[4]]
    sam = b"\n".join(lines) + b"\n"
    _same_pss(ora, em, sam)


def test_scan11_matches_glibc(world):
    """emul_scan11 (the sscanf restatement the kernels fall back to) against glibc's sscanf, value by value."""
    _, ora, _ = world
    olib, elib = Oracle.lib(), Emul.lib()
    olib.ora_parse_line.argtypes = [C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_ulong), C.POINTER(C.c_uint),
                                    C.POINTER(C.c_int), C.c_char_p, C.c_char_p, C.c_char_p]
    rng = random.Random(77)
    nums = [b"0", b"7", b"16", b"99", b"-5", b"+5", b"007", b"0x1F", b"0X1f", b"0x", b"0xg", b"08", b"010", b"-010",
            b"4294967295", b"4294967296", b"-4294967296", b"2147483647", b"2147483648", b"-2147483648", b"-2147483649",
            b"9223372036854775807", b"9223372036854775808", b"-9223372036854775808", b"-9223372036854775809",
            b"18446744073709551615", b"18446744073709551616", b"-18446744073709551615", b"-18446744073709551616",
            b"123456789012345678901234567890", b"12a", b"1-2", b"--1", b"+-1", b"-", b"+", b"1.5", b"1e5", b"0b1", b"\xd9\xa1"]
    toks = [b"r1", b"chr1", b"30M", b"*", b"=", b"ACGT", b"IIII", b"AC", b"I"]
    seps = [b"\t", b" ", b"\t\t", b" \t", b"\x0b", b"\x0c", b"\r"]
    n_ok = 0
    for it in range(6000):
        f = [rng.choice(toks), rng.choice(nums), rng.choice(toks), rng.choice(nums), rng.choice(nums), rng.choice(toks),
             rng.choice(toks), rng.choice(nums), rng.choice(nums), rng.choice([b"ACGT", b"AC"]), rng.choice([b"IIII", b"II"])]
        if it % 3 == 0:     # mostly-clean lines so that many conversions succeed
            f[1], f[3], f[4], f[7] = b"16", b"100", b"30", b"0"
        line = b"".join(x + rng.choice(seps if rng.random() < 0.3 else [b"\t"]) for x in f[:rng.choice([11, 11, 11, 10, 12])])
        line = line.rstrip(b"\t") + rng.choice([b"", b"\n", b"\tNM:i:0\n"])
        fl, ps, mq, isz = C.c_uint(), C.c_ulong(), C.c_uint(), C.c_int()
        rn, cg, sq = C.create_string_buffer(2048), C.create_string_buffer(2048), C.create_string_buffer(2048)
        want = olib.ora_parse_line(line, C.byref(fl), C.byref(ps), C.byref(mq), C.byref(isz), rn, cg, sq)
        efl, eps, emq, etl = C.c_uint32(), C.c_uint64(), C.c_uint32(), C.c_int32()
        offs = (C.c_int32 * 6)()
        got = elib.emul_scan11(line, len(line), C.byref(efl), C.byref(eps), C.byref(emq), C.byref(etl), offs)
        code = {0: 0, 1: -2, 2: -3}[want]
        assert got == code, (line, want, got)
        if want == 0:
            n_ok += 1
            assert (efl.value, eps.value, emq.value, etl.value) == (fl.value, ps.value, mq.value, isz.value), line
            assert line[offs[0]:offs[0] + offs[1]] == rn.value and line[offs[2]:offs[2] + offs[3]] == cg.value
            assert line[offs[4]:offs[4] + offs[5]] == sq.value
    assert n_ok > 500
