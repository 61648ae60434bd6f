"""N > 1 on real GPUs: world_size-2 (and, when the box has them, 4 / 8) torchrun jobs over NCCL -- reads sharded at
newline boundaries, genome replicated, genome-sharded spectrum, tables all-reduced -- must equal the oracle and the
single-GPU result (pss-bam.c:401-420, fragkon.c:137-146, genome-kmer-count.c:56-64).  Skipped on a one-GPU box; the
CPU suite covers the same host logic over gloo (tests/test_multi_rank.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_equals_single_gpu_and_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 1000) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and "spectrum k=12 == oracle: ok" in r.stdout


def _group_case(devices, env_extra):
    """Python Group + the three CLIs over `devices`, against the oracle / the golden files."""
    import importlib
    import json
    import numpy as np
    from pss_testlib import FkParams, Oracle, PssParams, Synth, reads_cfg_config2, tmpdir
    os.environ.update(env_extra)
    try:
        pkg = importlib.import_module("pss-bam_b200")
        g = Synth.genome(93, [900_000, 400_000, 77_777], n_frac=0.01, lower_frac=0.03)
        ora = Oracle(fasta=g.fasta_bytes())
        sam = Synth.sam(reads_cfg_config2(seed=94), g, 0, 120_000)
        os.environ["PSSGPU_BAM_BATCH_MB"] = "1"               # read when a context first ingests BAM
        grp = pkg.Group(devices)
        assert grp.size == len(devices)
        grp.upload_genome(ora.contigs())
        lines = sam.split(b"\n")[:-1]
        f, r, st = ora.pss(sam, PssParams())
        grp.pss_begin(pkg.PssOptions())
        for i in range(0, len(lines), 7001):                   # pieces of whole lines, dealt in turn
            grp.feed(b"\n".join(lines[i:i + 7001]) + b"\n")
        gf, gr = grp.pss_finish()
        assert grp.stats() == st and np.array_equal(gf, f) and np.array_equal(gr, r)
        fp, tp, fst = ora.fragkon(sam, FkParams(klen=8))
        grp.fragkon_begin(pkg.FragkonOptions(klen=8))
        for i in range(0, len(lines), 9973):
            grp.feed(b"\n".join(lines[i:i + 9973]) + b"\n")
        gfp, gtp = grp.fragkon_finish()
        assert grp.stats() == fst and np.array_equal(gfp, fp) and np.array_equal(gtp, tp)
        for k in (4, 8, 12):                                  # genome-sharded spectrum + sum of the 4^k counters
            assert np.array_equal(grp.kmer_spectrum(k), ora.kmer_spectrum(k)), k
        # a BAM FILE over the group (pssgpu_group_feed_bam): batches of BGZF blocks are inflated by the members in turn,
        # member 0 fetches, frames, renders and tallies; 1 MiB batches so that this small file makes a dozen of them
        bam = Synth.bam(sam, list(zip(g.names, g.lens)) + [("chrUn_synthetic_decoy", 1000)], level=6, qual_mode=1)
        for chunk in (None, 700_001):
            grp.pss_begin(pkg.PssOptions())
            if chunk is None:
                grp.feed_bam(bam, last=True)
            else:
                for off in range(0, len(bam), chunk):
                    grp.feed_bam(bam[off:off + chunk])
                grp.feed_bam(b"", last=True)
            gf, gr = grp.pss_finish()
            assert grp.stats() == st and np.array_equal(gf, f) and np.array_equal(gr, r)
            info = grp.bam_info()
            per = info["batches_per_member"]
            assert info["records"] == 120_000 and sum(per) == info["batches"] - 1 and info["batches"] > 6, info
            assert per[1] > 0 and max(per) < sum(per), info       # dealt: nobody inflated everything
        grp.bam_read_group("no_such_group")
        grp.pss_begin(pkg.PssOptions())
        grp.feed_bam(bam, last=True)
        gf, gr = grp.pss_finish()
        assert gf.sum() == 0 and grp.bam_info()["dropped_by_read_group"] == 120_000
        grp.bam_read_group(None)
        bad = bytearray(bam)                                  # a damaged block inflated on another GPU is reported here
        bad[len(bam) // 2] ^= 0x10
        grp.pss_begin(pkg.PssOptions())
        with pytest.raises(pkg.PssGpuError):
            grp.feed_bam(bytes(bad), last=True)
            grp.pss_finish()
        grp.pss_begin(pkg.PssOptions())
        grp.feed_bam(bam, last=True)
        gf, gr = grp.pss_finish()
        assert grp.stats() == st and np.array_equal(gf, f)
        backend = grp.reduce_backend
        grp.close()
    finally:
        for k in env_extra:
            os.environ.pop(k, None)
        os.environ.pop("PSSGPU_BAM_BATCH_MB", None)
    # the host programs with $PSSGPU_DEVICES: byte-identical golden outputs
    gold = os.path.join(ROOT, "tests", "golden", "v1")
    man = json.load(open(os.path.join(gold, "manifest.json")))
    bindir = os.path.join(ROOT, "pss-bam_b200", "bin")
    subprocess.run(["make", "-C", os.path.join(ROOT, "pss-bam_b200", "host")], check=True, capture_output=True)
    d = tmpdir()
    shim = os.path.join(d, "samtools")
    with open(os.path.join(ROOT, "oracle", "samtools_shim.sh")) as fi, open(shim, "w") as fo:
        fo.write(fi.read())
    os.chmod(shim, 0o755)
    e = dict(os.environ, PATH=d + os.pathsep + os.environ.get("PATH", ""), PSSGPU_DEVICES=",".join(str(x) for x in devices), **env_extra)
    for fn in [man["fasta"], *man["sams"].values()]:
        os.symlink(os.path.join(gold, fn), os.path.join(d, fn))
    rd = lambda name: open(os.path.join(gold, name), "rb").read()
    for case in man["pss"][:5]:
        rr = subprocess.run([os.path.join(bindir, "pss-bam"), "-F", "genome.fa", "-B", case["sam"] + ".sam", "-o", "out", *case["args"]],
                            cwd=d, env=e, capture_output=True)
        assert rr.returncode == 0, rr.stderr[-2000:]
        assert open(os.path.join(d, "out.pss.counts.txt"), "rb").read() == rd(case["counts"])
        assert open(os.path.join(d, "out.pss.rates.txt"), "rb").read() == rd(case["rates"])
    for case in [c for c in man["fragkon"] if not c["sparse"]][:3]:
        rr = subprocess.run([os.path.join(bindir, "fragkon"), "-F", "genome.fa", "-B", case["sam"] + ".sam", *case["args"]],
                            cwd=d, env=e, capture_output=True)
        assert rr.returncode == 0 and rr.stdout == rd(case["out"]), rr.stderr[-2000:]
    for case in man["gkc"]:
        rr = subprocess.run([os.path.join(bindir, "genome-kmer-count"), "-f", "genome.fa", "-k", str(case["k"])], cwd=d, env=e, capture_output=True)
        assert rr.returncode == 0 and rr.stdout == rd(case["out"]), rr.stderr[-2000:]
    # the host program reading a BAM FILE with the group (its bytes go to the GPUs) against the samtools-shim text path
    # on one GPU: same file name in two directories, so that the outputs must be byte-identical
    d1, d2 = tmpdir(), tmpdir()
    for dd in (d1, d2):
        with open(os.path.join(dd, "genome.fa"), "wb") as fo:
            fo.write(g.fasta_bytes())
    with open(os.path.join(d1, "reads.bam"), "wb") as fo:
        fo.write(sam)                                         # (the shim `cat`s it)
    with open(os.path.join(d2, "reads.bam"), "wb") as fo:
        fo.write(bam)
    e1 = dict(e, PSSGPU_DEVICES=str(devices[0]))
    e2 = dict(e, PSSGPU_BAM_BATCH_MB="1")
    for dd, ee in ((d1, e1), (d2, e2)):
        rr = subprocess.run([os.path.join(bindir, "pss-bam"), "-F", "genome.fa", "-B", "reads.bam", "-o", "out"], cwd=dd, env=ee, capture_output=True)
        assert rr.returncode == 0, rr.stderr[-2000:]
    for name in ("out.pss.counts.txt", "out.pss.rates.txt"):
        assert open(os.path.join(d1, name), "rb").read() == open(os.path.join(d2, name), "rb").read(), name
    return backend


def test_group_of_member_contexts_on_one_gpu():
    """The group code path -- dealing, per-member tallies, the sum (peer copies + add kernel), genome-sharded spectrum,
    the CLIs with $PSSGPU_DEVICES -- with three member contexts on GPU 0 (PSSGPU_GROUP_ALLOW_DUP), so that a one-GPU
    box covers it too."""
    assert "peer" in _group_case([0, 0, 0], {"PSSGPU_GROUP_ALLOW_DUP": "1"})


@pytest.mark.parametrize("reduce", ["nccl", "peer"])
def test_group_on_two_gpus(reduce):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    backend = _group_case([0, 1], {"PSSGPU_GROUP_REDUCE": reduce} if reduce == "peer" else {})
    assert (reduce in backend) or reduce == "nccl"          # (without libnccl.so.2 the peer path takes over)
