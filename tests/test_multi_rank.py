"""N > 1 host logic on the CPU: world_size-2 gloo process group, reads sharded at newline boundaries, per-rank
tables summed with all_reduce, compared with the single-process result.  (On GPUs the same code path runs over
NCCL; the per-rank tally itself is covered by the GPU parity tests.)"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pss_testlib import FkParams, Oracle, PssParams, Synth, reads_cfg_config2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fasta, sam, out):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("pss-bam_b200.dist")
    lo, hi = d.shard_sam_bytes(sam, rank, world)
    ora = Oracle(fasta=fasta)                               # stands in for the per-rank GPU tally
    f, r, st = ora.pss(sam[lo:hi], PssParams())
    fp, tp, _ = ora.fragkon(sam[lo:hi], FkParams(klen=6))
    t = torch.from_numpy(np.concatenate([f.reshape(-1), r.reshape(-1)]).astype(np.int64))
    k = torch.from_numpy(np.concatenate([fp, tp]).astype(np.int64))
    n = torch.tensor([st["lines"], st["counted"], hi - lo], dtype=torch.int64)
    d.allreduce_tables(t)
    d.allreduce_tables(k)
    d.allreduce_tables(n)
    g0, g1 = d.shard_range(1000, rank, world)
    span = torch.tensor([g1 - g0], dtype=torch.int64)
    d.allreduce_tables(span)
    if rank == 0:
        np.savez(out, t=t.numpy(), k=k.numpy(), n=n.numpy(), span=span.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_tally_equals_single_run(tmp_path, world):
    g = Synth.genome(51, [30000, 20000])
    fasta = g.fasta_bytes()
    sam = Synth.sam(reads_cfg_config2(seed=52, min_len=20, max_len=90), g, 0, 4000)
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, fasta, sam, out), nprocs=world, join=True)
    res = np.load(out)
    ora = Oracle(fasta=fasta)
    f, r, st = ora.pss(sam, PssParams())
    fp, tp, _ = ora.fragkon(sam, FkParams(klen=6))
    assert np.array_equal(res["t"], np.concatenate([f.reshape(-1), r.reshape(-1)]).astype(np.int64))
    assert np.array_equal(res["k"], np.concatenate([fp, tp]).astype(np.int64))
    assert res["n"].tolist() == [st["lines"], st["counted"], len(sam)]
    assert res["span"].tolist() == [1000]


def test_shard_sam_bytes_edge_cases():
    d = importlib.import_module("pss-bam_b200.dist")
    sam = b"aaaa\nbb\n\ncccccccccc\nd"
    for world in (1, 2, 3, 5, 8, 40):
        cuts = [d.shard_sam_bytes(sam, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == len(sam)
        for (a, b), (c, e) in zip(cuts, cuts[1:]):
            assert b == c and a <= b
        for a, b in cuts:
            assert a == 0 or sam[a - 1:a] == b"\n"
    assert d.shard_sam_bytes(b"", 0, 2) == (0, 0)
    assert [d.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
