"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multi.py under torchrun, one process per GPU,
NCCL).  Every rank tallies its newline-cut shard of the same seeded SAM text through the C ABI on its own GPU; the
tables are summed with an NCCL all-reduce; the spectrum is computed on this rank's slice of the packed genome and
all-reduced as well.  Rank 0 then compares with (a) the oracle and (b) the single-GPU result.  Exit status 0 = equal."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pss_testlib import FkParams, Oracle, PssParams, Synth, oracle_parallel, oracle_spectrum_parallel, reads_cfg_config2  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("pss-bam_b200")
    d = importlib.import_module("pss-bam_b200.dist")
    cores = max(1, (os.cpu_count() or 1) // world)
    Synth.set_threads(cores)

    g = Synth.genome(91, [60_000_000, 41_000_000, 9_000_000, 123_457], n_frac=0.01, lower_frac=0.03)
    sam = np.frombuffer(Synth.sam(reads_cfg_config2(seed=92), g, 0, 600_000), dtype=np.uint8)
    ctx = pkg.Context(local)
    ctx.upload_genome(list(zip(g.names, g.seqs)))
    lo, hi = d.shard_sam_bytes(sam, rank, world)
    mine = sam[lo:hi]
    dev = torch.device("cuda", local)

    def reduced(arrs):
        t = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs]).astype(np.int64)).to(dev)
        d.allreduce_tables(t)                                   # NCCL sum over NVLink
        return t.cpu().numpy().astype(np.uint64)

    R, K, k = 15, 8, 12
    f, r = ctx.pss(mine)
    st = ctx.stats()
    pss_all = reduced([f, r])
    fp, tp = ctx.fragkon(mine, pkg.FragkonOptions(klen=K))
    fk_all = reduced([fp, tp])
    # device-to-device: the shard's spectrum lands in a torch buffer and is all-reduced in place
    spec = torch.zeros(1 << (2 * k), dtype=torch.int64, device=dev)
    ctx.kmer_spectrum_device(k, spec.data_ptr(), rank, world)
    dist.all_reduce(spec)
    spec_all = spec.cpu().numpy().astype(np.uint64)
    stats_all = reduced([np.array([st[x] for x in ("lines", "counted", "no_contig", "filtered", "parse_fail", "undefined")], dtype=np.uint64)])

    ok = True
    if rank == 0:
        ora = Oracle(contigs=list(zip(g.names, g.seqs)))
        of, orv, ost = oracle_parallel(ora, sam, "pss", PssParams(), cores)
        ofp, otp, _ = oracle_parallel(ora, sam, "fragkon", FkParams(klen=K), cores)
        ospec = oracle_spectrum_parallel(ora, k, cores)
        checks = {
            "pss == oracle": np.array_equal(pss_all, np.concatenate([of.reshape(-1), orv.reshape(-1)])),
            "stats == oracle": stats_all.tolist() == [ost[x] for x in ("lines", "counted", "no_contig", "filtered", "parse_fail", "undefined")],
            "fragkon == oracle": np.array_equal(fk_all, np.concatenate([ofp, otp])),
            "spectrum k=12 == oracle": np.array_equal(spec_all, ospec),
        }
        sf, sr = ctx.pss(sam)                                   # the whole input on one GPU
        checks["pss == single GPU"] = np.array_equal(pss_all, np.concatenate([sf.reshape(-1), sr.reshape(-1)]))
        checks["spectrum == single GPU"] = np.array_equal(spec_all, ctx.kmer_spectrum(k))
        for name, v in checks.items():
            print(f"[multi-gpu world={world}] {name}: {'ok' if v else 'MISMATCH'}", flush=True)
        ok = all(checks.values())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
