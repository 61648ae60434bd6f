#!/usr/bin/env python3
"""Host-to-device copy ceiling of a box, next to the library's own staging.

    python tools/h2d_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Every rank copies a pinned host buffer to its GPU -- as one cudaMemcpyAsync, in 64 MiB pieces (what pssgpu_feed
issues), and through pssgpu_feed itself (copy + tally) -- all ranks at the same time, timed with CUDA events, the
aggregate taken over the slowest rank.  Prints one JSON line (rank 0): the plain-copy figure is the box's ceiling
for any host-fed path; the gap to the pssgpu_feed figure is what the library's staging costs."""
import importlib
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    affinity = "default"
    if os.environ.get("PROBE_AFFINITY", "1") == "1":
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            affinity = "nvml"
        except Exception as ex:
            affinity = f"default ({type(ex).__name__})"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb = float(os.environ.get("PROBE_GB", "4"))
    n = int(gb * (1 << 30))
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(65)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            w0 = time.perf_counter()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return best

    piece = 64 << 20
    res = {}
    res["one_copy"] = timed(lambda: dev.copy_(host, non_blocking=True))
    res["pieces_64MiB"] = timed(lambda: [dev[o:o + piece].copy_(host[o:o + piece], non_blocking=True) for o in range(0, n, piece)])
    # the library's own staging: SAM-shaped lines so that the tally kernel has records to walk (they fail the parse early)
    pkg = importlib.import_module("pss-bam_b200")
    ctx = pkg.Context(local)
    ctx.upload_genome([("chr1", b"ACGT" * 4096)])
    line = b"r\t0\tchrZ\t100\t30\t50M\t*\t0\t0\t" + b"A" * 50 + b"\t" + b"I" * 50 + b"\tNM:i:0\n"
    reps = n // len(line)
    hv = host[:reps * len(line)].view(reps, len(line))
    hv[:] = torch.frombuffer(bytearray(line), dtype=torch.uint8)
    nfeed = reps * len(line)

    def feed():
        ctx.pss_begin(pkg.PssOptions())
        ctx.feed_ptr(host.data_ptr(), nfeed, last=True)
        ctx.sync()
    feed()
    res["pssgpu_feed"] = timed(feed)
    ctx.close()
    if rank == 0:
        out = {"n_gpus": world, "bytes_per_gpu": n, "cpu_affinity": affinity,
               "host_cores": os.cpu_count(),
               "gb_per_s_per_gpu": {k: (nfeed if k == "pssgpu_feed" else n) / (v * 1e-3) / 1e9 for k, v in res.items()},
               "gb_per_s_aggregate": {k: world * (nfeed if k == "pssgpu_feed" else n) / (v * 1e-3) / 1e9 for k, v in res.items()}}
        try:
            out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            out["topo"] = [ln for ln in topo.splitlines() if ln.startswith("GPU")][:8]
        except Exception:
            pass
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
