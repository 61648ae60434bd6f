#!/usr/bin/env python3
"""One-off timing of the UNMODIFIED reference (oracle/_ref/pss-bam) on the full configs[1] genome (BASELINE.md "Scale":
a >= 10 M-read prefix, extrapolated linearly).  P single-threaded processes, each loading the whole 3.1 Gb FASTA
(the reference has no threads and no shared genome) and tallying its own shard of the first N reads of the 200 M-read
set; load time is measured the same way with an empty SAM and reported next to the total.  Writes one JSON object.

    python tools/reference_fullscale.py [--reads 12000000] [--procs 12] > gpurun_out/r2_reference_fullscale.json
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bench
    from pss_testlib import Synth, reads_cfg_config2
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=12_000_000)
    ap.add_argument("--procs", type=int, default=min(12, os.cpu_count() or 1))
    ap.add_argument("--binary", default="pss-bam")
    a = ap.parse_args()
    d = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(d, a.binary)
    Synth.set_threads(os.cpu_count() or 1)
    plan = bench.contig_plan(1.0)
    g = Synth.genome(bench.GENOME_SEED, [l for _, l in plan], names=[n for n, _ in plan], n_frac=0.01, lower_frac=0.03)
    work = tempfile.mkdtemp(prefix="pssref_full_")
    try:
        with open(os.path.join(work, "genome.fa"), "wb") as f:
            f.write(g.fasta_bytes())
        cfg = reads_cfg_config2(seed=bench.READS_SEED)
        per = a.reads // a.procs
        for p in range(a.procs):
            with open(os.path.join(work, f"s{p}.sam"), "wb") as f:
                f.write(Synth.sam(cfg, g, p * per, (p + 1) * per))
        open(os.path.join(work, "empty.sam"), "wb").close()
        env = dict(os.environ)
        env["PATH"] = d + os.pathsep + env.get("PATH", "")

        def run(sams):
            t0 = time.perf_counter()
            ps = [subprocess.Popen([exe, "-F", "genome.fa", "-B", s, "-o", f"o{i}"], cwd=work, env=env,
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i, s in enumerate(sams)]
            rc = [p.wait() for p in ps]
            assert not any(rc), rc
            return time.perf_counter() - t0
        load = run(["empty.sam"] * a.procs)
        total = run([f"s{p}.sam" for p in range(a.procs)])
        reads = per * a.procs
        print(json.dumps({"what": f"unmodified reference {a.binary} (stock flags unless -O2 is in the name), {a.procs} processes x {per} reads "
                                  "of the configs[1] read set on the full 3.1 Gb / 24-contig genome",
                          "reads": reads, "procs": a.procs, "host_cores": os.cpu_count(),
                          "total_s": total, "fasta_load_s": load, "tally_s": total - load,
                          "reads_per_s_net_of_load": reads / max(total - load, 1e-9),
                          "reads_per_s_per_core": reads / max(total - load, 1e-9) / a.procs,
                          "reads_per_s_including_load": reads / total,
                          "extrapolated_200M_reads_s": 200e6 / (reads / max(total - load, 1e-9)),
                          "extrapolated_1B_reads_s": 1e9 / (reads / max(total - load, 1e-9))}))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
