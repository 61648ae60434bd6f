"""Developer probe for compute-sanitizer: one small pss-bam + fragkon + fused tally incl. malformed and long lines."""
import importlib, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pss_testlib import Synth, reads_cfg_config2
from test_record_logic import _mutate
pkg = importlib.import_module("pss-bam_b200")
g = Synth.genome(7, [300000, 200000, 5000], n_frac=0.01, lower_frac=0.05)
good = Synth.sam(reads_cfg_config2(seed=8), g, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 6000).split(b"\n")[:-1]
rng = random.Random(1)
lines = []
for i, ln in enumerate(good):
    lines.append(_mutate(rng, ln) if rng.random() < 0.05 else ln)
    if i % 1500 == 700: lines.append(b"z" * 70000)
    if i % 1500 == 900: lines.extend([b"a\tb"] * 50)
sam = b"\n".join(lines) + b"\n"
ctx = pkg.Context(0)
ctx.upload_genome(list(zip(g.names, g.seqs)))
f, r = ctx.pss(sam); print("pss", ctx.stats())
fp, tp = ctx.fragkon(sam, pkg.FragkonOptions(klen=8)); print("fragkon", int(fp.sum()), int(tp.sum()))
print("spectrum", int(ctx.kmer_spectrum(8).sum()), int(ctx.kmer_spectrum(5).sum()), int(ctx.kmer_spectrum(11).sum()))
ctx.close()
