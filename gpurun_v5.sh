set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for v in 3 4; do
  echo "== variant c$v"
  PSSGPU_LIB=pss-bam_b200/lib/variants/libpssgpu_c$v.so python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
done
python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_v4 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
