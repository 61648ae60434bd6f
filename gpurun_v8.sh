set -x
for v in match match3; do
  echo "== variant $v"
  PSSGPU_LIB=pss-bam_b200/lib/variants/libpssgpu_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
  PSSGPU_LIB=pss-bam_b200/lib/variants/libpssgpu_$v.so timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
done
echo "== default"
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
PSSGPU_LIB=pss-bam_b200/lib/variants/libpssgpu_match.so timeout 300 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
PSSGPU_LIB=pss-bam_b200/lib/variants/libpssgpu_match.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_match python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
