set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
PSSGPU_TALLY_KERNEL=a timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for k in a b; do
  echo "== kernel $k"
  PSSGPU_TALLY_KERNEL=$k timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
  PSSGPU_TALLY_KERNEL=$k timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 --mode fragkon 2>&1 | grep -E "iter"
done
timeout 300 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_v5 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
