set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 --config 1 2>&1 | grep -E "iter 2"
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 --mode fragkon 2>&1 | grep -E "iter 2"
timeout 300 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_v6 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
