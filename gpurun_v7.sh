set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter"
timeout 300 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 --config 1 2>&1 | grep -E "iter"
timeout 900 python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -c 2500 gpurun_out/bench_v7.json
