set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python tools/quick_bench.py --genome-mb 200 --reads 4000000 --iters 3 2>&1 | grep -E "iter|host feed"
( time python bench.py ) > gpurun_out/bench_full.log 2>&1
tail -c 3000 gpurun_out/bench_full.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
tail -c 1500 gpurun_out/bench_ref.log
