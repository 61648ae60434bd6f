set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python tools/quick_bench.py --genome-mb 200 --reads 4000000 --iters 3 2>&1 | tail -12
python tools/quick_bench.py --genome-mb 200 --reads 4000000 --mode fragkon --iters 3 2>&1 | grep iter
python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_v2 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
python bench.py --reads-per-gpu 2000000 --genome-scale 0.05 --steps 3 --warmup 1 --cpu-sample-reads 200000 --cpu-genome-mb 20 2>&1 | tail -5
