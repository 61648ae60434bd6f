set -x
nvidia-smi -L
nproc; free -g | head -2
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
python tools/quick_bench.py --genome-mb 200 --reads 4000000 2>&1 | tail -30
python tools/quick_bench.py --genome-mb 200 --reads 4000000 --config 1 --iters 3 2>&1 | tail -12
python tools/quick_bench.py --genome-mb 200 --reads 4000000 --mode fragkon --iters 3 2>&1 | tail -12
