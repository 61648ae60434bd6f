set -x
python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/qb_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_tally_v1 python tools/quick_bench.py --genome-mb 200 --reads 1000000 --iters 2 > gpurun_out/ncu_tally.log 2>&1
tail -5 gpurun_out/ncu_tally.log
