set -x
for f in pss-bam_b200/lib/variants/*.so; do
  echo "== $f"
  PSSGPU_LIB=$f timeout 200 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 2>&1 | grep -E "iter 2"
  PSSGPU_LIB=$f timeout 200 python tools/quick_bench.py --genome-mb 3000 --reads 8000000 --iters 3 --config 1 2>&1 | grep -E "iter 2"
done
# traffic + launch list of the bench launch (default build)
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.json 2>gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tally -s 1 -c 1 -o gpurun_out/prof_bench_tally python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_tally.log 2>&1
