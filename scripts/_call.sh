mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; python scripts/fmt_bench.py gpurun_out/bench.log | cut -c1-400
python scripts/bench_configs.py --configs c1,ragged,c3,c4 > gpurun_out/configs.log 2>&1; cut -c1-400 gpurun_out/configs.log
