mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; python scripts/fmt_bench.py gpurun_out/bench.log | cut -c1-400
python scripts/bench_configs.py --configs c1,ragged,c4 > gpurun_out/configs.log 2>&1; cut -c1-330 gpurun_out/configs.log
