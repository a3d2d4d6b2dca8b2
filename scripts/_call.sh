mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python scripts/fmt_bench.py gpurun_out/bench.log | cut -c1-400
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; python scripts/fmt_bench.py gpurun_out/bench2.log | cut -c1-400
