#!/bin/bash
# Run the GPU test groups in separate processes (a faulting kernel poisons its CUDA context), each
# under its own timeout, logging to gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run t_simt   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "simt and not candidates"
run t_select python -m pytest tests/test_gpu_parity.py -q -m gpu -k "topk or rrf or synthetic"
run t_tc     python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc and match_oracle"
run t_rest   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not match_oracle and not topk and not rrf and not synthetic"
run smoke    python -c "import __graft_entry__ as g; g.smoke()"
run bench    python bench.py --steps 5 --warmup 3
