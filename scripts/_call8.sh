mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2953$n scripts/check_sharded_nccl.py > gpurun_out/nccl_check_$n.log 2>&1; echo "nccl check n=$n exit $?"; grep "world=" gpurun_out/nccl_check_$n.log
done
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n exit $?"; python scripts/fmt_bench.py gpurun_out/bench_n$n.log | cut -c1-330
done
HRC_TC_M64=1 timeout 400 $TR --nproc-per-node 8 --master-port 29550 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_m64.log 2>&1; echo "bench n=8 M64 exit $?"; python scripts/fmt_bench.py gpurun_out/bench_n8_m64.log | cut -c1-330
timeout 500 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 --docs-per-gpu 1250000 > gpurun_out/bench_c5_n8.log 2>&1; echo "C5 (10M docs / 8 GPUs) exit $?"; python scripts/fmt_bench.py gpurun_out/bench_c5_n8.log | cut -c1-330
nvidia-smi --query-gpu=index,power.limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpus8.txt; cat gpurun_out/gpus8.txt
