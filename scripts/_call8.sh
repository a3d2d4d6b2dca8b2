mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2953$n scripts/check_sharded_nccl.py > gpurun_out/nccl_check_$n.log 2>&1; echo "nccl check n=$n exit $?"; grep "world=" gpurun_out/nccl_check_$n.log
done
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench n=1 exit $?"; python scripts/fmt_bench.py gpurun_out/bench_n1.log | cut -c1-330
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n exit $?"; python scripts/fmt_bench.py gpurun_out/bench_n$n.log | cut -c1-330
done
timeout 500 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 --docs-per-gpu 1250000 > gpurun_out/bench_c5_n8.log 2>&1; echo "C5 (10M docs / 8 GPUs) exit $?"; python scripts/fmt_bench.py gpurun_out/bench_c5_n8.log | cut -c1-330
timeout 200 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_ref_n8.log 2>&1; echo "reference arm n=8 exit $?"; tail -1 gpurun_out/bench_ref_n8.log | cut -c1-200
