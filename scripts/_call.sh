P=/root/repo/hybrid-rag-colbertv2_b200
for rep in 1 2 3; do for lib in libhrc.so libhrc_prev.so libhrc_new.so; do
  HRC_LIB_PATH=$P/$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench.log 2>&1; echo -n "$lib  "; python scripts/fmt_bench.py gpurun_out/ab_bench.log | cut -c1-120
done; done
