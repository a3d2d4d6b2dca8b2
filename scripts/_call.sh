mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
python - <<'PY'
import torch, time
import hybrid_rag_colbertv2_b200 as hrc
from hybrid_rag_colbertv2_b200 import _lib
from hybrid_rag_colbertv2_b200.synth import synth_store, synth_queries
dev=torch.device('cuda:0')
store=synth_store(1_000_000,128,128,seed=20260102,device=dev)
q=synth_queries(1,32,device=dev)
out=torch.empty((1,store.n_docs),dtype=torch.float32,device=dev)
for _ in range(3): _lib.meanpool_cosine_scores(store.tokens,store.offsets,q,out=out)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _lib.meanpool_cosine_scores(store.tokens,store.offsets,q,out=out)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('meanpool_cosine C2: ms',ms,'GB/s',store.total_tokens*256/ms/1e6)
PY
