P=/root/repo/hybrid-rag-colbertv2_b200
cat > /tmp/one.py <<'PY'
import torch, sys
sys.path.insert(0,'/root/repo')
from hybrid_rag_colbertv2_b200 import _lib
from hybrid_rag_colbertv2_b200.synth import synth_store, synth_queries
dev=torch.device('cuda:0')
store=synth_store(300000,32,512,seed=12,device=dev)
q=synth_queries(64,32,device=dev)
out=torch.empty((64,store.n_docs),dtype=torch.float32,device=dev)
_lib.maxsim_scores(store.tokens,store.offsets,q,out=out); torch.cuda.synchronize()
PY
for pair in 1; do for dbg in 0 1 7; do echo "== PAIR=$pair DEBUG=$dbg"; HRC_LIB_PATH=$P/libhrc_prof.so HRC_TC_PAIR=$pair HRC_TC_DEBUG=$dbg python /tmp/one.py 2>&1 | sort | uniq -c | sort -k2 | head -12; done; done
