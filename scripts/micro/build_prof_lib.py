#!/usr/bin/env python
"""Build hybrid-rag-colbertv2_b200/libhrc_prof.so: a copy of the library whose tensor-core MaxSim kernel accounts
clock64() cycles per role (MMA warp: wait tempty / wait full / issue; epilogue warps: wait tfull / walk+math /
finish_doc) and prints the per-tile averages from CTA 8.  Development aid only; the product library is untouched.

    python scripts/micro/build_prof_lib.py && HRC_LIB_PATH=.../libhrc_prof.so python <anything that launches the kernel>
"""
import os
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "csrc")


def rep(s, a, b):
    assert s.count(a) == 1, (s.count(a), a)
    return s.replace(a, b)


src = open(os.path.join(CSRC, "maxsim_tc.cu")).read()
src = rep(src, "  const uint32_t tmem_base = *tmem_slot;",
          "  long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0, prof_e = 0;\n  int prof_docs = 0;\n"
          "  const long long prof_t0 = clock64();\n  const uint32_t tmem_base = *tmem_slot;")
src = rep(src, """        mbar_wait_wd(&tempty[ts], tphase ^ 1);
        mbar_wait_wd(&full[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {""", """        long long c0 = clock64();
        mbar_wait_wd(&tempty[ts], tphase ^ 1);
        long long c1 = clock64();
        mbar_wait_wd(&full[stage], phase);
        long long c2 = clock64();
        prof_a += c1 - c0; prof_b += c2 - c1;
        tc_fence_after_sync();
        if (elect_one()) {""")
src = rep(src, """        __syncwarp();
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
      }
    }
  } else if (ZP == 0""", """        __syncwarp();
        prof_c += clock64() - c2;
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        if (++ts == kTileStages) { ts = 0; tphase ^= 1; }
      }
      if (blockIdx.x == 8 && lane == 0)
        printf("MMA warp: tiles %d  per tile: wait tempty %lld  wait full %lld  issue+commit %lld  total %lld\\n", n_tiles,
               prof_a / n_tiles, prof_b / n_tiles, prof_c / n_tiles, (clock64() - prof_t0) / n_tiles);
    }
  } else if (ZP == 0""")
src = rep(src, """      mbar_wait_wd(&tfull[ts], tphase);
      tc_fence_after_sync();
      const int tile0 = t * TN, tile1 = tile0 + TN;""", """      long long c0 = clock64();
      mbar_wait_wd(&tfull[ts], tphase);
      long long c1 = clock64();
      prof_a += c1 - c0;
      tc_fence_after_sync();
      long long c2 = clock64();
      prof_d += c2 - c1;
      const int tile0 = t * TN, tile1 = tile0 + TN;""")
src = rep(src, "        if (e_tok <= tile1) finish_doc(); else break;",
          "        if (e_tok <= tile1) { long long f0 = clock64(); finish_doc(); prof_e += clock64() - f0; ++prof_docs; } else break;")
src = rep(src, """      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(&tempty[ts], 0);""", """      long long c3 = clock64();
      prof_b += c3 - c2;
      tc_fence_before_sync();
      __syncwarp();
      prof_c += clock64() - c3;
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(&tempty[ts], 0);""")
src = rep(src, "    while (have_doc) finish_doc();          // trailing empty documents",
          """    if (blockIdx.x == 8 && lane == 0 && n_tiles > 0 && (warp == 2 || warp == 6))
      printf("epi warp %d: per tile: wait tfull %lld  fence_after %lld  walk+math (incl finish) %lld  finish_doc %lld (docs %d, per doc %lld)  fence_before+syncwarp %lld  total %lld\\n",
             warp, prof_a / n_tiles, prof_d / n_tiles, prof_b / n_tiles, prof_e / n_tiles, prof_docs,
             prof_e / (prof_docs ? prof_docs : 1), prof_c / n_tiles, (clock64() - prof_t0) / n_tiles);
    while (have_doc) finish_doc();          // trailing empty documents""")

tmp = tempfile.mkdtemp(prefix="hrc_prof_")
for f in os.listdir(CSRC):
    if f.endswith((".cu", ".cuh")):
        shutil.copy(os.path.join(CSRC, f), tmp)
open(os.path.join(tmp, "maxsim_tc.cu"), "w").write(src.replace('#include "hrc_common.cuh"', '#include "hrc_common.cuh"'))
# the sources include ../../include/hrc.h relative to csrc/: mirror that layout
os.makedirs(os.path.join(tmp, "..", "..", "include"), exist_ok=True) if False else None
inc = os.path.join(ROOT, "include")
srcs = ["capi.cu", "maxsim_tc.cu", "maxsim_simt.cu", "meanpool.cu", "topk.cu", "rrf.cu", "synth.cu"]
for f in os.listdir(tmp):
    if f.endswith((".cu", ".cuh")):
        t = open(os.path.join(tmp, f)).read().replace('"../../include/hrc.h"', f'"{inc}/hrc.h"')
        open(os.path.join(tmp, f), "w").write(t)
out = os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_prof.so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-cudart", "static", "-o", out] +
               [os.path.join(tmp, f) for f in srcs], check=True)
shutil.rmtree(tmp)
print("built", out)
