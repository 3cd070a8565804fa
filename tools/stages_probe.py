import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import gemm_sweep
for st in ("16", "6", "4"):
    gemm_sweep.run(f"LD_GEMM_STAGES={st}", {"LD_GEMM_STAGES": st, "LD_GEMM_PROF": "1"}, 10.0, detail=True)
