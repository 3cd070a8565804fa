"""GEMM tuning sweep on a B200: for each knob setting (environment, read at context creation) time the conv stack on one
synthetic channel and print per-role cycle counters.  Usage: python tools/gemm_sweep.py [minutes] ; not part of the product."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laughter_detection_icsi_b200 import synth  # noqa: E402
from laughter_detection_icsi_b200.engine import Engine  # noqa: E402

NAMES = ["prod_wait", "mma_wait_full", "mma_wait_acc", "mma_issue", "epi_wait", "epi_work", "cta", "tiles"]


def run(label, env, minutes, detail=False, chunk_rows=0):
    for k in ("LD_GEMM_SPAN", "LD_GEMM_STAGES", "LD_GEMM_PROF", "LD_GEMM_STAGE_BYTES", "LD_GEMM_MAX_OUTS", "LD_GEMM_RINGS", "LD_GEMM_DBG",
              "LD_GEMM_ISSUERS_WIDE", "LD_GEMM_ISSUERS_NARROW", "LD_GEMM_PIPE", "LD_GEMM_PIPE_W", "LD_GEMM_PIPE_LEAD", "LD_GEMM_PIPE_DBG", "LD_GEMM_PIPE_ROLES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = Engine(0, chunk_rows=chunk_rows)
    eng.load_state_dict(synth.synthetic_state_dict())
    T = int(minutes * 6000)
    feats = torch.randn(T, 44, device="cuda") * 3 - 4
    for _ in range(2):
        eng.infer_windows(feats)
    torch.cuda.synchronize()
    eng.timing_read(reset=True); eng.timing_read_convs(reset=True); eng.gemm_counters(reset=True)
    eng.timing_enable(True)
    reps = 3
    for _ in range(reps):
        eng.infer_windows(feats)
    torch.cuda.synchronize()
    eng.timing_enable(False)
    convs = eng.timing_read_convs(reset=True)
    t = eng.timing_read(reset=True)
    total = sum(v for _, v in convs) / reps
    print(f"== {label}: conv stack {total:.2f} ms per {minutes:g} min channel (stem {t['stem'][0] / reps:.2f} head {t['head'][0] / reps:.2f}) "
          f"-> {minutes / 60 / (total / 1e3):.2f} audio-h/s conv-only", flush=True)
    if detail:
        sync = eng.gemm_sync_wait()
        spread = eng.gemm_cta_spread()
        cnt = eng.gemm_counters()
        groups = eng.conv_pipeline_groups()
        for name, ms in convs:
            print(f"   {name:22s} {ms / reps:7.3f} ms")
        for i, (name, c) in enumerate(cnt):
            if c[6]:
                cta = c[6]
                print(f"      {name:22s} group {groups[i][0]:2d} ctas {groups[i][1]:3d}  " + " ".join(f"{n}={c[k] / cta:5.2f}" for k, n in enumerate(NAMES[:6]))
                      + f"  sync_wait={sync[i][0] / cta:5.2f} (up {sync[i][1] / cta:5.2f})  cyc/tile={cta / max(c[7], 1):7.0f} (CTA min {spread[i][0]:6.0f} max {spread[i][1]:6.0f})")
    eng.close()
    del eng
    torch.cuda.empty_cache()


if __name__ == "__main__":
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    if not (len(sys.argv) > 2 and sys.argv[2] == "pipe"):
        run("default", {"LD_GEMM_PROF": "1"}, minutes, detail=True)
    if len(sys.argv) > 2 and sys.argv[2] == "dbg":
        for bits, what in ((2, "no MMAs"), (3, "no MMAs, one copy per stage"), (6, "no MMAs, no stores"), (14, "no MMAs, no stores, no TMEM reads"),
                           (15, "barrier skeleton only"), (4, "no stores"), (1, "one copy per stage")):
            run(f"LD_GEMM_DBG={bits}: {what} (garbage results)", {"LD_GEMM_PROF": "1", "LD_GEMM_DBG": str(bits)}, minutes, detail=bits in (15, 14))
    elif len(sys.argv) > 2 and sys.argv[2] == "knobs":
        for env in ({"LD_GEMM_ISSUERS_WIDE": "4"}, {"LD_GEMM_ISSUERS_NARROW": "2"}, {"LD_GEMM_STAGE_BYTES": "9000"},
                    {"LD_GEMM_STAGE_BYTES": "36000"}, {"LD_GEMM_MAX_OUTS": "2"}, {"LD_GEMM_MAX_OUTS": "3"}, {"LD_GEMM_SPAN": "0"}):
            run(" ".join(f"{k}={v}" for k, v in env.items()), dict(env, LD_GEMM_PROF="1"), minutes, detail=True)
    elif len(sys.argv) > 2 and sys.argv[2] == "chunks":
        for c in (2048, 4096, 8192, 16384, 65536):
            run(f"chunk_rows={c}", {}, minutes, detail=(c == 4096), chunk_rows=c)
    elif len(sys.argv) > 2 and sys.argv[2] == "pipe":
        run("LD_GEMM_PIPE=0", {"LD_GEMM_PROF": "1", "LD_GEMM_PIPE": "0"}, minutes, detail=False)
        for extra in sys.argv[3:]:
            env = dict(kv.split("=", 1) for kv in extra.split(";"))
            run(extra, dict(env, LD_GEMM_PROF="1"), minutes, detail=True)
    else:
        run("one ring", {"LD_GEMM_RINGS": "1"}, minutes)
