"""One policy step at batch B (for ncu / quick timing):  python scripts/prof_step.py B precision [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpinets_b200 import scenes, _lib
from mpinets_b200.engine import Engine
from oracle import oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
prec = _lib.PREC_BF16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else _lib.PREC_FP32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = Engine()
eng.load_state_dict(O.reference_state_dict(0))
p = scenes.config_problems(2, B)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sc = {k: dev(p[k]) for k in scenes.SCENE_KEYS}
q0, tg = dev(p["q0"]), dev(p["target"])
cloud = eng.build_cloud(sc, q0, tg)
qn = eng.normalize(q0)
eng.profile(True)
for _ in range(reps):
    dq = eng.policy_forward(cloud, qn, prec)
torch.cuda.synchronize()
st = eng.profile_read()
print({k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in st.items() if v["launches"]})
print("tc_error", eng.tc_error())
if os.environ.get("MPN_TC_TIMELINE"):
    tl = eng.tc_timeline()
    n2 = ["top/write rows", "prefetch(bq+loads)", "-", "fence+sync", "issue L1", "wait L1", "ep1", "fence+sync ", "issue L2", "wait L2",
          "ep2+fence+sync", "issue L3", "wait L3", "ep3", "sync"]
    n = 64 * reps
    print("SA2 cycles/centroid (CTA0/WG0):", {k: int(v / n) for k, v in zip(n2, tl[:15])}, "total", int(sum(tl[:16]) / n))
    n1 = ["top", "ball query", "gather", "fence+sync", "issue L1", "wait(x2)", "ep(x2)", "fence+sync(x2)", "issue(x2)", "wait L3", "pool"]
    n = 128 * reps
    print("SA1 cycles/centroid (CTA0/WG0):", {k: int(v / n) for k, v in zip(n1, tl[16:27])}, "total", int(sum(tl[16:32]) / n))
    ng = ["prologue", "wait loads", "fence+sync", "mma issue", "wait prev mma", "issue loads", "drain", "epilogue"]
    k = max(tl[40], 1)
    print("GEMM cycles per CTA-0 tile (all row-GEMM launches):", {a: int(v / k) for a, v in zip(ng, tl[32:40])}, "tiles", tl[40])
