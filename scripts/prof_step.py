"""One policy step at batch B (for ncu / quick timing):  python scripts/prof_step.py B precision [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpinets_b200 import scenes, _lib
from mpinets_b200.engine import Engine
from oracle import oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
prec = _lib.PRECISIONS[sys.argv[2]] if len(sys.argv) > 2 else _lib.PREC_FP32
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
print(sys.argv[1:], {k: round(v["ms"] / reps, 4) for k, v in st.items() if v["launches"]})
print("tc_error", eng.tc_error())
