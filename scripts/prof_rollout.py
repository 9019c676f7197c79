"""Cloud build + a short lock-step rollout at batch B (for ncu captures of the HBM-side kernels):
python scripts/prof_rollout.py B steps [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpinets_b200 import scenes, _lib
from mpinets_b200.engine import Engine
from oracle import oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = Engine()
eng.load_state_dict(O.reference_state_dict(0))
p = scenes.config_problems(cfg, B)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sc = {k: dev(p[k]) for k in scenes.SCENE_KEYS}
q0, tg = dev(p["q0"]), dev(p["target"])
cloud = eng.build_cloud(sc, q0, tg)
torch.cuda.synchronize()
eng.profile(True)
traj, metrics = eng.rollout(sc, cloud, q0, tg, T, check_every_step=True, precision=_lib.PREC_BF16)
ev = eng.evaluate(sc, traj, tg)
torch.cuda.synchronize()
st = eng.profile_read()
print({k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in st.items() if v["launches"]})
print("collisions", float(metrics[:, 0].mean()), "eval success", float(ev[:, 9].mean()))
