"""compute-sanitizer target: one tiny training step + Adam (run as `compute-sanitizer --tool memcheck python scripts/sanitize_train.py`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpinets_b200 import scenes  # noqa: E402
from mpinets_b200.engine import Engine  # noqa: E402
from oracle import oracle as O  # noqa: E402  (weights init only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
PREC = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # 0 fp32, 1 bf16 (tensor-core backward)
os.environ["MPN_TRAIN_CHUNK"] = "1"
eng = Engine(device=0)
eng.load_state_dict(O.reference_state_dict(0))
p = scenes.config_problems(4, B)
d = {k: torch.from_numpy(np.ascontiguousarray(p[k])).cuda() for k in scenes.SCENE_KEYS + ("q0", "target")}
sc = {k: d[k] for k in scenes.SCENE_KEYS}
cloud = eng.build_cloud(sc, d["q0"], d["target"])
qn = eng.normalize(d["q0"])
sup = torch.clamp(qn + 0.05, -1, 1)
losses, y, g = eng.train_step_grads(sc, cloud, qn, sup, precision=PREC)
n = eng.adam_step(g, 1)
torch.cuda.synchronize()
print("sanitize_train: losses", losses.cpu().tolist(), "grad norm", float(n), "finite", bool(torch.isfinite(g).all()))
