import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mpinets_b200 import scenes, _lib
from mpinets_b200.engine import Engine
from oracle import oracle as O
eng = Engine(); eng.load_state_dict(O.reference_state_dict(0))
B = 64
p = scenes.config_problems(4, B)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sc = {k: dev(p[k]) for k in scenes.SCENE_KEYS}
q0, tg = dev(p["q0"]), dev(p["target"])
cloud = eng.build_cloud(sc, q0, tg)
qn = eng.normalize(q0)
sup = torch.clamp(qn + 0.05 * torch.randn_like(qn), -1, 1)
for prec in (_lib.PREC_FP32, _lib.PREC_BF16):
    l, y, g = eng.train_step_grads(sc, cloud, qn, sup, precision=prec)
    a1, a2, a3 = eng.train_pooled_rows(B)
    for name, a in (("SA1", a1), ("SA2", a2)):
        a = a.cpu().numpy()            # [B, groups, channels]
        d = np.array([[len(np.unique(a[b, g_])) for g_ in range(a.shape[1])] for b in range(B)])
        print(prec, name, "distinct winning rows per group: mean %.1f  median %d  p90 %d  max %d  (slots %d)" % (d.mean(), np.median(d), np.percentile(d, 90), d.max(), 64 if name == "SA1" else 128))
