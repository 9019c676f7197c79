"""FPS 6272 -> 512 at batch B for each launch variant (MPN_FPS_VARIANT), bit-equality against variant 0:
python scripts/prof_fps.py [B] [variants...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpinets_b200 import scenes
from mpinets_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
variants = sys.argv[2:] or ["0", "1", "2", "3"]
eng = Engine()
for cfg in (2, 4):
    p = scenes.config_problems(cfg, B)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sc = {k: dev(p[k]) for k in scenes.SCENE_KEYS}
    cloud = eng.build_cloud(sc, dev(p["q0"]), dev(p["target"]))
    ref = None
    for v in variants:
        os.environ["MPN_FPS_VARIANT"] = v
        idx = eng.fps(cloud, 512)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx = eng.fps(cloud, 512); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = idx.clone()
        print(f"config {cfg} B={B} variant {v}: {np.median(ts):.3f} ms  equal_to_first={bool(torch.equal(idx, ref))}", flush=True)
