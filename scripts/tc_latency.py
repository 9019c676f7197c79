"""tcgen05 round-trip latencies (cycles) from the single-CTA self-test kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpinets_b200.engine import Engine
e = Engine()
names = ["fence.proxy.async", "__syncthreads", "issue K/16 MMAs + commit", "mbar wait (first)", "fence+LDTM.x32+wait", "bar.sync 128",
         "issue 1 MMA + commit", "mbar wait (1 MMA)"]
for N, K in ((64, 64), (128, 128), (256, 128)):
    a = torch.randn(128, K).to(torch.bfloat16).cuda(); b = torch.randn(N, K).to(torch.bfloat16).cuda()
    for rep in range(2):
        lat, to = e.tc_selftest(a, b, 0x100)
    print(f"N={N} K={K}:", {n: int(v) for n, v in zip(names, lat)}, "timeout", to)
