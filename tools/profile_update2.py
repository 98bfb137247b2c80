import sys, os, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch, bench
dev = torch.device("cuda:0")
for k in range(3):
    r = bench.update_loop_profile(dev)
    print({a: (round(b, 2) if isinstance(b, float) else b) for a, b in r.items() if a != "what"}, flush=True)
