"""A small run through every new device path of round 2, for `compute-sanitizer --tool memcheck`:
slabbed upload, rounds of paths, chain certificate (and its literal fallback), batched gather (merge), two-shard handle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Problem, scripted_moves
from dmphyclus_b200.engine import Engine
from dmphyclus_b200.workloads import load_packaged

pk = load_packaged()
os.environ["DMPC_UPLOAD_SLABS"] = "2"
for n, S, k, R, kind, seed in ((10000, 300, 100, 3, "evolved", 1), (700, 330, 6, 3, "iid", 2), (400, 300, 5, 5, "iid", 3)):
    p = Problem(pk, n, S, k, R, kind, seed=seed)
    for eng in (Engine(), Engine(chain_cert=False), Engine(devices=[0, 0])):
        h, ll = p.create(eng)
        out = scripted_moves(eng, h, p, False)
        st = eng.stats(h)
        print(n, S, R, kind, type(eng).__name__, ll, out[-1], "cert", st["last_cert"], flush=True)
        if eng.opts.n_devices == 0:
            big = [m for m in p.mrcas if m > n][:6]
            got, nb = eng.evalProposals(h, p.W[6], p.B[4], p.pi, [("split", m) for m in big])
            print("  batch", nb, got[:2], flush=True)
        eng.finalDeallocate(h)
print("sanitize case done")
