import cProfile, pstats, sys, io, time
sys.path.insert(0, "/root/repo")
import numpy as np
from dmphyclus_b200 import chain, synth
from dmphyclus_b200.engine import Engine
from dmphyclus_b200.workloads import CONFIGS, Problem, load_packaged
n, s, k, r = CONFIGS["config3"]
p = Problem(load_packaged(), n, s, k, r, "evolved", "random", seed=2001)
eng = Engine(device=0)
al = eng.getConvertedAlignment(list(synth.STATES), p.chars)
cs = chain.ChainSettings(limProbs=p.pi, withinTransMatAll=p.W, betweenTransMatAll=p.B, shapeForAlpha=1000,
                         scaleForAlpha=0.1, alphaMin=10, poisRateNumClus=2, numSplitMergeMoves=3)
chain.run_chain(eng, al, p.edge, p.clus, 100.0, cs, 3, seed=2001)
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
chain.run_chain(eng, al, p.edge, p.clus, 100.0, cs, 12, seed=2001)
dt = time.perf_counter() - t0
pr.disable()
print("12 iterations", dt)
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue())
