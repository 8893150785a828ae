import sys, numpy as np
sys.path.insert(0,'/root/repo')
from path_planner_b200 import EdgeEngine, synth
eng=EdgeEngine(0)
for name in ["c2","c3","c3b","c5"]:
    world=synth.WORLDS[name]()
    edges=synth.make_edges(world, 20000, seed=5)
    edges["ribbon_set"]=world.upload(eng)
    r=eng.true_cost_batch(edges)
    ch=np.ceil(r["n_samples"]/32)
    print(name,"chunks/edge %.1f culled/edge %.1f frac %.3f cps %.2f"%(ch.mean(), r["reserved"].mean(), r["reserved"].sum()/ch.sum(), r["n_checkpoints"].mean()))
