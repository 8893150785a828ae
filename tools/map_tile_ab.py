"""N2 A/B (GPU): static-map look-ups of the thread walker from (a) global memory through __ldg (L2-resident bitmap),
(b) the same with a persisting L2 access-policy window, (c) a shared-memory tile of the batch's bounding box staged with
TMA bulk copies.  Workload: the C4 frontier-batch shape -- 2^19 edges whose sources lie within 150 m of the start state
(what the top vertices of the open list x samples produce), device-resident, CUDA events.  Prints one JSON line per mode."""
import ctypes as C
import json
import os
import subprocess
import sys

sys.path.insert(0, "/root/repo")
MODES = {"ldg": {}, "l2_persist": {"PPE_MAP_L2_PERSIST": "1"}, "smem_tile": {"PPE_MAP_TILE": "1"}}


def run(mode, world_name, n):
    import numpy as np
    import torch
    from path_planner_b200 import EdgeEngine, abi, synth
    world = synth.WORLDS[world_name]()
    eng = EdgeEngine(0)
    sid = world.upload(eng)
    cx, cy = world.start[0], world.start[1]
    edges = synth.make_edges(world, n, seed=7)
    edges["src"][:, 0] = cx + (edges["src"][:, 0] % 300.0) - 150.0
    edges["src"][:, 1] = cy + (edges["src"][:, 1] % 300.0) - 150.0
    edges["dst"][:, 0] = edges["src"][:, 0] + (edges["dst"][:, 0] % 120.0) - 60.0
    edges["dst"][:, 1] = edges["src"][:, 1] + (edges["dst"][:, 1] % 120.0) - 60.0
    edges["ribbon_set"] = sid
    eng._lib.ppe_set_map_window.argtypes = [C.c_void_p] + [C.c_double] * 4
    eng._lib.ppe_set_map_window(eng._ctx, cx - 230.0, cy - 230.0, cx + 230.0, cy + 230.0)
    dev = torch.device("cuda", 0)
    d_e = torch.from_numpy(edges.view(np.uint8).reshape(n, abi.EDGE_DTYPE.itemsize)).to(dev)
    d_r = torch.empty((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    sh = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.true_cost_batch_device(n, d_e.data_ptr(), d_r.data_ptr(), sh)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    steps = 10
    for _ in range(steps):
        eng.true_cost_batch_device(n, d_e.data_ptr(), d_r.data_ptr(), sh)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    r32 = d_r.view(torch.int32).reshape(n, abi.RESULT_DTYPE.itemsize // 4)
    chk = int(r32[:, 45].to(torch.int64).sum().item()) * 1000003 + int(r32[:, 47].to(torch.int64).sum().item())
    print(json.dumps({"mode": mode, "world": world_name, "edges": n, "ms_per_step": ms, "edges_per_s": n / (ms * 1e-3),
                      "infeasible_and_samples_checksum": chk}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3:
        run(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 19)
    else:
        for world in ("c4", "c5"):
            for mode, env in MODES.items():
                e = dict(os.environ)
                e.update(env)
                subprocess.run([sys.executable, __file__, mode, world], env=e, check=True)
