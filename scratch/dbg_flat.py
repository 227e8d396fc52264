import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnflow_b200 import DynamicGraph
from gnnflow_b200.synth import synth_stream
n_iter, batch = 200, 20000
src, dst, ts, eid = synth_stream(50000, 0, n_iter * batch, seed=3, t_max=float(n_iter))
g = DynamicGraph(initial_pool_size=8 << 20, maximum_pool_size=4 << 30, mem_resource_type="cuda", minimum_block_size=16,
                 blocks_to_preallocate=1024, insertion_policy="insert")
for it in range(n_iter):
    sl = slice(it * batch, (it + 1) * batch)
    g.add_edges(*[torch.from_numpy(x[sl]).cuda() for x in (src, dst, ts, eid)])
    nb = g.offload_old_blocks(float(ts[sl][0]) - 12.0)
    if it % 20 == 19:
        print(it, nb, g.get_device_memory_usage(), g.get_memory_breakdown(), flush=True)
