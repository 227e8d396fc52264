import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref"))
import libgnnflow as ref
g = ref._DynamicGraph(1<<20, 1<<26, ref.MemoryResourceType.CUDA, 4, 64, ref.InsertionPolicy.INSERT, 0, True)
src=np.array([0,0,0,1,2,2],dtype=np.int64); dst=np.array([3,4,5,3,4,5],dtype=np.int64)
ts=np.arange(6,dtype=np.float32); eid=np.arange(6,dtype=np.int64)
g.add_edges(src,dst,ts,eid)
print("list", g.out_degree([0,1,2,3]))
print("np", g.out_degree(np.array([0,1,2,3])))
print("single", [g.out_degree([i]) for i in range(4)])
