"""The product path has no CPU fallback: without a CUDA device every compute entry fails loudly instead of silently
computing on the host (the oracle under oracle/ is test infrastructure and is never imported by the package)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnnflow_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)
                assert "oracle." not in src.replace("oracle/", ""), os.path.join(d, f)


def test_unique_inverse_rejects_host_tensors():
    from gnnflow_b200 import unique_inverse
    with pytest.raises(ValueError):
        unique_inverse(torch.tensor([3, 1, 3]), 10)


@no_gpu
def test_graph_creation_fails_without_a_device():
    from gnnflow_b200 import DynamicGraph
    with pytest.raises((RuntimeError, MemoryError, ValueError)):
        g = DynamicGraph(initial_pool_size=1 << 20, maximum_pool_size=1 << 22, mem_resource_type="cuda",
                         minimum_block_size=4, blocks_to_preallocate=16, insertion_policy="insert")
        g.add_edges(np.array([0, 1]), np.array([1, 2]), np.array([0.0, 1.0], np.float32))
