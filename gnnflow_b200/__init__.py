"""gnnflow_b200: B200-native dynamic-graph store, temporal sampler and feature-cache gather behind GNNFlow's
Python API (reference gnnflow/__init__.py:1-2)."""
from .dynamic_graph import DynamicGraph  # noqa: F401
from .temporal_sampler import Block, SamplingResult, TemporalSampler  # noqa: F401
from .mfg_ops import gather_rows, prepare_memory_input, unique_inverse  # noqa: F401
