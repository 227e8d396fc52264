"""ctypes binding of the C ABI in include/gnnflow_b200.h (gnnflow_b200/lib/libgnnflow_b200.so).

There is no CPU fallback: if the CUDA library is missing, importing the package raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GNNFLOW_B200_LIB: load another build of the same library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get("GNNFLOW_B200_LIB") or os.path.join(_HERE, "lib", "libgnnflow_b200.so")

GF_OK, GF_EINVAL, GF_EORDER, GF_ENOMEM, GF_ECUDA, GF_ECAPACITY, GF_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
GF_PTR_HOST, GF_PTR_DEVICE = 0, 1
INSERTION = {"insert": 0, "replace": 1}
SAMPLING = {"recent": 0, "uniform": 1}
MEM = {"cuda": 0, "unified": 1, "pinned": 2, "shared": 3}


class GraphConfig(C.Structure):
    _fields_ = [("initial_pool_size", C.c_uint64), ("maximum_pool_size", C.c_uint64),
                ("mem_resource_type", C.c_int32), ("minimum_block_size", C.c_uint64),
                ("blocks_to_preallocate", C.c_uint64), ("insertion_policy", C.c_int32),
                ("device", C.c_int32), ("adaptive_block_size", C.c_int32)]


class SamplingResultC(C.Structure):
    _fields_ = [("all_nodes", C.c_void_p), ("all_timestamps", C.c_void_p), ("delta_timestamps", C.c_void_p),
                ("eids", C.c_void_p), ("row", C.c_void_p), ("col", C.c_void_p),
                ("capacity_dst", C.c_uint64), ("num_dst", C.c_uint64), ("num_edges", C.c_uint64)]


class CacheStateC(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("flag", C.c_void_p), ("map", C.c_void_p), ("index_to_id", C.c_void_p),
                ("count", C.c_void_p), ("capacity", C.c_uint64), ("num_items", C.c_uint64), ("dim", C.c_uint32)]


# name -> (restype, argtypes); every symbol declared in include/gnnflow_b200.h
_vp, _u64, _i64, _u32, _i32, _f32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_int, C.c_float
_P = C.POINTER
SIGNATURES = {
    "gf_last_error": (C.c_char_p, []),
    "gf_abi_version": (_i32, []),
    "gf_graph_create": (_i32, [_P(GraphConfig), _P(_vp)]),
    "gf_graph_destroy": (_i32, [_vp]),
    "gf_graph_clear": (_i32, [_vp, _vp]),
    "gf_graph_add_edges": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _i32, _vp]),
    "gf_graph_add_edges_async": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "gf_graph_flush": (_i32, [_vp]),
    "gf_graph_offload_old_blocks": (_i32, [_vp, _f32, _i32, _P(_u64), _vp]),
    "gf_block_file_read": (_i32, [C.c_char_p, _P(_u64), _P(_u64), _P(_f32), _P(_f32), _vp, _vp, _vp, _u64]),
    "gf_graph_num_vertices": (_i32, [_vp, _P(_u64)]),
    "gf_graph_num_source_vertices": (_i32, [_vp, _P(_u64)]),
    "gf_graph_num_edges": (_i32, [_vp, _P(_u64)]),
    "gf_graph_max_vertex_id": (_i32, [_vp, _P(_i64)]),
    "gf_graph_avg_linked_list_length": (_i32, [_vp, _P(_f32)]),
    "gf_graph_memory_usage": (_i32, [_vp, _P(_f32)]),
    "gf_graph_metadata_memory_usage": (_i32, [_vp, _P(_f32)]),
    "gf_graph_device_bytes": (_i32, [_vp, _P(_u64)]),
    "gf_graph_memory_breakdown": (_i32, [_vp, _vp]),
    "gf_graph_save": (_i32, [_vp, C.c_char_p]),
    "gf_graph_load": (_i32, [C.c_char_p, _i32, _P(_vp)]),
    "gf_graph_out_degree": (_i32, [_vp, _vp, _u64, _vp]),
    "gf_graph_nodes": (_i32, [_vp, _vp, _u64, _P(_u64)]),
    "gf_graph_src_nodes": (_i32, [_vp, _vp, _u64, _P(_u64)]),
    "gf_graph_edges": (_i32, [_vp, _vp, _u64, _P(_u64)]),
    "gf_graph_get_temporal_neighbors": (_i32, [_vp, _i64, _vp, _vp, _vp, _u64, _P(_u64)]),
    "gf_graph_block_shapes": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _u64, _P(_u64)]),
    "gf_sampler_create": (_i32, [_vp, _P(_u32), _u32, _i32, _u32, _f32, _i32, _u64, _P(_vp)]),
    "gf_sampler_destroy": (_i32, [_vp]),
    "gf_sampler_sample_layer": (_i32, [_vp, _vp, _vp, _u64, _u32, _u32, _P(SamplingResultC), _i32, _i32, _vp]),
    "gf_sampler_sample": (_i32, [_vp, _vp, _vp, _u64, _P(SamplingResultC), _i32, _i32, _vp]),
    "gf_sampler_sample_layer_batched": (_i32, [_vp, _vp, _vp, _u64, _vp, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _i32, _vp]),
    "gf_sampler_sample_layer_batched_ids32": (_i32, [_vp, _vp, _vp, _u64, _vp, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _vp,
                                                       _vp, _vp]),
    "gf_sampler_chain_batched": (_i32, [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "gf_sampler_get_launch_index": (_i32, [_vp, _P(_u64)]),
    "gf_sampler_set_launch_index": (_i32, [_vp, _u64]),
    "gf_sampler_set_variant": (_i32, [_vp, _i32]),
    "gf_sampler_set_host_output_mode": (_i32, [_vp, _i32]),
    "gf_sampler_bind_host_outputs": (_i32, [_vp, _vp, _u64]),
    "gf_peer_create": (_i32, [_i32, _u32, _u32, _u64, _u32, _P(_vp)]),
    "gf_peer_export": (_i32, [_vp, _vp]),
    "gf_peer_connect": (_i32, [_vp, _vp]),
    "gf_peer_destroy": (_i32, [_vp]),
    "gf_sampler_sample_layer_partitioned": (_i32, [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _u32, _u32, _P(SamplingResultC),
                                                   _vp]),
    "gf_cache_gather": (_i32, [_vp, _u64, _u64, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "gf_gather_rows": (_i32, [_vp, _u64, _u64, _vp, _u32, _vp, _vp, _vp]),
    "gf_cache_update_lru": (_i32, [_P(CacheStateC), _vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp]),
    "gf_cache_update_fifo": (_i32, [_P(CacheStateC), _vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "gf_cache_update_lfu": (_i32, [_P(CacheStateC), _vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp]),
    "gf_cache_count_distinct": (_i32, [_vp, _u64, _vp, _u64, _vp]),
    "gf_cache_fill_topk": (_i32, [_P(CacheStateC), _vp, _vp, _vp, _u64, _vp]),
    "gf_cache_update_scratch_bytes": (_u64, [_u64, _u64, _u64]),
    "gf_cache_fetch": (_i32, [_P(CacheStateC), _vp, _u64, _vp, _u64, _i32, _vp, _u64, _vp, _i32, _vp, _vp, _vp, _vp, _u64,
                              _vp]),
    "gf_cache_fill_scratch_bytes": (_u64, [_u64]),
    "gf_unique_inverse": (_i32, [_vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, _vp]),
    "gf_unique_scratch_bytes": (_u64, [_u64]),
    "gf_shared_alloc": (_i32, [_i32, _u64, _P(_vp)]),
    "gf_shared_free": (_i32, [_vp]),
    "gf_shared_export": (_i32, [_vp, _vp]),
    "gf_shared_open": (_i32, [_i32, _vp, _P(_vp)]),
    "gf_shared_close": (_i32, [_vp]),
    "gf_gather_rows_partitioned": (_i32, [_vp, _u64, _vp, _vp, _u64, _vp, _u32, _u32, _vp, _vp]),
    "gf_host_register": (_i32, [_vp, _u64, _P(C.c_int)]),
    "gf_host_unregister": (_i32, [_vp]),
    "gf_sampler_set_profiling": (_i32, [_vp, _i32]),
    "gf_sampler_get_profile": (_i32, [_vp, _P(C.c_double), _P(_u64), _i32]),
    "gf_graph_set_profiling": (_i32, [_vp, _i32]),
    "gf_graph_get_profile": (_i32, [_vp, _P(C.c_double), _P(_u64), _i32]),
    "gf_debug_launch_count": (_u64, []),
    "gf_dispatch_edges": (_i32, [_vp, _vp, _vp, _vp, _u64, _vp, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _P(_u64), _vp]),
    "gf_l2_fetch_granularity": (_i32, [_i32, _u64, _P(_u64)]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "gnnflow_b200: CUDA library not built ({}). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gnnflow_b200/csrc`. There is no CPU fallback.".format(LIB_PATH))
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().gf_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a gf_status to the exception the reference's Python layer documents (dynamic_graph.py:99-101)."""
    if rc == GF_OK:
        return
    msg = last_error()
    if rc in (GF_EINVAL, GF_EORDER):
        raise ValueError(msg)
    if rc == GF_ENOMEM:
        raise MemoryError(msg)
    if rc == GF_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError("gnnflow_b200 error {}: {}".format(rc, msg))
