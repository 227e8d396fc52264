"""The C-ABI boundary without a GPU: the header, the ctypes table and the built library agree.

  * every function declared in include/gnnflow_b200.h has an entry in gnnflow_b200._lib.SIGNATURES (and vice versa),
    with the same number of parameters;
  * libgnnflow_b200.so loads on a CPU-only box and exports every declared symbol;
  * the ctypes mirrors of the ABI structs have the C layout (checked by compiling a tiny C program against the header);
  * the header is plain C (compiles with gcc -std=c99 -pedantic), i.e. nothing but pointers and sizes crosses it;
  * no compute entry is called here -- that is what the -m gpu tests do."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gnnflow_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    src = re.sub(r"typedef\s+enum\s*\w*\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(gf_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        nargs = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[name] = nargs
    return out


def test_header_and_ctypes_table_agree():
    from gnnflow_b200 import _lib
    decl = _declared_functions()
    assert len(decl) >= 40, sorted(decl)
    assert set(decl) == set(_lib.SIGNATURES), (sorted(set(decl) ^ set(_lib.SIGNATURES)))
    for name, nargs in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_library_loads_and_exports_every_symbol():
    from gnnflow_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = C.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(L, name), "libgnnflow_b200.so does not export " + name
    L.gf_abi_version.restype = C.c_int
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gnnflow_b200.h")).read()
    assert L.gf_abi_version() == int(re.search(r"#define GF_ABI_VERSION (\d+)", hdr).group(1)) == 7
    # no torch / pybind / libstdc++-typed symbol is exported: the dynamic symbol table holds gf_* only
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert exported and all(s.startswith("gf_") for s in exported), [s for s in exported if not s.startswith("gf_")][:5]


def test_missing_library_fails_loudly(monkeypatch):
    from gnnflow_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "gnnflow_b200", "lib", "does_not_exist.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_struct_layouts_match_c(tmp_path):
    from gnnflow_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "gnnflow_b200.h"
#define F(s, f) printf(#s "." #f " %zu\n", offsetof(s, f));
int main(void) {
  printf("gf_graph_config %zu\n", sizeof(gf_graph_config));
  F(gf_graph_config, initial_pool_size) F(gf_graph_config, maximum_pool_size) F(gf_graph_config, mem_resource_type)
  F(gf_graph_config, minimum_block_size) F(gf_graph_config, blocks_to_preallocate) F(gf_graph_config, insertion_policy)
  F(gf_graph_config, device) F(gf_graph_config, adaptive_block_size)
  printf("gf_sampling_result %zu\n", sizeof(gf_sampling_result));
  F(gf_sampling_result, all_nodes) F(gf_sampling_result, all_timestamps) F(gf_sampling_result, delta_timestamps)
  F(gf_sampling_result, eids) F(gf_sampling_result, row) F(gf_sampling_result, col) F(gf_sampling_result, capacity_dst)
  F(gf_sampling_result, num_dst) F(gf_sampling_result, num_edges)
  printf("gf_cache_state %zu\n", sizeof(gf_cache_state));
  F(gf_cache_state, buffer) F(gf_cache_state, flag) F(gf_cache_state, map) F(gf_cache_state, index_to_id)
  F(gf_cache_state, count) F(gf_cache_state, capacity) F(gf_cache_state, num_items) F(gf_cache_state, dim)
  return 0;
}''')
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(prog), "-o", str(exe)])
    got = dict(ln.rsplit(" ", 1) for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in (("gf_graph_config", _lib.GraphConfig), ("gf_sampling_result", _lib.SamplingResultC),
                      ("gf_cache_state", _lib.CacheStateC)):
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
