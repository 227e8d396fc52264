#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference extension (libgnnflow) from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored, not gpurun-ignored).  Nothing in the product
# path links or loads it; tests/ and bench.py --impl reference may.
#
# The reference's own CMakeLists.txt is not used (it pins C++14 and FindCUDA/select_compute_arch, both
# broken with CUDA 12.9 / CMake 4); this is the direct nvcc/g++ recipe.
set -euo pipefail
R=${GNNFLOW_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
[ -d "$R/gnnflow/csrc" ] || { echo "reference not present at $R; keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OBJ"
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c "import torch,os;print(os.path.dirname(torch.__file__))")
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC="-I$R/gnnflow/csrc -I$R/third_party/rmm/include -I$R/third_party/spdlog/include \
 -I$R/third_party/pybind11/include -I$R/third_party/abseil-cpp -I$PYINC \
 -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include"
NVF="-std=c++17 -gencode arch=compute_100a,code=sm_100a -rdc=true --use_fast_math -lineinfo -O3 \
 -Xcompiler -fopenmp,-fPIC,-w -w -DTORCH_EXTENSION_NAME=libgnnflow"
pids=()
for f in sampling_kernels utils doubly_linked_list temporal_block_allocator dynamic_graph temporal_sampler; do
  nvcc $NVF $INC -c "$R/gnnflow/csrc/$f.cu" -o "$OBJ/$f.o" & pids+=($!)
done
for f in api kvstore logging; do
  nvcc $NVF $INC -x cu -c "$R/gnnflow/csrc/$f.cc" -o "$OBJ/$f.cc.o" & pids+=($!)
done
A=$R/third_party/abseil-cpp
for f in absl/container/internal/raw_hash_set.cc absl/hash/internal/hash.cc absl/hash/internal/city.cc \
         absl/hash/internal/low_level_hash.cc absl/base/internal/raw_logging.cc absl/numeric/int128.cc; do
  g++ -std=c++17 -O2 -fPIC -w -I"$A" -c "$A/$f" -o "$OBJ/absl_$(basename "$f" .cc).o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libgnnflow$EXT" "$OBJ"/*.o \
  -L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -lc10 -ltorch_python -lgomp -lrt \
  -Xlinker -rpath -Xlinker "$TORCH_DIR/lib"
rm -rf "$OBJ"
echo "built $OUT/libgnnflow$EXT"
