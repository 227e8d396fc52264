#!/usr/bin/env bash
# scratch/build_variant.sh <name> <extra nvcc flags...>: builds scratch/variants/lib_<name>.so (kernel experiments)
set -e
NAME=$1; shift
mkdir -p /root/repo/scratch/variants
make -s -C /root/repo/gnnflow_b200/csrc -j4 OUT=/root/repo/scratch/variants/lib_$NAME.so OBJDIR=/tmp/gfobj_$NAME EXTRA="$*" 2>&1 | grep -v deprecated || true
ls -la /root/repo/scratch/variants/lib_$NAME.so
