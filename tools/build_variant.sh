#!/bin/bash
# Builds a variant of the library with extra nvcc flags into elections.dtree_b200/lib/libdtree_b200_<name>.so
# (selected at run time with DTREE_B200_LIB). usage: tools/build_variant.sh <name> <flags...>
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
T=/tmp/dtb_variant_$NAME
rm -rf $T && mkdir -p $T/pkg/x $T/include
cp -r $ROOT/include/. $T/include/
mkdir -p $T/a/b && cp $ROOT/elections.dtree_b200/csrc/*.cu $ROOT/elections.dtree_b200/csrc/*.cuh $ROOT/elections.dtree_b200/csrc/*.h \
  $ROOT/elections.dtree_b200/csrc/*.cpp $ROOT/elections.dtree_b200/csrc/Makefile $T/a/b/
mkdir -p $T/include && rm -rf $T/pkg
# the Makefile looks for ../../include and writes ../lib
cp -r $ROOT/include $T/ 2>/dev/null || true
make -C $T/a/b -j 8 EXTRA="$*" >/dev/null
cp $T/a/lib/libdtree_b200.so $ROOT/elections.dtree_b200/lib/libdtree_b200_$NAME.so
echo built elections.dtree_b200/lib/libdtree_b200_$NAME.so
