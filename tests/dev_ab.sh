#!/bin/bash
# Developer A/B (GPU): the ray-stream rate of several builds of libb2rt.so (variants/libb2rt_<name>.so, see dev_variant.py) on one box.
# usage: dev_ab.sh <rays> <name> [<name> ...]   ("current" = the in-tree library)
n=$1; shift
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = current ]; then unset B2RT_LIB; else export B2RT_LIB=$PWD/variants/libb2rt_$lib.so; fi
  echo "== lib=$lib rep=$rep n=$n"; python tests/dev_stream.py $n 2>&1 | tail -4
done; done
