for rep in 1 2; do
for lib in "" r1 notail; do
  if [ -z "$lib" ]; then unset B2RT_LIB; else export B2RT_LIB=$PWD/variants/libb2rt_$lib.so; fi
  echo "== lib=${lib:-current} rep=$rep n=$1"; python tests/dev_stream.py $1 2>&1 | tail -4
done; done
