# Developer A/B (GPU): stage times of a rank's frame share under several builds of libb2rt.so (variants/, see dev_variant.py)
for lib in "$@"; do
  if [ "$lib" = current ]; then unset B2RT_LIB; else export B2RT_LIB=$PWD/variants/libb2rt_$lib.so; fi
  echo "== lib=$lib"; COOPS=8:0 timeout 400 python tests/dev_tail3.py 10000000 8 0 2>&1 | grep "stages"
  QUICK=1 timeout 400 python tests/dev_tail2.py 10000000 8 2>&1 | grep -A1 "every rank\|full frame" | grep -v "^--\|==" | head -3
done
