for rep in 1 2; do for lib in current pf; do
  if [ "$lib" = current ]; then unset B2RT_LIB; else export B2RT_LIB=$PWD/variants/libb2rt_$lib.so; fi
  echo "== lib=$lib rep=$rep"; QUICK=1 timeout 400 python tests/dev_tail2.py 10000000 8 2>&1 | grep -A1 "every rank\|full frame" | grep -v "^--" | head -6
done; done
for lib in current pf; do if [ "$lib" = current ]; then unset B2RT_LIB; else export B2RT_LIB=$PWD/variants/libb2rt_$lib.so; fi; echo "== stream lib=$lib"; python tests/dev_stream.py 50000000 2>&1 | tail -2; done
