#!/bin/bash
for v in v1 v3; do for n in 500 2000 5000 9996; do
  echo -n "$v n=$n: "; PMM_LIB=$PWD/variants/libpmm_$v.so timeout 300 python scripts/prof_nni_ultra.py $n 2>&1 | tail -1
done; done
