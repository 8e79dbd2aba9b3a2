#!/bin/bash
for rev in 270be60 9729375 e03df0e 45d5180 d565c2f; do
  echo "== $rev"; PMM_LIB=$PWD/variants/libpmm_$rev.so timeout 300 python tests/peer_local_worker.py 4 2000 520 6 2>&1 | tail -2 | cut -c1-200
done
