#!/bin/bash
# AddressSanitizer + UBSan, then ThreadSanitizer, on the host-only part of the library.
set -e
cd "$(dirname "$0")/../.."
out=${TMPDIR:-/tmp}
for san in "address,undefined" "thread"; do
  g++ -std=c++17 -O1 -g -fsanitize=$san -fno-omit-frame-pointer -pthread \
      scripts/native/ultrametric_harness.cpp laml_b200/csrc/pmm_ultrametric.cpp -o $out/ultrametric_$$.bin
  echo "== -fsanitize=$san"
  PMM_HOST_THREADS=8 $out/ultrametric_$$.bin
  rm -f $out/ultrametric_$$.bin
done
echo "sanitizers clean"
