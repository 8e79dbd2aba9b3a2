#!/bin/bash
# AddressSanitizer + UBSan over the lumped kernels' own source (pmm_lumped.cuh, pmm_nni_lumped.cuh, pmm_nni_ultra.cuh),
# compiled for the host by tests/test_kernels_emulated.py: every record / statistics array of those tests is a
# numpy (malloc) buffer, so an out-of-range index in kernel code lands in an ASan red zone.  This is the memcheck
# that can run without a GPU; compute-sanitizer on the device build is scripts/r2_first_call.sh step 2.
set -e
cd "$(dirname "$0")/../.."
asan=$(g++ -print-file-name=libasan.so)
ubsan=$(g++ -print-file-name=libubsan.so)
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 LD_PRELOAD="$asan $ubsan" \
PMM_EMU_CXXFLAGS="-fsanitize=address,undefined -fno-sanitize-recover=all -fno-omit-frame-pointer -g" \
    python -m pytest tests/test_kernels_emulated.py -q -x -p no:cacheprovider "$@"
