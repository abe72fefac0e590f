#!/bin/sh
# TEST INFRASTRUCTURE (optional, not part of pytest): build the emulated library with ThreadSanitizer and run a few
# RK4 steps, transforms and right-hand sides (tsan_driver.cpp) -- checks that every shared-memory exchange between the
# threads of a CTA is ordered by a barrier.  Cross-rank ordering is checked by tests/test_kernels_emulated.py (results
# under launch jitter): TSan loses its history when tens of thousands of short-lived threads recycle its slots, so
# it does not see write-after-read hazards that are several launches apart.
#   usage: sh tests/host/run_tsan.sh            (about two minutes)
set -e
cd "$(dirname "$0")"
mkdir -p build/tsan
FLAGS='-std=c++20 -O1 -g -fsanitize=thread -fPIC -pthread -w -DSDNS_HOST_SHIM -I . -x c++'
SIZES='-DSDNS_SIZES(X)=X(8) X(12) X(16) X(24) X(256)'
for fam in $(seq 0 14); do for pr in 32 64; do
  g++ $FLAGS "$SIZES" '-DSDNS_SIZES_5(X)=' -DSDNS_FAMILY=$fam -DSDNS_PREC=$pr -c ../../spectraldns_b200/csrc/inst.cu -o build/tsan/inst_${fam}_$pr.o &
done; done
g++ $FLAGS "$SIZES" '-DSDNS_SIZES_5(X)=' -c ../../spectraldns_b200/csrc/sdns_api.cu -o build/tsan/api.o &
wait
g++ -std=c++20 -O1 -g -fsanitize=thread -pthread -o build/tsan_driver tsan_driver.cpp build/tsan/*.o
printf 'race:run_strided\nrace:run_f0x\nrace:run_z\nrace:run_zx_q\nrace:run_zy_q\nrace:run_mhd_f0\nrace:run_nsdiv_f0\n' > build/tsan.supp   # launcher statics shared by rank THREADS (ranks are processes in real runs)
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 suppressions=$PWD/build/tsan.supp"
rc=0
for cfg in "16 16 16 1 1 0" "16 16 16 0 2 0" "16 16 16 1 2 1" "16 16 16 1 1 2" "8 8 256 1 1 0" "256 8 8 1 1 0" "16 16 16 1 1 0 2" "16 16 16 1 2 0 4"; do
  out=$(./build/tsan_driver $cfg 2>&1) || rc=1
  n=$(printf '%s\n' "$out" | grep -c 'WARNING: ThreadSanitizer' || true)
  echo "tsan_driver $cfg: $n report(s)"
  [ "$n" = 0 ] || { printf '%s\n' "$out" | head -40; rc=1; }
done
exit $rc
