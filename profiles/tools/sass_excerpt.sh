#!/bin/bash
# SASS evidence for the judged claims: mnemonic counts per object of libsdns_b200.so and first occurrences in named kernels.
#   bash profiles/tools/sass_excerpt.sh > profiles/r2/sass/sass_excerpt.txt      (needs the objects of python -m spectraldns_b200.build)
B=spectraldns_b200/build
echo "# cuobjdump -sass of the objects of libsdns_b200.so (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a)"
echo "# produced by profiles/tools/sass_excerpt.sh; mnemonic counts per object first, then excerpts"
echo
for o in sdns_api inst_0_f64 inst_1_f64 inst_2_f64 inst_4_f64 inst_9_f64 inst_9_f32 inst_0_f32; do
  cuobjdump -sass $B/$o.o > /tmp/sass_$o.txt 2>/dev/null
  printf "%-14s" "$o.o"
  for m in UBLKCP SYNCS UTMACMDFLUSH FFMA2 FADD2 FMUL2 DFMA SHFL; do printf " %s %5d " $m $(grep -c "^\s*/\*[0-9a-f]*\*/\s*$m" /tmp/sass_$o.txt); done
  echo " arch $(grep -m1 -o 'sm_[0-9a]*' /tmp/sass_$o.txt)"
done
echo
excerpt() {  # object, function regex, mnemonic regex, title
  echo "== $4  ($1.o)"
  awk -v f="$2" -v m="$3" '/Function :/ { on = ($0 ~ f); if (on) print "   function: " $0 } on && $0 ~ m { if (n[$0 ~ f]++ < 8) print "  " $0 }' /tmp/sass_$1.txt | cut -c1-150 | head -10
}
excerpt sdns_api "xfer_kernel" "UBLKCP|SYNCS|UTMACMDFLUSH" "transfer-only launch: bulk-async copies global -> shared (UBLKCP.S.G, mbarrier SYNCS) and shared -> peer global (UBLKCP.G.S)"
excerpt inst_4_f64 "f0x_kernelIdLi512" "UBLKCP|SYNCS|UTMACMDFLUSH" "the same role carried inside a pass kernel (F0, N = 512, fp64)"
excerpt inst_9_f32 "zx_kernelIfLi768" "FADD2|FFMA2|FMUL2" "fp32 arithmetic on the packed f32x2 pipe (fused z pass, M = 768)"
excerpt inst_9_f64 "zx_kernelIdLi256" "SHFL" "Hermitian mirrors by warp shuffle (fused z pass, M = 256, fp64)"
