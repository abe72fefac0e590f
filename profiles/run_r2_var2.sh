# round 2, 1 GPU: the GPU tests that failed in the previous call (fixed since), then per-pass times of
#   default (composite radix-12/24 first stage on the 3*2^k lengths)   nox3 = without it
#   plain84 = fp64 plain passes limited to 84 registers (3 CTAs per SM at N = 512)   nof0x = serial-field F0 instead of field-parallel
O=gpurun_out/r2_var2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -q -k "standalone or scripts_run_unchanged or dealiased or long_axis or golden" > $O/pytest_fixed.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_fixed.log
for v in default nox3; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 600 python profiles/tools/passbench.py --only rk4 --tag $v --configs 512:single:3/2-rule 256:double:3/2-rule 256:single:3/2-rule 256:double:3/2-rule:VV 2> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
for v in default plain84 nof0x; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 600 python profiles/tools/passbench.py --only rk4 --tag $v --configs 512:double:2/3-rule 256:double:2/3-rule 2> $O/pb2_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
