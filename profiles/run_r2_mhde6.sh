# round 2, 1 GPU, last GPU seconds: MHD parity on the final library (fp64 software-pipelined mhd_f0), and fp32 mhd_f0 on the
# 3*2^k lengths with 6 instead of 12 elements per thread (1 KB of spills per thread at E = 12), with its parity.
O=gpurun_out/r2_mhde6; mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -q -k "golden or mhd" > $O/pytest_mhd.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest_mhd.log
for v in default mhde6; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 60 python profiles/tools/passbench.py --only rk4 --tag $v --configs 256:single:3/2-rule:MHD 512:single:3/2-rule:MHD 256:double:2/3-rule:MHD 2> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden" > $O/pytest_mhde6.log 2>&1; echo "pytest mhde6 rc=$?"; tail -1 $O/pytest_mhde6.log
