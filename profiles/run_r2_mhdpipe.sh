# round 2, 1 GPU, the round's last GPU seconds: MHD F0 with the next field's loads software-pipelined under the current
# transform (default build) against the serial version (-DSDNS_NO_MHD_F0_PIPE), and MHD parity of the default build.
O=gpurun_out/r2_mhdpipe; mkdir -p $O
for v in default mhdnopipe; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 100 python profiles/tools/passbench.py --only rk4 --tag $v --configs 512:double:2/3-rule:MHD 256:double:2/3-rule:MHD 256:single:3/2-rule:MHD 2> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
unset SDNS_LIBPATH
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -q -k "golden or mhd" > $O/pytest_mhd.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_mhd.log
