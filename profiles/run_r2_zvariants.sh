# round 2: the bulk-async staged z pass (zb_kernel) against zx_kernel, per variant library (1 GPU)
O=gpurun_out/r2_zvar; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_parity.log
for v in default zb1 zb1l4 zb2l1 nozb; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  for cfg in "256 double 2/3-rule" "512 single 3/2-rule" "512 single 2/3-rule" "256 double 3/2-rule"; do
    set -- $cfg
    timeout 300 python profiles/tools/passbench.py --grid $1 --precision $2 --dealias $3 --tag $v 2> $O/pb_${v}_$1_$2.err | grep -E "^rk4" 
  done
done
