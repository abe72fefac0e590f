# round 2, 1 GPU: the whole -m gpu suite on the default build (field-parallel B0, stand-alone operators, 1024/2048 on the
# strided axes, 60/90), then per-pass times of the default build and three experiment builds:
#   v1 = strided_kernel B0 (as in round 1) + F0 with the L2 prefetch of its epilogue state
#   v2 = F0 prefetch + unrolled epilogue      v3 = F0 unrolled epilogue only
O=gpurun_out/r2_var; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for v in default v1 v2 v3; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 600 python profiles/tools/passbench.py --only rk4 --tag $v --configs 256:double:2/3-rule 512:double:2/3-rule 512:single:2/3-rule 512:single:3/2-rule 256:double:2/3-rule:VV 256:double:3/2-rule 2> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
