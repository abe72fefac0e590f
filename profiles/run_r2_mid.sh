# round 2, 1 GPU: the whole -m gpu suite (new: stand-alone operators, device diagnostics, lazy state mirror, 2-D solvers,
# 1024/2048 on the strided axes), the forced-run comparison eager vs lazy, the default bench line, the ncu launch list
# of the default bench command and one ncu --set full capture of a right-hand side at 512^3 fp64.
O=gpurun_out/r2_mid; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
for lazy in 0 1; do
  timeout 300 python profiles/tools/forced_run_bench.py --grid 256 --precision double --steps 10 --lazy $lazy 2> $O/forced_$lazy.err | grep FORCED | tee -a $O/forced.txt
  timeout 300 python profiles/tools/forced_run_bench.py --grid 256 --precision single --dealias 3/2-rule --steps 10 --lazy $lazy 2>> $O/forced_$lazy.err | grep FORCED | tee -a $O/forced.txt
done
timeout 600 python bench.py --steps 10 > $O/bench_1gpu_512d.json 2> $O/bench_1gpu_512d.err; echo "bench rc=$?"; tail -c 600 $O/bench_1gpu_512d.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/launches_r2.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"strided_kernel|zy_kernel|zx_kernel|f0x_kernel" -s 65 -c 5 -f -o $O/prof_r2_512d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_r2_512d.ncu-rep --page raw --csv > $O/ncu_full_512d_raw.csv 2>/dev/null
ls -la $O | tail -12
