# round 2, 1 GPU, final code: the whole -m gpu suite, the default bench line (512^3 fp64), BASELINE configs[1] (256^3 fp64) and
# configs[2]'s per-GPU load (512^3 fp32 3/2-rule), the ncu launch list of the default bench command and one ncu --set full
# capture of a right-hand side at 512^3 fp64.
O=gpurun_out/r2_final1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 > $O/bench_1gpu_512d.json 2> $O/bench_1gpu_512d.err; echo "bench rc=$?"; tail -c 400 $O/bench_1gpu_512d.json
timeout 300 python bench.py --steps 20 --grid 256 --no-cpu-baseline > $O/bench_1gpu_256d.json 2> $O/bench_1gpu_256d.err; echo "bench256 rc=$?"
timeout 300 python bench.py --steps 10 --grid 512 --precision single --dealias 3/2-rule --no-cpu-baseline > $O/bench_1gpu_512s32.json 2> $O/bench_1gpu_512s32.err; echo "bench512s32 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/launches_r2_final.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"strided_kernel|zy_kernel|zx_kernel|f0x_kernel" -s 65 -c 5 -f -o $O/prof_r2_final_512d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_r2_final_512d.ncu-rep --page raw --csv > $O/ncu_full_512d_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ('bench_1gpu_512d', 'bench_1gpu_256d', 'bench_1gpu_512s32'):
    try:
        d = json.loads(open('gpurun_out/r2_final1/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f' % d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', (d.get('parity') or {}).get('worst_err_over_tol_all_ranks'))
        print('   ' + '  '.join('%s %.0fus %.0fGB/s' % (k, v['ms_per_launch']*1e3, v['GBps']) for k, v in d['roofline']['all_kernels'].items()))
    except Exception as e:
        print(f, 'FAILED', e)
PY
