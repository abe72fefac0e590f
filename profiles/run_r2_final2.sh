# round 2, 1 GPU, last call: the -m gpu suite on the final library, the default bench line, TG-MHD 512^3 fp64 with the stable dt,
# and the F0 experiment: serial-field F0 with 512-thread CTAs (128-byte rows at N0 = 512, 64-byte rows at N0 = 1024) against
# the field-parallel F0 (64 / 32-byte rows), per-pass times at 256^3, 512^3 and on 1024-long axis-0 lines, plus parity of the variant.
O=gpurun_out/r2_final2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 > $O/bench_1gpu_512d.json 2> $O/bench_1gpu_512d.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 6 --solver MHD --grid 512 --no-cpu-baseline > $O/bench_1gpu_mhd512d.json 2> $O/bench_1gpu_mhd512d.err; echo "bench mhd512 rc=$?"; tail -c 300 $O/bench_1gpu_mhd512d.err
for v in default f0wide; do
  if [ "$v" = default ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 300 python profiles/tools/passbench.py --only rk4 --tag $v --configs 512:double:2/3-rule 256:double:2/3-rule 256:double:3/2-rule 2> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
  timeout 300 python profiles/tools/passbench.py --only rk4 --tag $v --grid 1024 128 512 2>> $O/pb_$v.err | grep -E "^rk4" | tee -a $O/passbench.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or long_axis or full_size or pressure" > $O/pytest_f0wide.log 2>&1; echo "pytest f0wide rc=$?"; tail -2 $O/pytest_f0wide.log
unset SDNS_LIBPATH
python - <<'PY'
import json
for f in ('bench_1gpu_512d', 'bench_1gpu_mhd512d'):
    try:
        d = json.loads(open('gpurun_out/r2_final2/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f' % d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', (d.get('parity') or {}).get('worst_err_over_tol_all_ranks'), 'E', d['config'].get('kinetic_energy_after_run'), 'dt', d['config'].get('dt'))
        print('   ' + '  '.join('%s %.0fus %.0fGB/s' % (k, v['ms_per_launch']*1e3, v['GBps']) for k, v in d['roofline']['all_kernels'].items()))
    except Exception as e:
        print(f, 'FAILED', e)
PY
