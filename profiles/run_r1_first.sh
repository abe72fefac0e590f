set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"strided_kernel|z_kernel" -s 60 -c 10 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
