# usage: bash profiles/run_r1_final.sh   (under gpurun, 1 GPU): the whole -m gpu suite, then the ncu launch list of a default bench run
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r1b.log 2>&1
tail -c 300 gpurun_out/launches_r1b.log
