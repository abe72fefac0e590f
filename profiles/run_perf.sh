# usage: bash profiles/run_perf.sh <tag>   (under gpurun) -- parity + the two single-GPU bench configurations
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_256d.json 2> gpurun_out/bench_${TAG}_256d.err
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_256d.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f  value %.3e  e2e ms %.2f  step_frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["step_frac"]))
        for k, v in d["roofline"]["all_kernels"].items():
            print("   %-14s %.1f us  %.0f GB/s  share %.3f" % (k, v["ms_per_launch"]*1e3, v["GBps"], v["share"]))
        print("   clocks", d["clocks"])
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-2000:])
PY
timeout 600 python bench.py --no-cpu-baseline --grid 512 --precision single --dealias 3/2-rule --steps 10 > gpurun_out/bench_${TAG}_512s.json 2> gpurun_out/bench_${TAG}_512s.err
python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_512s.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f  value %.3e  e2e ms %.2f  step_frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["step_frac"]))
        for k, v in d["roofline"]["all_kernels"].items():
            print("   %-14s %.1f us  %.0f GB/s  share %.3f" % (k, v["ms_per_launch"]*1e3, v["GBps"], v["share"]))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-2000:])
PY
