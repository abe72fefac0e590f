# usage: bash profiles/run_variants.sh <tag> "<bench args>" variant...   bench one configuration with experiment builds of the library
TAG=$1; ARGS=$2; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset SDNS_LIBPATH; else export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline $ARGS > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  python - <<PY
import json
f = "gpurun_out/bench_${TAG}_$v.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("$v", d["config"]["workload"], "ms/step %.3f step_frac %.3f" % (d["ms_per_step"], d["roofline"]["step_frac"]))
    print("   " + "  ".join("%s %.0fus %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-2000:])
PY
done
