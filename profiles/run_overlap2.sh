# usage: bash profiles/run_overlap2.sh <ngpus> <tag> "<chunks xcap zcap>" ...   bench only, several pipeline settings
NG=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg; c=$1; x=$2; z=$3
  SDNS_CHUNKS=$c SDNS_XCAP=$x SDNS_ZCAP=$z timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_${TAG}_g${NG}_c${c}x${x}z${z}.json 2> gpurun_out/bench_${TAG}_g${NG}_c${c}x${x}z${z}.err
  python - <<PY
import json
f = "gpurun_out/bench_${TAG}_g${NG}_c${c}x${x}z${z}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("chunks=${c} xcap=${x} zcap=${z}", d["config"]["workload"], "ms/step %.3f  value %.3e  E=%.14f" % (d["ms_per_step"], d["value"], d["config"]["kinetic_energy_after_run"]))
    print("   " + "  ".join("%s %.0fus" % (k, v["ms_per_launch"]*1e3*v["launches_per_step"]/4) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-3000:])
PY
done
