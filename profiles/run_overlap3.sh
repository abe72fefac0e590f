# usage: bash profiles/run_overlap3.sh <ngpus> <tag> "<exchange chunks>" ...   multi-GPU parity once, then the bench per setting
NG=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5; fi
for cfg in "$@"; do
  set -- $cfg; x=$1; c=$2; sp=${3:-1}
  SDNS_EXCHANGE=$x SDNS_CHUNKS=$c SDNS_SPLIT=$sp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_${TAG}_g${NG}_${x}${c}s${sp}.json 2> gpurun_out/bench_${TAG}_g${NG}_${x}${c}s${sp}.err
  python - <<PY
import json
f = "gpurun_out/bench_${TAG}_g${NG}_${x}${c}s${sp}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("exchange=${x} chunks=${c} split=${sp}", d["config"]["workload"], "ms/step %.3f  value %.3e  E=%.14f" % (d["ms_per_step"], d["value"], d["config"]["kinetic_energy_after_run"]))
    print("   " + "  ".join("%s %.0fus" % (k, v["ms_per_launch"]*1e3*v["launches_per_step"]/4) for k, v in d["roofline"]["all_kernels"].items()))
    nv = d.get("nvlink", {})
    print("   nvlink GB/s %s  sustained %s  busy ms/step %s" % (nv.get("achieved_GBps_per_direction"), nv.get("sustained_over_step_GBps"), nv.get("copy_stream_busy_ms_per_step")))
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-3000:])
PY
done
