# usage: bash profiles/run_overlap.sh <ngpus> <tag> [chunk counts...]: multi-GPU parity, then the bench at several
# pipeline chunk counts (SDNS_CHUNKS=1 is the serial schedule)
NG=$1; TAG=$2; shift; shift
CH=${@:-1 4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
for c in $CH; do
  SDNS_CHUNKS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_${TAG}_g${NG}_c${c}.json 2> gpurun_out/bench_${TAG}_g${NG}_c${c}.err
  python - <<PY
import json
f = "gpurun_out/bench_${TAG}_g${NG}_c${c}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("chunks=${c}", d["config"]["workload"], "ms/step %.3f  value %.3e  step_frac %.3f" % (d["ms_per_step"], d["value"], d["roofline"]["step_frac"]))
    for k, v in d["roofline"]["all_kernels"].items():
        print("   %-14s %.1f us  %.0f GB/s  share %.3f" % (k, v["ms_per_launch"]*1e3, v["GBps"], v["share"]))
    print("   nvlink", d.get("nvlink", {}).get("achieved_GBps_per_direction"))
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-3000:])
PY
done
