# usage: bash profiles/run_scale.sh <ngpus> <tag> [extra bench args]: multi-GPU parity + bench under torchrun
NG=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG "$@" > gpurun_out/bench_${TAG}_g${NG}.json 2> gpurun_out/bench_${TAG}_g${NG}.err
python - <<PY
import json
f = "gpurun_out/bench_${TAG}_g${NG}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["config"]["workload"], "ms/step %.3f  value %.3e  e2e ms %.2f  step_frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["step_frac"]))
    for k, v in d["roofline"]["all_kernels"].items():
        print("   %-14s %.1f us  %.0f GB/s  share %.3f" % (k, v["ms_per_launch"]*1e3, v["GBps"], v["share"]))
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-3000:])
PY
