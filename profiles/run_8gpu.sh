# usage: bash profiles/run_8gpu.sh <ngpus>  (under gpurun --gpus N): slab parity at N ranks, the weak-scaling bench line
# (256^3 per GPU) and -- on 8 GPUs -- BASELINE.json's target configuration 1024^3 double
NG=${1:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29508 tests/mp/slab_worker.py 2>&1 | grep -E "OK|FAIL|RESULT|rror" | cut -c1-200 | tail -14
summ() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "| ms/step %.3f  value %.3e  step_frac %.3f  E=%.12f" % (d["ms_per_step"], d["value"], d["roofline"]["step_frac"], d["config"]["kinetic_energy_after_run"]))
    print("   " + "  ".join("%s %.0fus/rhs" % (k, v["ms_per_launch"]*1e3*v["launches_per_step"]/4) for k, v in d["roofline"]["all_kernels"].items()))
    print("   nvlink", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.get("nvlink", {}).items() if k != "note"})
except Exception as e:
    print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-2500:])
PY
}
run 29533 bench.py --gpus $NG --no-cpu-baseline --steps 10 > gpurun_out/bench_r1_g${NG}_weak.json 2> gpurun_out/bench_r1_g${NG}_weak.err; summ gpurun_out/bench_r1_g${NG}_weak.json
for v in $VARIANTS; do
  env $(echo $v | tr ',' ' ') timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $NG --no-cpu-baseline --steps 10 > gpurun_out/bench_r1_g${NG}_weak_$v.json 2> gpurun_out/bench_r1_g${NG}_weak_$v.err; echo "variant $v"; summ gpurun_out/bench_r1_g${NG}_weak_$v.json
done
if [ "$NG" = 8 ]; then
  run 29534 bench.py --gpus 8 --grid 1024 --scaling strong --steps 3 --no-cpu-baseline > gpurun_out/bench_r1_g8_1024d.json 2> gpurun_out/bench_r1_g8_1024d.err; summ gpurun_out/bench_r1_g8_1024d.json
fi
