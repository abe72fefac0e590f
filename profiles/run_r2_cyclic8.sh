# round 2, 8 GPUs (every second costs eight): the slab worker with the cyclic axis-1 ownership, the 1024^3 north-star
# configuration with cyclic (bench.py's choice on 3 or more GPUs) against the reference's contiguous blocks, then the two
# BASELINE configurations that had never run: TG-MHD 512^3 fp64 (configs[3]) and NS 2048^3 fp32 (configs[4]).
NG=${1:-8}
O=gpurun_out/r2_cyclic${NG}; mkdir -p $O
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
show() { python - <<PY
import json
try:
    d = json.loads(open("$O/$1.out").read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("$1", d["config"]["workload"], d["config"].get("k1_layout", "")[:6], "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "nvlink", nv.get("sustained_GBps_per_direction"), "parity", (d.get("parity") or {}).get("worst_err_over_tol_all_ranks"), "E", d["config"].get("kinetic_energy_after_run"))
    print("   " + "  ".join("%s %.0fus x%.0f %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["launches_per_step"], v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$1 FAILED", e, open("$O/$1.err").read()[-1200:])
PY
}
SLAB_K1_LAYOUT=cyclic SLAB_CASES=0,1,4,7,9,10,14 run slab_cyclic tests/mp/slab_worker.py; tail -9 $O/slab_cyclic.out
if grep -q "SLAB_WORKER_RESULT fails=0" $O/slab_cyclic.out; then
  run bench_1024_cyclic bench.py --gpus $NG --steps 6 --no-cpu-baseline --timeline $O/timeline_cyclic; show bench_1024_cyclic
  LAY=""
else
  echo "cyclic slab worker FAILED: the remaining runs use the contiguous blocks"; tail -c 1500 $O/slab_cyclic.err
  LAY="--k1-layout blocks"
fi
run bench_1024_blocks bench.py --gpus $NG --steps 6 --no-cpu-baseline --no-parity --k1-layout blocks; show bench_1024_blocks
run bench_mhd512 bench.py --gpus $NG --steps 10 --grid 256 --solver MHD --no-cpu-baseline $LAY; show bench_mhd512
run bench_2048s bench.py --gpus $NG --steps 4 --grid 1024 --precision single --no-cpu-baseline --no-parity $LAY; show bench_2048s
python profiles/tools/timeline_report.py $O/timeline_cyclic 2>&1 | head -12
