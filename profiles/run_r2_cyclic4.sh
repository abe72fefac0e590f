# round 2, 4 GPUs: the cyclic axis-1 ownership with closed-form rows in the axis-1 passes (slab worker subset incl. 128^3),
# NS weak scaling at 512^3 per GPU (1024 x 1024 x 512) cyclic against blocks, and the per-GPU load and strides of BASELINE
# configs[4] (NS 2048^3 fp32 on 8 GPUs) on half the box: 2048 x 2048 x 1024 fp32 (the 32-bit row-offset guard used to refuse it).
NG=${1:-4}
O=gpurun_out/r2_cyclic${NG}; mkdir -p $O
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
show() { python - <<PY
import json
try:
    d = json.loads(open("$O/$1.out").read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("$1", d["config"]["workload"], (d["config"].get("k1_layout") or "")[:6], "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "nvlink", nv.get("sustained_GBps_per_direction"), "parity", (d.get("parity") or {}).get("worst_err_over_tol_all_ranks"), "E", d["config"].get("kinetic_energy_after_run"))
    print("   " + "  ".join("%s %.0fus x%.0f %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["launches_per_step"], v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$1 FAILED", e, open("$O/$1.err").read()[-1200:])
PY
}
SLAB_K1_LAYOUT=cyclic SLAB_CASES=0,1,4,6,8,12 run slab_cyclic tests/mp/slab_worker.py; tail -8 $O/slab_cyclic.out
run bench_cyclic bench.py --gpus $NG --steps 6 --no-cpu-baseline; show bench_cyclic
run bench_blocks bench.py --gpus $NG --steps 6 --no-cpu-baseline --no-parity --k1-layout blocks; show bench_blocks
run bench_2048x2048x1024s bench.py --gpus $NG --steps 4 --grid 1024 --precision single --no-cpu-baseline --no-parity; show bench_2048x2048x1024s
