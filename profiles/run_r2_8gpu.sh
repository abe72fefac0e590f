# round 2, 8 GPUs: parity of the slab worker, then the 1024^3 north-star configuration (bench.py default grid, weak
# scaling) with the transfer-role exchange against the round-1 copy-engine exchange, plus a few knobs.
NG=${1:-8}
O=gpurun_out/r2_${NG}gpu; mkdir -p $O
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
show() { python - <<PY
import json
try:
    d = json.loads(open("$O/$1.out").read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("$1", d["config"]["workload"], "ms/step %.3f" % d["ms_per_step"], "nvlink", nv.get("sustained_GBps_per_direction", nv.get("sustained_over_step_GBps")), "flushes", nv.get("transfer_only_launches_per_step"), "parity", (d.get("parity") or {}).get("worst_err_over_tol_all_ranks"), "E", d["config"].get("kinetic_energy_after_run"))
    print("   " + "  ".join("%s %.0fus x%.0f %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["launches_per_step"], v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$1 FAILED", e, open("$O/$1.err").read()[-1500:])
PY
}
run slab_tma tests/mp/slab_worker.py; tail -11 $O/slab_tma.out
run bench_tma_1024 bench.py --gpus $NG --steps 4 --timeline $O/timeline_tma; show bench_tma_1024
SDNS_EXCHANGE=ce run bench_ce_1024 bench.py --gpus $NG --steps 4 --no-parity --timeline $O/timeline_ce; show bench_ce_1024
SDNS_XRATIO=0.6 run bench_tma_1024_r6 bench.py --gpus $NG --steps 4 --no-parity; show bench_tma_1024_r6
SDNS_XRATIO=0.16 run bench_tma_1024_r16 bench.py --gpus $NG --steps 4 --no-parity; show bench_tma_1024_r16
SDNS_CHUNKS=4 run bench_tma_1024_k4 bench.py --gpus $NG --steps 4 --no-parity; show bench_tma_1024_k4
SDNS_CHUNKS=10 run bench_tma_1024_k10 bench.py --gpus $NG --steps 4 --no-parity; show bench_tma_1024_k10
SDNS_GRAPH=1 run bench_tma_1024_graph bench.py --gpus $NG --steps 4 --no-parity; show bench_tma_1024_graph
run bench_tma_512 bench.py --gpus $NG --steps 10 --grid 256 --no-parity; show bench_tma_512
SDNS_EXCHANGE=ce run bench_ce_512 bench.py --gpus $NG --steps 10 --grid 256 --no-parity; show bench_ce_512
python profiles/tools/timeline_report.py $O/timeline_tma 2>&1 | head -30
python profiles/tools/timeline_report.py $O/timeline_ce 2>&1 | head -30
