# round 2: transfer role after the issue-rate fix (2 ranks): ce vs tma, CTA count and budget sweeps
NG=${1:-2}
O=gpurun_out/r2_xfer${NG}b; mkdir -p $O
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
run slab_tma tests/mp/slab_worker.py; tail -2 $O/slab_tma.out
show() { python - <<PY
import json
try:
    d = json.loads(open("$O/$1.out").read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("$1", d["config"]["workload"], "ms/step %.3f" % d["ms_per_step"], "nvlink sustained", nv.get("sustained_GBps_per_direction", nv.get("sustained_over_step_GBps")), "flushes", nv.get("transfer_only_launches_per_step"))
    print("   " + "  ".join("%s %.0fus x%.0f %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["launches_per_step"], v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$1 FAILED", e, open("$O/$1.err").read()[-1500:])
PY
}
G=${2:-256}
SDNS_EXCHANGE=ce run bench_ce bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_ce
run bench_tma bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_tma
for c in 8 16 64; do SDNS_XCTAS=$c run bench_tma_c$c bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_tma_c$c; done
for r in 0.08 0.3 1.0; do SDNS_XRATIO=$r run bench_tma_r$r bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_tma_r$r; done
for k in 3 4 8; do SDNS_CHUNKS=$k run bench_tma_k$k bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_tma_k$k; done
SDNS_GRAPH=1 run bench_tma_graph bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $G; show bench_tma_graph
run bench_tma_512 bench.py --gpus $NG --no-cpu-baseline --steps 6 --grid 512; show bench_tma_512
