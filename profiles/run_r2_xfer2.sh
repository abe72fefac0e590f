# round 2: first run of the transfer-role exchange on real GPUs (2 ranks): parity, then ce vs tma vs ldst
NG=${1:-2}
O=gpurun_out/r2_xfer$NG; mkdir -p $O
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
run slab_tma tests/mp/slab_worker.py; tail -12 $O/slab_tma.out
SDNS_EXCHANGE=ldst run slab_ldst tests/mp/slab_worker.py; tail -2 $O/slab_ldst.out
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
for g in 256 512; do
for x in ce tma ldst; do
  SDNS_EXCHANGE=$x run bench_${x}_$g bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid $g; show bench_${x}_$g
done
done
SDNS_XRATIO=0.3 run bench_tma_r30 bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid 256; show bench_tma_r30
SDNS_XRATIO=0.08 run bench_tma_r08 bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid 256; show bench_tma_r08
SDNS_XCTAS=64 run bench_tma_c64 bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid 256; show bench_tma_c64
SDNS_XCTAS=16 run bench_tma_c16 bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid 256; show bench_tma_c16
SDNS_GRAPH=1 run bench_tma_graph bench.py --gpus $NG --no-cpu-baseline --steps 10 --grid 256; show bench_tma_graph
