# round 2: transfer role with the CTA count derived from the ring capacity; new bench.py (parity gate, 512 default)
NG=${1:-2}
O=gpurun_out/r2_xfer${NG}c; mkdir -p $O
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 "$@" > $O/$tag.out 2> $O/$tag.err; echo "$tag rc=$?"; }
show() { python - <<PY
import json
try:
    d = json.loads(open("$O/$1.out").read().strip().splitlines()[-1])
    nv = d.get("nvlink", {})
    print("$1", d["config"]["workload"], "ms/step %.3f" % d["ms_per_step"], "nvlink", nv.get("sustained_GBps_per_direction", nv.get("sustained_over_step_GBps")), "flushes", nv.get("transfer_only_launches_per_step"), "parity", (d.get("parity") or {}).get("worst_err_over_tol_all_ranks"))
    print("   " + "  ".join("%s %.0fus x%.0f %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["launches_per_step"], v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$1 FAILED", e, open("$O/$1.err").read()[-1500:])
PY
}
run bench_tma bench.py --gpus $NG --steps 6; show bench_tma
SDNS_EXCHANGE=ce run bench_ce bench.py --gpus $NG --steps 6 --no-parity; show bench_ce
SDNS_XINFLIGHT_KB=1024 run bench_tma_i1 bench.py --gpus $NG --steps 6 --no-parity; show bench_tma_i1
SDNS_XINFLIGHT_KB=4096 run bench_tma_i4 bench.py --gpus $NG --steps 6 --no-parity; show bench_tma_i4
SDNS_XRATIO=0.3 run bench_tma_r3 bench.py --gpus $NG --steps 6 --no-parity; show bench_tma_r3
SDNS_CHUNKS=4 run bench_tma_k4 bench.py --gpus $NG --steps 6 --no-parity; show bench_tma_k4
run bench_tma_256 bench.py --gpus $NG --steps 10 --grid 256 --no-parity; show bench_tma_256
