# usage: bash profiles/run_timeline.sh <ngpus> [bench args]   (under gpurun --gpus N)
# One RK4 step with every kernel, peer copy and barrier time-stamped on the device clock (bench.py --timeline), for
# the copy-engine exchange, its copy-kernel variant and the fused peer stores, then the per-rank summary: how much of the copy time lies
# underneath kernels and where the plan stream idles.  First thing to run in round 2.
NG=${1:-8}; shift
mkdir -p gpurun_out
for x in ce kcopy store; do
  SDNS_EXCHANGE=$x timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $NG --no-cpu-baseline --steps 10 --timeline gpurun_out/timeline_${x}_g$NG "$@" > gpurun_out/bench_tl_${x}_g$NG.json 2> gpurun_out/bench_tl_${x}_g$NG.err
  echo "== exchange=$x"; python profiles/tools/timeline_report.py gpurun_out/timeline_${x}_g$NG | head -40
done
# the same bench line with the step replayed as a CUDA graph (no timeline: profiling bypasses the graph)
SDNS_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $NG --no-cpu-baseline --steps 10 "$@" > gpurun_out/bench_tl_graph_g$NG.json 2> gpurun_out/bench_tl_graph_g$NG.err
python - <<PY
import json
for x in ('ce', 'kcopy', 'store', 'graph'):
    try:
        d = json.loads(open('gpurun_out/bench_tl_%s_g$NG.json' % x).read().strip().splitlines()[-1])
        print('%-6s ms/step %.3f' % (x, d['ms_per_step']))
    except Exception as e:
        print(x, 'FAILED', e)
PY
