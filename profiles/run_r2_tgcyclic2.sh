# round 2, 2 GPUs, the last GPU seconds: the reference's own tests/TG.py (unchanged; its k / w asserts run on rank 0) on the
# drop-in layer with the cyclic axis-1 ownership (SDNS_K1_LAYOUT=cyclic: local_slice, K, K2, masks follow the strided slice).
O=gpurun_out/r2_tgcyclic2; mkdir -p $O
cd /tmp && SDNS_K1_LAYOUT=cyclic PYTHONPATH=$GRAFT_REPO_ROOT timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 -m spectraldns_b200.run $GRAFT_REPO_ROOT/baseline/_ref/tests/TG.py NS > $GRAFT_REPO_ROOT/$O/tg_ns.out 2>&1; echo "TG.py NS cyclic rc=$?"; grep -E "Fastest|Error|assert|Traceback" $GRAFT_REPO_ROOT/$O/tg_ns.out | head -5; tail -3 $GRAFT_REPO_ROOT/$O/tg_ns.out
