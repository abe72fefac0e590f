# round 2, first 1-GPU call: baseline of the round-1 code on this round's box, plus everything round 1 left unrun
# (60/90 lengths, the fp32 column-pair kernels, VV / MHD timings, 512^3).
mkdir -p gpurun_out/r2_first
O=gpurun_out/r2_first
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
SDNS_TEST_NEW=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
b() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["config"]["workload"], "ms/step %.3f step_frac %.3f e2e %.2f" % (d["ms_per_step"], d["roofline"]["step_frac"], d["e2e"]["ms_per_step"]))
    print("   " + "  ".join("%s %.0fus %.0fGB/s" % (k, v["ms_per_launch"]*1e3, v["GBps"]) for k, v in d["roofline"]["all_kernels"].items()))
except Exception as e:
    print("$tag FAILED", e, open("$O/bench_$tag.err").read()[-1500:])
PY
}
b 256d --grid 256 --steps 20
b 512d --grid 512 --steps 10
b 512s32 --grid 512 --precision single --dealias 3/2-rule --steps 10
b 512s23 --grid 512 --precision single --steps 10
b vv256d --grid 256 --solver VV --steps 20
b mhd256d --grid 256 --solver MHD --steps 10
export SDNS_LIBPATH=$PWD/spectraldns_b200/variants/libsdns_pairs.so
b 512s32_pairs --grid 512 --precision single --dealias 3/2-rule --steps 10
b 512s23_pairs --grid 512 --precision single --steps 10
SDNS_TEST_NEW=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single or float or f32 or golden" > $O/pytest_pairs.log 2>&1; echo "pytest pairs rc=$?"; tail -3 $O/pytest_pairs.log
