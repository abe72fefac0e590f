# usage: bash profiles/run_ncu.sh <tag> [bench args]   (under gpurun): ncu --set full of one RHS worth of kernels
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"strided_kernel|z_kernel" -s 60 -c 5 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
