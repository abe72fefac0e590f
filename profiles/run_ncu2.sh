# usage: bash profiles/run_ncu2.sh <tag> [bench args]   (under gpurun): ncu --set full of one right-hand side (5 kernels)
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"strided_kernel|z_kernel|zx_kernel|zy_kernel|f0x_kernel|mhd_f0_kernel" -s 65 -c 5 -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log | cut -c1-300
