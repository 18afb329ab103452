set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ns_kernel -s 260 -c 1 -o gpurun_out/s4_ns -f \
    python bench.py --steps 20 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:aec_kernel -s 420 -c 1 -o gpurun_out/s4_aec -f \
    python tools/bench_aec.py --steps 30 --warmup 400 --no-ns > /dev/null 2>&1
ls -la gpurun_out
