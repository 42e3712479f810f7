set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 15267 -c 5089 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 400 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/ | tail -5
