# BASELINE configs[2] on the final commit: one replica of the 100x100 grid (core step + MPNN blocks), 1024 replicas in
# one store, and the ncu launch list of the single-replica command.
set -x
T=r02_final4
python bench.py --steps 20 --warmup 5 --workload grid100 --no-ppo --no-cpu-baseline > gpurun_out/bench_${T}_grid100.json 2> gpurun_out/bench_${T}_grid100.err; tail -c 200 gpurun_out/bench_${T}_grid100.json
python bench.py --steps 20 --warmup 5 --workload grid100 --replicas 1024 --no-ppo --no-mpnn --no-cpu-baseline > gpurun_out/bench_${T}_grid100_x1024.json 2> gpurun_out/bench_${T}_grid100_x1024.err; tail -c 200 gpurun_out/bench_${T}_grid100_x1024.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${T}_grid100.csv python bench.py --steps 20 --warmup 5 --workload grid100 --no-cpu-baseline --no-ppo --no-mpnn > gpurun_out/ncu_launch_${T}_grid100.log 2>&1
