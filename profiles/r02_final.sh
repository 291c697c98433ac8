# The round's final single-GPU measurements (run under gpurun from the repo root): tests, smoke, bench lines for the
# three single-GPU configurations of BASELINE.json, reference arm, ncu launch list of the bench command, ncu --set full of
# the two store kernels (-> profiles/traffic_r02.json), of the MPNN kernels and of the train-mode kernels.
set -x
T=${1:-r02_final}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -3 gpurun_out/smoke_$T.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 300 gpurun_out/bench_$T.json
python bench.py --steps 20 --warmup 5 --workload grid100 --no-ppo --no-cpu-baseline > gpurun_out/bench_${T}_grid100.json 2> gpurun_out/bench_${T}_grid100.err; tail -c 200 gpurun_out/bench_${T}_grid100.json
python bench.py --steps 20 --warmup 5 --workload grid100 --replicas 1024 --no-ppo --no-mpnn --no-cpu-baseline > gpurun_out/bench_${T}_grid100_x1024.json 2> gpurun_out/bench_${T}_grid100_x1024.err; tail -c 200 gpurun_out/bench_${T}_grid100_x1024.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ppo > gpurun_out/ncu_launch_$T.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${T}_grid100.csv python bench.py --steps 20 --warmup 5 --workload grid100 --no-cpu-baseline --no-ppo --no-mpnn > gpurun_out/ncu_launch_${T}_grid100.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ell_" -s 20 -c 2 -o /tmp/store_$T -f python profiles/tune_step.py 1 20 > gpurun_out/ncu_store_$T.log 2>&1
ncu -i /tmp/store_$T.ncu-rep --page raw --csv > gpurun_out/store_${T}_raw.csv 2>/dev/null && python profiles/summarise_ncu.py gpurun_out/store_${T}_raw.csv gpurun_out/${T}_store_ell_ncu_full_summary.csv && python profiles/make_traffic.py gpurun_out/store_${T}_raw.csv gpurun_out/traffic_$T.json ring_radial_1m 1 "profiles/${T}_store_ell_ncu_full_summary.csv (ncu --set full --clock-control none, default cache control: every launch measured cold, i.e. WITHOUT the L2 residency the pipelined step has; one launch each)"
ncu --set full --clock-control none --import-source on -k regex:"k_gd_|k_policy_|k_value_" -s 32 -c 40 -o /tmp/mpnn_$T -f python profiles/mpnn_ncu.py 32 2 > gpurun_out/ncu_mpnn_$T.log 2>&1
ncu -i /tmp/mpnn_$T.ncu-rep --page raw --csv > /tmp/mpnn_raw.csv 2>/dev/null && python profiles/summarise_ncu.py /tmp/mpnn_raw.csv gpurun_out/${T}_mpnn_ncu_full_summary.csv
ncu --set full --clock-control none -k regex:"k_value_" -s 7 -c 7 -o /tmp/train_$T -f python profiles/value_train_once.py > gpurun_out/ncu_train_$T.log 2>&1
ncu -i /tmp/train_$T.ncu-rep --page raw --csv > /tmp/train_raw.csv 2>/dev/null && python profiles/summarise_ncu.py /tmp/train_raw.csv gpurun_out/${T}_value_train_ncu_full_summary.csv
ncu --set full --clock-control none -k regex:"k_edge_mlp" -c 6 -o /tmp/emlp_$T -f python profiles/edge_mlp_once.py > gpurun_out/ncu_emlp_$T.log 2>&1
ncu -i /tmp/emlp_$T.ncu-rep --page raw --csv > /tmp/emlp_raw.csv 2>/dev/null && python profiles/summarise_ncu.py /tmp/emlp_raw.csv gpurun_out/${T}_edge_mlp_ncu_full_summary.csv
python profiles/rollout_timeline.py 128 2>&1 | grep -v Warn | tail -20 > gpurun_out/rollout_timeline_${T}_128.txt
python profiles/rollout_timeline.py 1024 2>&1 | grep -v Warn | tail -20 > gpurun_out/rollout_timeline_${T}_1024.txt
python profiles/pcie_diag.py
python profiles/value_mlp_time.py 1024 59600 int 2>&1 | grep 'k_value_mlp\|forward\|host' > gpurun_out/value_mlp_time_$T.txt
python profiles/value_mlp_time.py 8192 59600 int 2>&1 | grep 'k_value_mlp\|forward\|host' >> gpurun_out/value_mlp_time_$T.txt
./profiles/micro/mma_rate > gpurun_out/mma_rate_$T.txt 2>&1
ls -la gpurun_out/*$T* | tail -30
