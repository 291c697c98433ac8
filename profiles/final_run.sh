set -x
T=aj
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -4 gpurun_out/smoke_$T.log
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_r01_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
python bench.py > gpurun_out/bench_r01_$T.json 2> gpurun_out/bench_$T.err; tail -c 300 gpurun_out/bench_r01_$T.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ell_|k_gd_|k_policy_|k_value_" -c 400 --csv --log-file gpurun_out/launches_r01_$T.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_launch_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gd_|k_policy_|k_value_" -s 32 -c 32 -o /tmp/mpnn_$T python profiles/mpnn_ncu.py 32 2 > gpurun_out/ncu_mpnn_$T.log 2>&1
ncu -i /tmp/mpnn_$T.ncu-rep --page raw --csv > /tmp/mpnn_raw.csv 2>/dev/null && python profiles/summarise_ncu.py /tmp/mpnn_raw.csv gpurun_out/r01_${T}_mpnn_ncu_full_summary.csv
ncu --set full --clock-control none --import-source on -k regex:"k_gd_sample_bcast|k_ell_|k_withdraw|k_insert_|k_observe" -s 70 -c 14 -o /tmp/roll_$T python profiles/rollout_ncu.py 1024 > gpurun_out/ncu_roll_$T.log 2>&1
ncu -i /tmp/roll_$T.ncu-rep --page raw --csv > /tmp/roll_raw.csv 2>/dev/null && python profiles/summarise_ncu.py /tmp/roll_raw.csv gpurun_out/r01_${T}_rollout_ncu_full_summary.csv
ls -la gpurun_out/*_$T* | tail -12
