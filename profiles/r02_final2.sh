# Last single-GPU set of the round (after the uniform-weight hint and the PPO loss kernel): both bench arms, the ncu
# launch list of the bench command, ncu --set full of the two store kernels (-> traffic json). Run under gpurun.
set -x
T=${1:-r02_final2}
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 300 gpurun_out/bench_$T.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ppo --no-mpnn > gpurun_out/ncu_launch_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ell_" -s 20 -c 2 -o /tmp/store_$T -f python profiles/tune_step.py 1 20 > gpurun_out/ncu_store_$T.log 2>&1
ncu -i /tmp/store_$T.ncu-rep --page raw --csv > gpurun_out/store_${T}_raw.csv 2>/dev/null && python profiles/summarise_ncu.py gpurun_out/store_${T}_raw.csv gpurun_out/${T}_store_ell_ncu_full_summary.csv && python profiles/make_traffic.py gpurun_out/store_${T}_raw.csv gpurun_out/traffic_$T.json ring_radial_1m 1 "profiles/${T}_store_ell_ncu_full_summary.csv (ncu --set full --clock-control none, default cache control: every launch measured cold, i.e. WITHOUT the L2 residency the pipelined step has; one launch each)"
ls -la gpurun_out/*$T* | tail
