set -x
T=r02_d
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py tests/test_core_step_gpu.py tests/test_sim_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
python profiles/tune_step.py 5 20 > gpurun_out/tune_$T.log 2>&1
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant mb8 "-DTARL_SELECT_MINBLOCKS=8"
build_variant mb10 "-DTARL_SELECT_MINBLOCKS=10"
build_variant nogather "-DTARL_ABLATE_GATHER"
for v in mb8 mb10 nogather; do TARL_TUNE="$v" TARL_B200_LIB=/tmp/$v/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1; done
python profiles/tune_step.py 3 20 grid100 1 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
cat gpurun_out/tune_$T.log
for v in base nogather; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_B200_LIB=$lib ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_ell_" -s 20 -c 4 --csv --log-file gpurun_out/launches_${T}_$v.csv python profiles/tune_step.py 1 20 > gpurun_out/ncu_$T.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_${T}_$v.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print("$v", r[4][:42], r[-3], r[-1])
PY
done
python bench.py --steps 20 --no-mpnn --no-ppo > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 2500 gpurun_out/bench_$T.json; tail -5 gpurun_out/bench_$T.err
