set -x
T=r02_c
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
python profiles/tune_step.py 5 20 > gpurun_out/tune_$T.log 2>&1
for mb in 9 12; do
  mkdir -p /tmp/v$mb && cp build/obj/*.o /tmp/v$mb/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false -DTARL_SELECT_MINBLOCKS=$mb -c tarl_simulator_b200/csrc/engine.cu -o /tmp/v$mb/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/v$mb/libtarl.so /tmp/v$mb/*.o
  TARL_TUNE="minblocks=$mb" TARL_B200_LIB=/tmp/v$mb/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
done
python profiles/tune_step.py 3 20 grid100 1 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
cat gpurun_out/tune_$T.log
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_ell_|k_deferred" -s 30 -c 6 --csv --log-file gpurun_out/launches_$T.csv python profiles/tune_step.py 1 20 > gpurun_out/ncu_$T.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r02_c.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:50], r[-3], r[-1])
PY
