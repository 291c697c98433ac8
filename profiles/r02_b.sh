set -x
T=r02_b
python profiles/tune_step.py 5 20 > gpurun_out/tune_$T.log 2>&1
for mb in 9 10 12; do
  mkdir -p /tmp/v$mb && cp build/obj/*.o /tmp/v$mb/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false -DTARL_SELECT_MINBLOCKS=$mb -c tarl_simulator_b200/csrc/engine.cu -o /tmp/v$mb/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/v$mb/libtarl.so /tmp/v$mb/*.o
  TARL_TUNE="minblocks=$mb" TARL_B200_LIB=/tmp/v$mb/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
done
TARL_TUNE="no-pdl" TARL_NO_PDL=1 python profiles/tune_step.py 3 20 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 100 ring_radial_1m 1 100 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 20 grid100 1 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
cat gpurun_out/tune_$T.log
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"k_ell_|k_deferred" -s 30 -c 30 --csv --log-file gpurun_out/launches_$T.csv python profiles/tune_step.py 1 20 > gpurun_out/ncu_$T.log 2>&1
tail -32 gpurun_out/launches_$T.csv | cut -d, -f5,10- | head -40
