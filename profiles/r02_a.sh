# round 2, first GPU call: parity of the reworked direction phase, variants of the streaming kernel, short bench
set -x
T=r02_a
python -m pytest tests/test_store_replay_gpu.py tests/test_link_store_gpu.py tests/test_core_step_gpu.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -15 gpurun_out/pytest_$T.log
python profiles/tune_step.py 5 300 > gpurun_out/tune_$T.log 2>&1
for mb in 9 10 12; do
  mkdir -p /tmp/v$mb && for f in tarl_simulator_b200/csrc/*.cu; do
    b=$(basename $f .cu); extra=""; case $b in engine|agents|core_step) extra="-fmad=false";; esac
    if [ $b = engine ]; then nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include $extra -DTARL_SELECT_MINBLOCKS=$mb -c $f -o /tmp/v$mb/$b.o; else cp build/obj/$b.o /tmp/v$mb/$b.o; fi
  done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/v$mb/libtarl.so /tmp/v$mb/*.o
  TARL_TUNE="minblocks=$mb" TARL_B200_LIB=/tmp/v$mb/libtarl.so python profiles/tune_step.py 5 300 >> gpurun_out/tune_$T.log 2>&1
done
TARL_TUNE="no-pdl" TARL_NO_PDL=1 python profiles/tune_step.py 3 300 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 100 grid100 1 >> gpurun_out/tune_$T.log 2>&1
python profiles/tune_step.py 3 20 grid100 1024 >> gpurun_out/tune_$T.log 2>&1
cat gpurun_out/tune_$T.log
python bench.py --steps 200 --no-mpnn --no-ppo > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 3000 gpurun_out/bench_$T.json; tail -5 gpurun_out/bench_$T.err
