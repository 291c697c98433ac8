set -x
T=r02_v
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant nohint "-DTARL_ABLATE_HINT"
build_variant nopost "-DTARL_ABLATE_POST"
: > gpurun_out/tune_$T.log
for v in base nohint nopost base; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_TUNE=$v TARL_B200_LIB=$lib python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
done
grep -v Warn gpurun_out/tune_$T.log
for v in base nohint; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_B200_LIB=$lib python bench.py --steps 20 --warmup 5 --no-mpnn --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v', d['ms_per_step'], d['roofline']['kernels_ms'], d['e2e']['value']/1e9, d['e2e']['copies_alone'])"
done
