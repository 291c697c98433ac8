set -x
T=r02_l
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant noattr "-DTARL_ABLATE_ATTR"
build_variant nogather "-DTARL_ABLATE_GATHER"
: > gpurun_out/tune_$T.log
TARL_TUNE=base python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=noattr TARL_B200_LIB=/tmp/noattr/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=nogather TARL_B200_LIB=/tmp/nogather/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=base python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
grep -v Warn gpurun_out/tune_$T.log
python bench.py --steps 20 --no-mpnn --no-ppo --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_l.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernels_ms'], d['roofline']['frac'], d['roofline']['step']['frac'], d['clocks'])
PY
