set -x
T=r02_m
build_variant() { # name, flags
  mkdir -p /tmp/$1 && cp build/obj/*.o /tmp/$1/
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false $2 -c tarl_simulator_b200/csrc/engine.cu -o /tmp/$1/engine.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/$1/libtarl.so /tmp/$1/*.o
}
build_variant nocontest "-DTARL_ABLATE_CONTEST"
build_variant nophilox "-DTARL_ABLATE_PHILOX"
: > gpurun_out/tune_$T.log
TARL_TUNE=base python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=nocontest TARL_B200_LIB=/tmp/nocontest/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=nophilox TARL_B200_LIB=/tmp/nophilox/libtarl.so python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=base python profiles/tune_step.py 5 20 >> gpurun_out/tune_$T.log 2>&1
grep -v Warn gpurun_out/tune_$T.log
for v in base nocontest; do
  lib=/tmp/$v/libtarl.so; [ $v = base ] && lib=tarl_simulator_b200/libtarl_b200.so
  TARL_B200_LIB=$lib ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --cache-control none --clock-control none -k regex:"k_ell_" -s 20 -c 2 --csv --log-file gpurun_out/launches_${T}_$v.csv python profiles/tune_step.py 1 20 > gpurun_out/ncu_$T.log 2>&1
  grep -E "k_ell" gpurun_out/launches_${T}_$v.csv | awk -F'","' '{print "'$v'", substr($5,1,40), $(NF-2), $NF}'
done
