mkdir -p /tmp/noocc && cp build/obj/*.o /tmp/noocc/
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -fmad=false -DTARL_ABLATE_OCC -c tarl_simulator_b200/csrc/engine.cu -o /tmp/noocc/engine.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/noocc/libtarl.so /tmp/noocc/*.o
for R in 128 1024; do
python profiles/rollout_kernels.py $R base 2>&1 | tail -1
TARL_B200_LIB=/tmp/noocc/libtarl.so python profiles/rollout_kernels.py $R noocc 2>&1 | tail -1
done
