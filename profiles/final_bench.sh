T=ak
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -4 gpurun_out/smoke_$T.log
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_r01_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
python bench.py > gpurun_out/bench_r01_$T.json 2> gpurun_out/bench_$T.err; wc -l gpurun_out/bench_r01_$T.json; tail -c 200 gpurun_out/bench_r01_$T.json
