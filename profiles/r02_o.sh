set -x
T=r02_o
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -3 gpurun_out/smoke_$T.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -3 gpurun_out/bench_$T.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err; tail -2 gpurun_out/bench_${T}_ref.err; cat gpurun_out/bench_${T}_ref.json | cut -c1-600
