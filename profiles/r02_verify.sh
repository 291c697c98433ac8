# Verification of the head commit the way the driver runs it: GPU tests, smoke, both bench arms (N = 1).
set -x
T=${1:-r02_verify}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -2 gpurun_out/smoke_$T.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 400 gpurun_out/bench_$T.json
