set -x
T=r02_n
: > gpurun_out/tune_$T.log
for w in ring_radial_250k ring_radial_1m ring_radial_4m; do
TARL_TUNE=base python profiles/tune_step.py 4 20 $w >> gpurun_out/tune_$T.log 2>&1
TARL_TUNE=pol00 TARL_L2_POLICY=00 TARL_AHEAD_SELECT=0 TARL_AHEAD_RESPOND=0 python profiles/tune_step.py 4 20 $w >> gpurun_out/tune_$T.log 2>&1
done
grep -v Warn gpurun_out/tune_$T.log
