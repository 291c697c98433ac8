set -x
T=r02_r
ncu --set full --import-source on --clock-control none -k regex:"k_value_message_dropout|k_value_edge_grad_dropout|k_value_keep_words|k_value_dz" -s 4 -c 4 -o gpurun_out/prof_$T -f python profiles/value_train_once.py > gpurun_out/ncu_$T.log 2>&1; tail -3 gpurun_out/ncu_$T.log
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv --kernel-name regex:k_value_message_dropout > gpurun_out/src_${T}_message.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv --kernel-name regex:k_value_edge_grad_dropout > gpurun_out/src_${T}_edge_grad.csv 2>/dev/null
ls -la gpurun_out/*$T*
