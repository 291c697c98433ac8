set -x
for V in "" "TARL_ROLLOUT_DRAW_AT=step" "TARL_ROLLOUT_SIDE_PRIORITY=1" "TARL_ROLLOUT_DRAW_AT=step TARL_ROLLOUT_SIDE_PRIORITY=1" "TARL_NO_ROLLOUT_OVERLAP=1"; do
  echo "== variant: $V"
  env $V python profiles/rollout_timeline.py 128 2>&1 | grep -E "^R |kernel time|checksum|k_ell|k_insert_direct|k_gd_sample"
done
echo "== 1024 replicas"
for V in "" "TARL_ROLLOUT_DRAW_AT=step TARL_ROLLOUT_SIDE_PRIORITY=1"; do
  echo "== variant: $V"
  env $V python profiles/rollout_timeline.py 1024 2>&1 | grep -E "^R |kernel time|checksum"
done
