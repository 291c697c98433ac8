set -x
for R in 1024 128; do python profiles/ppo_update_prof.py $R 2>&1 | grep -v Warn | tail -34; done
