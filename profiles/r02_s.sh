set -x
python profiles/ppo_update_prof.py 128 2>&1 | grep -v Warn | tail -45
