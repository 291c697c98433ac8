"""Mirror of the reference's src/rl/ppo_trainer.py."""
from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, ppo_train  # noqa: F401
