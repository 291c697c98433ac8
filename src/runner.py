"""Mirror of the reference's src/runner.py."""
from tarl_simulator_b200.runner import Runner, RunnerArgs  # noqa: F401
