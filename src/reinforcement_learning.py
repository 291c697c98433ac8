"""Mirror of the reference's src/reinforcement_learning.py."""
from tarl_simulator_b200.distribution import GraphDistribution  # noqa: F401
from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv, SimulatorEnv  # noqa: F401
