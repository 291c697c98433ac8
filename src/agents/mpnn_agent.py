"""Mirror of the reference's src/agents/mpnn_agent.py."""
from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet, MPNNValueNetSimple  # noqa: F401
