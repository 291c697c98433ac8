"""Mirror of the reference's src/agents/base.py."""
from tarl_simulator_b200.agents import Agents  # noqa: F401
