"""Mirror of the reference's src/direction_mpnn.py."""
from tarl_simulator_b200.core import DirectionMPNN  # noqa: F401
from tarl_simulator_b200.message_passing import MessagePassing  # noqa: F401
