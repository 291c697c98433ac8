"""Mirror of the reference's src/response_mpnn.py."""
from tarl_simulator_b200.core import ResponseMPNN  # noqa: F401
from tarl_simulator_b200.message_passing import MessagePassing  # noqa: F401
