"""Mirror of the reference's src/transportation_simulator.py."""
from tarl_simulator_b200.transportation_simulator import TransportationSimulator  # noqa: F401
