"""Mirror of the reference's src/simulation_core_model.py."""
from tarl_simulator_b200.core import SimulationCoreModel  # noqa: F401
from tarl_simulator_b200.data import Data  # noqa: F401
