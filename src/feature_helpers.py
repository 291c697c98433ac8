"""Mirror of the reference's src/feature_helpers.py."""
from tarl_simulator_b200.feature_helpers import AgentFeatureHelpers, FeatureHelpers, ObservationFeatureHelpers  # noqa: F401
