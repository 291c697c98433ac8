"""Import stub for torchrl 0.5.0: lets `src/reinforcement_learning.py` of the reference import so that
`GraphDistribution` and `SimulatorEnv._reset/_step` can be executed. No torchrl algorithm is restated here."""
