"""Column layout contracts of `graph.x`, `agent_features` and the observation tensor.

Same attribute names and values as the reference's src/feature_helpers.py:38-92 (they are the data format of the
drop-in boundary: user code indexes `graph.x` with them).
"""


class FeatureHelpers:
    """Row layout of `graph.x` (width 3*Nmax+7): three FIFO segments of Nmax slots each — agent ids (head at column
    0), arrival times, scheduled exit times — followed by MAXN, NUM, FFTT, LENGTH, MAX_FLOW, SELECTED_ROAD,
    ROAD_INDEX. (src/feature_helpers.py:38-54; NODE_TYPE is declared there too although rows stop one column
    earlier, src/transportation_simulator.py:135 — kept for attribute compatibility.)"""

    def __init__(self, Nmax=100):
        self.Nmax = Nmax
        self.AGENT_POSITION = slice(0, Nmax)
        self.AGENT_TIME_ARRIVAL = slice(Nmax, 2 * Nmax)
        self.AGENT_TIME_DEPARTURE = slice(2 * Nmax, 3 * Nmax)
        base = 3 * Nmax
        (self.MAX_NUMBER_OF_AGENT, self.NUMBER_OF_AGENT, self.FREE_FLOW_TIME_TRAVEL, self.LENGHT_OF_ROAD,
         self.MAX_FLOW, self.SELECTED_ROAD, self.ROAD_INDEX, self.NODE_TYPE) = range(base, base + 8)
        self.HEAD_FIFO = 0
        self.HEAD_FIFO_ARRIVAL_TIME = Nmax
        self.HEAD_FIFO_DEPARTURE_TIME = 2 * Nmax
        self.CONGESTION_FILE = 3  # jam buffer: slots kept free for gridlock resolution

    @property
    def num_features(self):
        return 3 * self.Nmax + 7


class AgentFeatureHelpers:
    """Columns of `agent_features` [A+1, 9] (src/feature_helpers.py:56-71). Row 0 is a dummy agent."""

    (ORIGIN, DESTINATION, DEPARTURE_TIME, ARRIVAL_TIME, AGE, SEX, EMPLOYMENT_STATUS, ON_WAY, DONE) = range(9)

    def __init__(self):
        pass

    def __len__(self):
        return 9


class ObservationFeatureHelpers:
    """Columns of the 16-wide observation = 7 link statics ‖ 9 agent features (src/feature_helpers.py:74-92)."""

    (MAX_NUMBER_OF_AGENT, NUMBER_OF_AGENT, FREE_FLOW_TIME_TRAVEL, LENGHT_OF_ROAD, MAX_FLOW, SELECTED_ROAD, ROAD_INDEX,
     ORIGIN, DESTINATION, DEPARTURE_TIME, ARRIVAL_TIME, AGE, SEX, EMPLOYMENT_STATUS, ON_WAY, DONE) = range(16)

    def __init__(self):
        pass
