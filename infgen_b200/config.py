"""Static hyper-parameters of the decode hot path.

Values are the ones `configs/ours_standard.yaml` / `configs/ours_long_term.yaml` feed into
`InfGenAgentDecoder.__init__` (reference `infgen/modules/agent_decoder.py:100-314`) plus the constants that
constructor hard-codes.  The CUDA library is compiled for exactly these widths (hidden 128, 8 heads x 16,
64 Fourier bands); everything that is a run-time knob in the reference stays a run-time knob here.
"""
from dataclasses import dataclass, field, asdict
from typing import Dict

HIDDEN = 128
NUM_HEADS = 8
HEAD_DIM = 16
NUM_FREQ = 64
FOURIER_IN = 2 * NUM_FREQ + 1       # cos | sin | x   (layers.py:153)
NUM_LAYERS = 6                      # num_agent_layers (ours_standard.yaml:76)
SEED_LAYERS = 3                     # agent_decoder.py:235
TOKEN_SIZE = 2048                   # ours_standard.yaml:6
SHIFT = 5                           # agent_decoder.py:299 / preprocess.py:13
NUM_HIST_STEPS = 11
AGENT_TYPE = ['veh', 'ped', 'cyc', 'seed']          # preprocess.py:19
AGENT_SHAPE = {                                      # preprocess.py:14-18 (key spelling is the reference's)
    'vehicle': [4.3, 1.8, 1.0],
    'pedstrain': [0.5, 0.5, 1.0],
    'cyclist': [1.9, 0.5, 1.0],
}
STATE_TOKEN = {'invalid': 0, 'valid': 1, 'enter': 2, 'exit': 3}   # ours_standard.yaml:11-15


@dataclass
class DecoderConfig:
    """Run-time configuration mirrored from the reference constructor arguments."""
    num_historical_steps: int = NUM_HIST_STEPS
    time_span: int = 60                 # -> temporal window time_span / shift = 12 columns
    pl2a_radius: float = 30.0
    pl2seed_radius: float = 75.0
    a2a_radius: float = 60.0
    a2sa_radius: float = 10.0
    pl2sa_radius: float = 10.0
    grid_range: float = 150.0
    grid_interval: float = 3.0
    angle_interval: float = 3.0
    num_recurrent_steps_val: int = -1   # -1: scene length - num_historical_steps (agent_decoder.py:1633-1635)
    disable_insertion: bool = False
    use_state_token: bool = True
    motion_beam_size: int = 5           # agent_decoder.py:300 (1 = greedy)
    insert_beam_size: int = 10          # agent_decoder.py:301
    debug_force_enter: bool = False     # the reference's DEBUG=1 env switch: seed head forced to 'enter' (:1888-1889)
    insert_row_reserve: int = 64        # rows kept free per scene for agents the insertion stage appends
    num_seed_feature: int = 10          # agent_decoder.py:292 (the "last 10 rows" temporal-edge quirk)
    max_pl2a_neighbors: int = 5         # agent_decoder.py:711
    max_a2a_neighbors: int = 300        # agent_decoder.py:633
    state_token: Dict[str, int] = field(default_factory=lambda: dict(STATE_TOKEN))

    @property
    def shift(self) -> int:
        return SHIFT

    @property
    def hist_cols(self) -> int:
        """Number of history token columns: (num_historical_steps - 1) // shift = 2."""
        return (self.num_historical_steps - 1) // SHIFT

    @property
    def window(self) -> int:
        return int(self.time_span // SHIFT)

    def to_dict(self):
        return asdict(self)
