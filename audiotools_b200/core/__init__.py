from . import util
from .audio_signal import AudioSignal
from .audio_signal import STFTParams
from .loudness import Meter
