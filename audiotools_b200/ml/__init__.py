"""Only the one layer of the reference's ``ml`` package that sits on the AudioSignal transform path
(``SpectralGate``, used by ``transforms.SpectralDenoising``); models, trainers and accelerators are out of scope."""
from . import layers
