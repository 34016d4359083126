"""``DSPMixin``: per-item windowed-sinc low-/high-pass (ref:audiotools/core/dsp.py:153-215).

The reference loops over the batch in Python, building one ``julius.LowPassFilter`` per item;
here the per-item tap design and the filtering are one grouped launch each (``csrc/fir.cu``).
Windowing / overlap-add and the spectral masks of the reference's DSPMixin are "next" tier
(SURVEY.md §8f)."""
from . import util


def _engine():
    from ..engine import get_engine

    return get_engine()


class DSPMixin:
    def low_pass(self, cutoffs, zeros: int = 51):
        """Low-pass each item at its own cutoff (Hz)."""
        cutoffs = util.ensure_tensor(cutoffs, 2, self.batch_size)
        self.audio_data = _engine().sinc_filter(self._materialized(), cutoffs[:, 0], self.sample_rate, zeros,
                                                highpass=False)
        self.stft_data = None
        return self

    def high_pass(self, cutoffs, zeros: int = 51):
        """High-pass each item at its own cutoff (Hz): ``x - low_pass(x)``."""
        cutoffs = util.ensure_tensor(cutoffs, 2, self.batch_size)
        self.audio_data = _engine().sinc_filter(self._materialized(), cutoffs[:, 0], self.sample_rate, zeros,
                                                highpass=True)
        self.stft_data = None
        return self
