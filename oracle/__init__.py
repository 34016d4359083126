"""CPU oracle for the AudioSignal DSP hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm for the hot
path named in BASELINE.json (descriptinc/audiotools @ 348ebf2, v0.7.4).  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``audiotools_b200/`` imports it, and the product path raises if
the CUDA library is missing instead of falling back to this code.

Layout
------
``third_party.py``  restated arithmetic of the *absent* dependencies
                    (julius 0.2.7, pyloudnorm 0.1.1, librosa 0.10 ``filters.mel``).
``signal_path.py``  restated first-party code of the reference
                    (``audiotools/core/{audio_signal,loudness,effects,dsp}.py``).

Pinning status (see DESIGN.md "Oracle")
---------------------------------------
* First-party code: PINNED.  ``tests/golden/make_golden.py`` imports the real
  reference from ``/root/reference`` (with the absent third-party modules
  shimmed by ``third_party.py``) and stores its outputs under ``tests/golden``
  (``make_golden_spectral.py`` does the same for the spectral masks and the
  SpectralTransform family); ``tests/test_oracle_golden.py`` checks
  ``signal_path.py`` against them.
* ``torch.stft`` / ``torchaudio.functional.lfilter`` / ``scipy.signal``: the
  oracle calls the very same installed functions the reference calls.
* julius / pyloudnorm / librosa restatements: pinned only by the reference's
  own property tests and known-answer constants (ITU-R BS.1770 coefficient
  table, torchaudio's Slaney mel filterbank, low/high-pass sine thresholds,
  band-sum identity) -- "parity unpinned" beyond those, because their source is
  not under /root/reference and they cannot be installed here (no network).
* SoX ``pitch``: no numeric pin exists anywhere in the reference
  (self-consistency tests only) -- "parity unpinned".
"""
