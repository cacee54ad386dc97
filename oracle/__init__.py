"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, float64) of the spectral-gating hot path of timsainb/noisereduce
(reference @ 51c8534, v3.0.3).  It exists to *check* the CUDA path, never to serve it:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import anything from this package;
  * the product package ``noisereduce_b200`` never imports it and has no CPU fallback --
    it raises if the CUDA library is missing.

Parity status: PINNED.  The reference ships no golden vectors (its tests have no assertions,
SURVEY.md section 8c), so the oracle is pinned against outputs of the reference itself, executed
in the build container by ``tests/golden/make_golden.py`` (imports /root/reference read-only,
numpy 2.3.5 / scipy 1.18.1 / torch 2.11.0) and committed under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays those fixtures against this restatement on every run.

The reference's arithmetic lives in un-vendored third-party code (scipy.signal.stft / istft /
fftconvolve / filtfilt, un-pinned in the reference's setup.py:24).  ``spectral_gate_oracle`` restates
those routines' published algorithms explicitly (framing, window, one-sided DFT, overlap-add,
zero-padded 2-D FIR, forward/backward one-pole IIR); the only library primitive it keeps is the
DFT itself (``scipy.fft.rfft/irfft`` = pocketfft, the same primitive scipy.signal reaches), which
is pinned separately against an O(N^2) matrix DFT in the tests.
"""
