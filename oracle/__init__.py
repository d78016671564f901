"""CPU oracle for the movenet WaveNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``movenet_b200/`` may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or
as the timed CPU baseline — never as the product path.

Parity status
-------------
* ``wavenet_oracle``: pinned.  Checked bit-for-bit (audio-only) against the
  reference's own ``movenet.wavenet.WaveNet`` imported from /root/reference in
  the authoring container, and against the committed fixtures that import
  produced (``tests/golden/*.pt``, generator ``tests/golden/make_golden.py``).
  The reference ships no golden vectors of its own (its only test asserts a
  shape, /root/reference/tests/test_model.py:60).
* video-conditioned paths: the unmodified reference raises at
  movenet/modules.py:76 (length-T context added to a length-(T-d) tensor), so
  there is nothing executable to pin against.  The oracle applies the same
  right-aligned crop the reference uses for the residual two lines later
  (movenet/modules.py:84).  Those fixtures are "reference + one-line crop":
  PARITY UNPINNED by the reference itself, pinned only to that patched build.
* ``mulaw_oracle``: the arithmetic lives in torchaudio (third-party,
  requirements.txt:12 ``torchaudio>=0.12.0``; 2.11.0 installed when the
  fixtures were made).  Restated from its published formula and pinned to the
  fixtures produced by the installed torchaudio CPU functions.
"""
