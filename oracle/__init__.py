"""CPU oracle for the Ravvent inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / the timed CPU baseline.  The product (``ravvent_basecaller_b200``)
never imports this package and fails loudly when its CUDA library is missing.

Parity status:
  * event detection + snippet builder: PINNED against the reference's own
    modules, imported from /root/reference in the build container
    (tools/make_golden.py -> tests/golden/*.npz).
  * NN path (encoders, attention decoder, greedy / beam search): PARITY
    UNPINNED.  The arithmetic lives in tensorflow / tensorflow_addons, neither
    of which is vendored in the reference nor installable offline; the oracle
    restates the published Keras / TFA algorithms and is cross-checked against
    torch.nn.LSTM for the encoder cells.
"""
