"""CPU oracles for the vocoder hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker.  The product path
(``tts_cube_b200``) never imports this package and fails loudly when the CUDA
library is missing.

Parity status:
  * Path H (HiFi-GAN generator): PINNED - ``hifigan_ref`` is checked against the
    reference ``hifigan/models.py:Generator`` executed in the build container
    (fixtures under tests/golden/, made by ``oracle/make_goldens.py``).
  * Path W heads (mu-law / RAW / MoL / Gaussian): PINNED against
    ``cube/networks/loss.py`` executed in the build container.
  * Upsamplers (UpsampleNet2/R/I): PINNED against ``cube/networks/modules.py``.
  * Path C (ClariNet teacher / ParallelWaveNet student): PARITY UNPINNED by the
    reference - the forward code is not in the reference tree (weights only);
    ``clarinet_ref`` restates upstream ksw0306/ClariNet (no pinned commit) and is
    anchored by (a) strict key/shape match with the shipped checkpoints and
    (b) the teacher-NLL self-consistency probe (tests/test_oracle.py::test_clarinet_self_consistency_probe).
"""
