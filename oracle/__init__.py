"""CPU oracle for the joint + RNN-T loss hot path of lucadellalib/ts-asr.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  ``tsasr_b200`` never imports it and has no CPU fallback.

Contents
--------
``rnnt_numpy``      float64 restatement of the transducer forward-backward recursion and of
                    the closed-form gradient (both reference semantics), plus the joint chain.
``rnnt_c``          ctypes loader of ``rnnt_oracle.c`` (the same algorithm in plain C, fp32 or
                    fp64 accumulation, OpenMP over utterances) for config-1-sized inputs.
``reference_chain`` the reference's own call path restated with torch ops on CPU tensors:
                    ``Transducer_joint.forward`` -> ``Linear`` -> ``losses.transducer_loss``
                    (``use_torchaudio=True`` -> ``torchaudio.functional.rnnt_loss`` on CPU).
``numba_route``     loads the reference's Numba kernels *by file path* from /root/reference under
                    ``NUMBA_ENABLE_CUDASIM=1`` (authoring container only; used by
                    ``make_golden.py``; never imported on the GPU box).
``make_golden``     regenerates ``tests/golden/*.npz`` from the two reference routes.

Parity pinning: the restatements are checked in ``tests/test_oracle.py`` against
(1) the reference's only known-answer test
    (vendor/speechbrain/tests/unittests/test_losses.py:109-152 -> 2.2478, Numba semantics),
(2) golden vectors produced in the authoring container by the reference's own Numba kernels
    (vendor/speechbrain/speechbrain/nnet/loss/transducer_loss.py) and by
    torchaudio.functional.rnnt_loss reached exactly as
    vendor/speechbrain/speechbrain/nnet/losses.py:58-59,72-79 reaches it, and
(3) torchaudio's CPU rnnt_loss run live (it is installed in the image, 2.11.0+cu128).
"""
