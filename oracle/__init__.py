"""CPU oracle for the diffusion-planning hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``autonomous_driving_with_diffusion_model_b200``) may import
this package.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.

What is in here
---------------
* ``weights``       deterministic, torch-RNG-independent random-init ``state_dict`` with the reference's
                    key names / shapes / registration order (``modeling/temporal.py:59-195``).
* ``unet``          functional fp32 restatement of ``TemporalMapUnet.forward`` (``modeling/temporal.py:197-245``),
                    its blocks (``modeling/helpers.py:62-112``), ``TrajPredict`` (``modeling/helpers.py:22-59``)
                    and the ResNet-34 conditioning encoder (``modeling/resnet.py:163-296``).
* ``schedulers``    restatement of the four ``scheduler/*.py`` ``step()`` bodies on top of a restated
                    ``diffusers==0.28.0`` DDIM/DDPM base (third party, not vendored in the reference).
* ``guidance``      ``control/guidance.py:35-59`` + ``control/guidance_loss.py:10-22``.
* ``plan``          the sampling loop ``interact.py:115-168`` generalised to a batch.
* ``diffusers_shim``/``reference_loader``/``make_golden``: used ONLY in the build container, where
                    ``/root/reference`` exists, to run the real reference code and (a) validate this restatement,
                    (b) write the golden vectors under ``tests/golden/``.

Parity status: the reference has no tests / golden vectors of its own, so the oracle is pinned
against outputs of the reference itself executed in the build container (``oracle/make_golden.py``); the
``diffusers`` base-class arithmetic is a restatement of the published 0.28.0 algorithm ("parity unpinned" for
that third-party part: its source is not in the container), self-checked against the known-answer constants
in ``tests/test_oracle.py``.
"""
