"""TEST-ONLY stand-in for the third-party ``diffusers==0.28.0`` package (reference ``requirements.txt:2``).

``diffusers`` is not vendored in the reference and not installed in this image.  The reference's four scheduler
files subclass ``DDIMScheduler`` / ``DDPMScheduler``; this shim restates ONLY the base-class members those files
touch (``__init__`` betas/alphas_cumprod, ``set_timesteps``, ``_get_variance``, ``previous_timestep``,
``_threshold_sample``, ``randn_tensor``, the two output dataclasses) from the published 0.28.0 algorithm so that
the reference ``step()`` bodies can be executed verbatim by ``oracle/make_golden.py``.  Never imported by the
product package.  "Parity unpinned" for this third-party part (no source / fixtures available offline).
"""
__version__ = "0.28.0+shim"
