"""Synthetic weights / inputs moved to the package (``autonomous_driving_with_diffusion_model_b200.synthetic``) in round 2 so
that bench.py's GPU arm and the scripts no longer import anything from ``oracle/``.  This alias keeps the oracle's own
modules and the tests (``from oracle import weights as W``) working.  TEST INFRASTRUCTURE."""
from autonomous_driving_with_diffusion_model_b200.synthetic import *  # noqa: F401,F403
from autonomous_driving_with_diffusion_model_b200.synthetic import (MODES, hash_normal, hash_symmetric, hash_uniform, make_state_dict,  # noqa: F401
                                                                    resnet34_specs, state_dict_digest, synth_image, synth_inputs,
                                                                    traj_predict_specs, unet_specs)
