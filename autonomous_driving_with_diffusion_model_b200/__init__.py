"""B200-native diffusion-planning hot path (drop-in for the reference's model forward / scheduler.step / generate_traj)."""
from .constant import GuidanceType
from .config import load_cfg, scheduler_kwargs
from .guidance import GuidanceLoss, TargetGuidance
from .modeling import TemporalMapUnet, build_model
from .planner import DiffusionPlanner
from .scheduler import (SCHEDULER_FUNC, GuidanceDDIMScheduler, GuidanceDDPMScheduler, InpaintingDDIMScheduler,
                        InpaintingDDPMScheduler)
from .sharding import shard, shard_bounds
from .checkpoint import copy_parameters, load_checkpoint
from .control import Controller, FleetController, PIDController, post_process_control, post_process_control_batch
from .inputs import preprocess_frames, process_next_waypoint

__all__ = ["GuidanceType", "load_cfg", "scheduler_kwargs", "GuidanceLoss", "TargetGuidance", "TemporalMapUnet", "build_model",
           "DiffusionPlanner", "SCHEDULER_FUNC", "GuidanceDDIMScheduler", "GuidanceDDPMScheduler", "InpaintingDDIMScheduler",
           "InpaintingDDPMScheduler", "shard", "shard_bounds", "copy_parameters", "load_checkpoint", "Controller", "FleetController", "PIDController",
           "post_process_control", "post_process_control_batch", "preprocess_frames", "process_next_waypoint"]
