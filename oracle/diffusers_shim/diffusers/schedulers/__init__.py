from .scheduling_ddim import DDIMScheduler, DDIMSchedulerOutput  # noqa: F401
from .scheduling_ddpm import DDPMScheduler, DDPMSchedulerOutput  # noqa: F401
