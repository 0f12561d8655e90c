import torch

from ._common import SchedulerBase, SchedulerOutput


class DDPMSchedulerOutput(SchedulerOutput):
    pass


class DDPMScheduler(SchedulerBase):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, variance_type="fixed_small", clip_sample=True, prediction_type="epsilon",
                 thresholding=False, dynamic_thresholding_ratio=0.995, clip_sample_range=1.0,
                 sample_max_value=1.0, timestep_spacing="leading", steps_offset=0, rescale_betas_zero_snr=False):
        self._register(**{k: v for k, v in locals().items() if k not in ("self", "__class__")})
        self._init_tables()
        self.one = torch.tensor(1.0)
        self.custom_timesteps = False
        self.variance_type = variance_type

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        assert timesteps is None and self.config.timestep_spacing == "leading"
        self._leading_timesteps(num_inference_steps, device)

    def previous_timestep(self, timestep):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def _get_variance(self, t, predicted_variance=None, variance_type=None):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        variance = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        variance = torch.clamp(variance, min=1e-20)
        vt = variance_type or self.config.variance_type
        if vt == "fixed_small_log":
            variance = torch.exp(0.5 * torch.log(variance))
        elif vt != "fixed_small":
            raise NotImplementedError(vt)
        return variance
