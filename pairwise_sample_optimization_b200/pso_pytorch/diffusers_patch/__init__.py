from .turbo_inference_with_logprob import turbo_step_with_logprob
from .distilled_inference_with_logprob import distilled_step_with_logprob, _get_x0_from_noise
from .sdxl_turbo_with_logprob import sdxl_turbo_pipeline_with_logprob
from .sdxl_dmd_with_logprob import sdxl_dmd_pipeline_with_logprob

__all__ = [
    "turbo_step_with_logprob",
    "distilled_step_with_logprob",
    "_get_x0_from_noise",
    "sdxl_turbo_pipeline_with_logprob",
    "sdxl_dmd_pipeline_with_logprob",
]
