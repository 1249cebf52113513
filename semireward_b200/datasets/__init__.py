"""Input side of the hot path: the reference's per-sample PIL transforms as one device kernel."""
from .gpu_augment import AugDecision, DeviceImagePipeline, DeviceSSLLoader, draw_strong, draw_weak  # noqa: F401
