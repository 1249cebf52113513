"""Input side of the hot path: the reference's per-sample PIL transforms as one device kernel."""
from .gpu_augment import (SAMPLE_DTYPE, AugDecision, DeviceImagePipeline, DeviceSSLLoader, draw_records, draw_strong, draw_weak,  # noqa: F401
                          pack_arrays, records_from_decisions)
