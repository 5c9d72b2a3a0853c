"""GPU registry — mirrors src/gpus.py:3-11 of the reference and adds the B200 this build targets."""
from typing import Literal

GpuT = Literal["geforce3090", "v100", "a6000", "a40", "l40", "a100", "h100", "b200"]

# dense bf16 / tf32 peaks in TFLOP/s, the table of experiments/training_time_analytic.py:24-47 plus a b200 row
PEAK_TFLOPS = {
    "h100": {"bf16": 756.0, "tf32": 378.0},
    "a100": {"bf16": 312.0, "tf32": 156.0},
    "a6000": {"bf16": 154.8, "tf32": 77.4},
    "geforce3090": {"bf16": 71.0, "tf32": 35.6},
    "b200": {"bf16": 2250.0, "tf32": 1125.0},
}


def ampere_or_newer_gpu(gpu_type: GpuT) -> bool:
    match gpu_type:
        case "geforce3090" | "a6000" | "a40" | "l40" | "a100" | "h100" | "b200":
            return True
        case "v100":
            return False
    raise ValueError(f"unknown gpu type {gpu_type!r}")
