"""count_flops_per_example — mirrors src/benchmarking/flops.py:9-37 (the repo's FLOP metric → MFU / analytic days).

The reference runs torch FlopCounterMode over one eager fwd+bwd on a GPU; for the in-scope GPT-NeoX family that count
has the closed form F(S) = 6 S W_lin + 12 L h S^2 (verified to the last digit there), which needs no
device. `count_flops_per_example_measured` is the reference's procedure verbatim for cross-checking."""
from __future__ import annotations

from ..models import BaseModelClass
from ..models.configs import neox_train_flops_per_sequence


def count_flops_per_example(model_class: BaseModelClass) -> float:
    cfg = model_class.config_dict()
    if "rotary_pct" in cfg:
        return float(neox_train_flops_per_sequence(cfg, model_class.sequence_length))
    return count_flops_per_example_measured(model_class)


def count_flops_per_example_measured(model_class: BaseModelClass, device: str = "meta") -> float:
    import torch
    from torch.utils.flop_counter import FlopCounterMode

    with torch.device(device):
        model = model_class.build_model(use_custom_kernels=False)
        S = model_class.sequence_length
        ids = torch.zeros(1, S, dtype=torch.long)
    with FlopCounterMode(display=False) as fc:
        model(input_ids=ids, labels=ids).get("loss").backward()
    return float(fc.get_total_flops())


def estimate_training_days_from_flops(num_nodes: int, gpus_per_node: int, gpu_type: str, model_class: BaseModelClass,
                                      training_flops: float | None = None) -> float:
    """experiments/training_time_analytic.py:14-53 with the b200 row added (gpus.PEAK_TFLOPS): training days if every GPU ran
    at its datasheet dense tensor peak (bf16 when the model trains in mixed precision, tf32 otherwise).
    `training_flops` defaults to flops-per-example x batch size x training steps (experiments/count_flops.py)."""
    from ..gpus import PEAK_TFLOPS

    if gpu_type not in PEAK_TFLOPS:
        raise NotImplementedError(gpu_type)
    peak = PEAK_TFLOPS[gpu_type]["bf16" if model_class.mixed_precision is not None else "tf32"]
    if training_flops is None:
        training_flops = count_flops_per_example(model_class) * model_class.batch_size * model_class.training_steps
    flops_per_day = num_nodes * gpus_per_node * peak * 1e12 * 86400.0
    return training_flops / flops_per_day
