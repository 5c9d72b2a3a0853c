"""Mirrors scripts/benchmark.py of the reference (same arguments: --num-nodes --gpus-per-node --gpu-type --model --methods)
on the B200 step engine: for every method combination of the reference's search space that is in this build's scope
(no sharding -> DDP, zero_1, zero_2, fsdp_shard_grad_op, each with/without activation checkpointing) it runs the reference's three steps
(experiments/training_time_empirical.py:43-138): find the largest power-of-two micro-batch -> benchmark the
accumulate / optimize times (device-timed) -> training days = training_steps * step_time / 86400.

    python scripts/benchmark.py --num-nodes 1 --gpus-per-node 1 --gpu-type b200 --model pythia-1b --methods free-lunch
    torchrun --nnodes 1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/benchmark.py --num-nodes 1 --gpus-per-node 8 ...

The reference launches its workers with torchrunx (experiments/utils/distribute.py:37-61); here the same role — one
process per GPU, rank 0 reports — is played by torchrun.
"""
import argparse
import json
import math
import os
import sys
from itertools import product
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from multimodal_llm_pretraining_b200.benchmarking.max_batch_size import find_max_mbs_pow2  # noqa: E402
from multimodal_llm_pretraining_b200.benchmarking.step_time import compute_training_days, estimate_step_time  # noqa: E402
from multimodal_llm_pretraining_b200.config import TrainingConfig  # noqa: E402
from multimodal_llm_pretraining_b200.gpus import ampere_or_newer_gpu  # noqa: E402
from multimodal_llm_pretraining_b200.models import get_model_class  # noqa: E402


def validate_arguments(num_nodes: int, gpus_per_node: int, gpu_type: str, model: str):  # scripts/benchmark.py:13-31
    model_class = get_model_class(model)
    num_gpus = num_nodes * gpus_per_node
    assert model_class.batch_size % num_gpus == 0, f"model batch size ({model_class.batch_size}) should be evenly divisible by total GPUs ({num_gpus})"
    assert math.log2(model_class.batch_size // num_gpus).is_integer(), f"batch size per gpu ({model_class.batch_size // num_gpus}) should be power of 2"
    if model_class.mixed_precision == "bf16":
        assert ampere_or_newer_gpu(gpu_type), "GPU must be ampere or newer to use mixed precision with bf16"


def search_space(methods: str):  # scripts/benchmark.py:45-64
    free_lunch, ckpt, sharding, offloading = [False], [False], [""], [False]
    if methods == "free-lunch":
        free_lunch = [True]
    elif methods == "all":
        free_lunch, ckpt = [True], [False, True]
        sharding = ["", "zero_1", "zero_2", "zero_3", "fsdp_shard_grad_op", "fsdp_full_shard"]
        offloading = [False, True]
    return list(product(free_lunch, ckpt, sharding, offloading))


RESULTS_FILE = Path(__file__).resolve().parent.parent / "results" / "training_time_empirical.jsonl"
PRECISION = "bf16"  # "bf16" (every BASELINE.json config) or "reference" (the model class's mixed_precision: fp16 for Pythia != 1b, RoBERTa)


def build_benchmarking_trainer(config: TrainingConfig, num_samples: int = 4096):  # experiments/training_time_empirical.py:17-40
    over = dict(bf16=True, fp16=False) if PRECISION == "bf16" else {}
    training_class = config.training_class(num_training_steps=1, micro_batch_size=1, gradient_accumulation_steps=1, **over)
    model_class = config.model_class()
    model = model_class.build_model(use_custom_kernels=True)  # the B200 module (the reference passes config.free_lunch)
    dataset = model_class.load_dummy_dataset(num_samples=num_samples, seed=0)
    return training_class.build_trainer(model, dataset)


def run_one(config: TrainingConfig, num_benchmarking_steps: int = 3):
    model_class = config.model_class()
    num_gpus = config.num_nodes * config.gpus_per_node
    target_mbs = model_class.batch_size // num_gpus
    trainer = build_benchmarking_trainer(config)
    mbs = find_max_mbs_pow2(trainer, limit=target_mbs)  # experiments/training_time_empirical.py:43-57
    if mbs == 0:
        return None
    step_time = estimate_step_time(trainer, mbs, target_mbs, num_benchmarking_steps)  # :66-130
    days = compute_training_days(model_class.training_steps, step_time)  # :133-138
    del trainer
    torch.cuda.empty_cache()
    return dict(micro_batch_size=mbs, gradient_accumulation_steps=target_mbs // mbs, step_time_s=step_time, training_days=days)


def run_benchmark(num_nodes: int, gpus_per_node: int, gpu_type: str, model: str, methods: str = "all"):
    validate_arguments(num_nodes, gpus_per_node, gpu_type, model)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    assert world == num_nodes * gpus_per_node, f"launch one process per GPU (WORLD_SIZE={world}, expected {num_nodes * gpus_per_node})"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    results = []
    for free_lunch, ckpt, sharding, offloading in search_space(methods):
        config = TrainingConfig(num_nodes, gpus_per_node, gpu_type, model, free_lunch, ckpt, sharding, offloading)
        tc = config.training_class()
        if not tc.is_valid() or (sharding != "" and world == 1):  # experiments/training_time_empirical.py:161-186
            continue
        if not tc.runs_on_b200_engine():
            if rank == 0:
                print(f"[skip] {config}: outside this build's scope (DDP / ZeRO-1 / ZeRO-2 / FSDP shard_grad_op)", flush=True)
            continue
        r = run_one(config)
        if rank == 0:
            # one row per configuration, the columns scripts/print_optimal_config.py selects (the reference keeps them in its Tango
            # step cache, experiments/training_time_empirical_sweep.py); appended to a JSON-lines file here
            row = dict(num_nodes=num_nodes, gpus_per_node=gpus_per_node, gpu_type=gpu_type, model=model, free_lunch=free_lunch,
                       activation_checkpointing=ckpt, sharding=sharding, offloading=offloading, precision=PRECISION,
                       **(r or {"micro_batch_size": 0, "training_days": None}))
            results.append(row)
            print(json.dumps(row), flush=True)
            RESULTS_FILE.parent.mkdir(parents=True, exist_ok=True)
            with open(RESULTS_FILE, "a") as fh:
                fh.write(json.dumps(row) + "\n")
    if rank == 0 and results:
        best = min((r for r in results if r.get("training_days")), key=lambda r: r["training_days"], default=None)
        print("optimal:", json.dumps(best))
    if world > 1:
        dist.destroy_process_group()
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-nodes", type=int, required=True)
    ap.add_argument("--gpus-per-node", type=int, required=True)
    ap.add_argument("--gpu-type", required=True)
    ap.add_argument("--model", required=True)
    ap.add_argument("--methods", default="all", choices=["naive", "free-lunch", "all"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "reference"])
    ap.add_argument("--results-file", type=Path, default=RESULTS_FILE)
    a = ap.parse_args()
    PRECISION, RESULTS_FILE = a.precision, a.results_file
    try:
        run_benchmark(a.num_nodes, a.gpus_per_node, a.gpu_type, a.model, a.methods)
    except KeyboardInterrupt:
        sys.exit(128 + 2)
