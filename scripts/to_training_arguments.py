"""Mirrors scripts/to_training_arguments.py of the reference: writes the HF TrainingArguments dict for a config.

    python scripts/to_training_arguments.py --output args.json --micro-batch-size 16 --gradient-accumulation-steps 16 \
        --num-nodes 1 --gpus-per-node 4 --gpu-type a100 --model pythia-1b --free-lunch --sharding zero_1
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from multimodal_llm_pretraining_b200.config import TrainingConfig  # noqa: E402


def save_arguments_to_file(output: Path, micro_batch_size: int, gradient_accumulation_steps: int, config: TrainingConfig) -> None:
    training_class = config.training_class(micro_batch_size=micro_batch_size, gradient_accumulation_steps=gradient_accumulation_steps)
    training_arguments = training_class._to_huggingface_args_dict()
    output.parent.mkdir(parents=True, exist_ok=True)
    json.dump(training_arguments, open(output, "w"))


def _cli():
    try:
        import tyro

        return tyro.cli(save_arguments_to_file, config=[tyro.conf.OmitArgPrefixes])
    except ImportError:
        ap = argparse.ArgumentParser()
        ap.add_argument("--output", type=Path, required=True)
        ap.add_argument("--micro-batch-size", type=int, required=True)
        ap.add_argument("--gradient-accumulation-steps", type=int, required=True)
        ap.add_argument("--num-nodes", type=int, required=True)
        ap.add_argument("--gpus-per-node", type=int, required=True)
        ap.add_argument("--gpu-type", required=True)
        ap.add_argument("--model", required=True)
        ap.add_argument("--free-lunch", action="store_true")
        ap.add_argument("--activation-checkpointing", action="store_true")
        ap.add_argument("--sharding", default="")
        ap.add_argument("--offloading", action="store_true")
        a = ap.parse_args()
        cfg = TrainingConfig(a.num_nodes, a.gpus_per_node, a.gpu_type, a.model, a.free_lunch, a.activation_checkpointing, a.sharding, a.offloading)
        return save_arguments_to_file(a.output, a.micro_batch_size, a.gradient_accumulation_steps, cfg)


if __name__ == "__main__":
    _cli()
