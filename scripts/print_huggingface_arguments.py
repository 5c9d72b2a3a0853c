"""Mirrors scripts/print_huggingface_arguments.py of the reference: pretty-prints the HF TrainingArguments dict for a config.

    python scripts/print_huggingface_arguments.py --micro-batch-size 16 --gradient-accumulation-steps 16 \
        --num-nodes 1 --gpus-per-node 4 --gpu-type a100 --model pythia-1b --free-lunch --sharding zero_1
"""
import argparse
import sys
from pathlib import Path
from pprint import pprint

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from multimodal_llm_pretraining_b200.config import TrainingConfig  # noqa: E402


def get_arguments(micro_batch_size: int, gradient_accumulation_steps: int, config: TrainingConfig) -> dict:
    return config.training_class(micro_batch_size=micro_batch_size,
                                 gradient_accumulation_steps=gradient_accumulation_steps)._to_huggingface_args_dict()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
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
    pprint(get_arguments(a.micro_batch_size, a.gradient_accumulation_steps, cfg))
