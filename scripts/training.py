"""Mirrors scripts/training.py of the reference (same entry points `get_model`, `get_dataset`, `get_data_collator`,
`get_optimizer_cls_and_kwargs`, `train`, `run`; same arguments --output-dir --model-type --training-arguments --data-path
--data-split) for the text families this build covers. The reference raises NotImplementedError for Pythia / RoBERTa in
`get_dataset` / `get_data_collator` (scripts/training.py:35-36,55-56); here those branches exist (text_data.py) and the HF
`Trainer(...).train()` of scripts/training.py:88-104 is the B200 step engine driven by the SAME training-arguments JSON that
`scripts/to_training_arguments.py` writes (max_steps, per_device_train_batch_size, gradient_accumulation_steps, lr schedule,
bf16 / fp16, deepspeed / fsdp sharding, max_grad_norm, + HF's logging_steps / save_steps / resume_from_checkpoint / seed).

    python scripts/training.py --output-dir out --model-type pythia-160m --training-arguments args.json --data-path corpus/ --data-split train
    torchrun --nnodes 1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/training.py ...     (one process per GPU; torchrunx in the reference)

Checkpoints follow HF Trainer's layout: <output_dir>/checkpoint-<global_step>/{pytorch_model.bin, optimizer*.pt, scheduler.pt,
trainer_state.json}; `resume_from_checkpoint: true` continues from the newest one with the same data order, dropout masks and
loss-scaler state (tests/test_training_script_gpu.py: interrupted + resumed == uninterrupted, bit for bit).
"""
import argparse
import json
import os
import re
import sys
import time
from pathlib import Path
from typing import Any

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from multimodal_llm_pretraining_b200.models import get_model_class  # noqa: E402
from multimodal_llm_pretraining_b200.text_data import EpochSampler, get_text_collator, get_text_dataset  # noqa: E402


def get_model(model_type: str):  # scripts/training.py:15-16
    return get_model_class(model_type).build_model(use_custom_kernels=True)


def get_dataset(model_type: str, data_path: Path, data_split: str):  # scripts/training.py:19-36
    return get_text_dataset(get_model_class(model_type), data_path, data_split)


def get_data_collator(model_type: str, model=None, seed: int = 0):  # scripts/training.py:39-56
    return get_text_collator(model_type, get_model_class(model_type), seed)


def get_optimizer_cls_and_kwargs(model_type: str, using_deepspeed: bool):  # scripts/training.py:59-70
    """The reference returns None under DeepSpeed (DeepSpeed then builds its own Adam from "auto" = TrainingArguments defaults,
    SURVEY App. C.3); this build always uses the model class's tuple so that every strategy optimises the same objective."""
    mc = get_model_class(model_type)
    return (mc.optimizer, mc.optimizer_kwargs)


def _sharding_of(training_arguments: dict) -> tuple[str, str]:
    ds = training_arguments.get("deepspeed") or None
    zero = str(ds.get("zero_optimization", {}).get("stage", 0)) if isinstance(ds, dict) else "0"
    fsdp = training_arguments.get("fsdp") or ""
    fsdp = fsdp[0] if isinstance(fsdp, (list, tuple)) and fsdp else (fsdp.split()[0] if isinstance(fsdp, str) and fsdp else "no_shard")
    return zero, fsdp


def latest_checkpoint(output_dir: Path) -> Path | None:
    cands = [(int(m.group(1)), p) for p in output_dir.glob("checkpoint-*") if (m := re.fullmatch(r"checkpoint-(\d+)", p.name)) and (p / "trainer_state.json").exists()]
    return max(cands)[1] if cands else None


def train(output_dir: str, model_type: str, training_arguments: dict[str, Any], data_path: Path, data_split: str) -> dict:  # scripts/training.py:73-104
    from multimodal_llm_pretraining_b200.benchmarking.utils import ManualTrainer

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    args = dict(training_arguments)
    seed = int(args.pop("seed", 42))
    logging_steps = int(args.pop("logging_steps", 500))
    save_steps = int(args.pop("save_steps", 500))
    resume = args.pop("resume_from_checkpoint", None)
    stop_after = args.pop("_stop_after_steps", None)  # test hook: simulate an interruption
    torch.manual_seed(seed)
    model = get_model(model_type)
    train_dataset = get_dataset(model_type, Path(data_path), data_split)
    collator = get_data_collator(model_type, model, seed)
    zero, fsdp = _sharding_of(args)
    args.setdefault("lr_scheduler_kwargs", {})
    args.setdefault("warmup_steps", 0)
    args.setdefault("gradient_checkpointing", False)
    args.setdefault("max_grad_norm", get_model_class(model_type).max_grad_norm)
    trainer = ManualTrainer(model=model, args=args, train_dataset=train_dataset,
                            optimizer_cls_and_kwargs=get_optimizer_cls_and_kwargs(model_type, args.get("deepspeed") is not None),
                            scheduler_type=args.get("lr_scheduler_type", "linear"), zero_stage=zero, fsdp_sharding=fsdp, seed=seed)
    eng = trainer.engine
    out = Path(output_dir)
    log_history: list[dict] = []
    if resume:
        ck = Path(resume) if isinstance(resume, str) and resume not in ("True", "true") else latest_checkpoint(out)
        if ck is not None:
            eng.load_checkpoint(ck)
            st = json.loads((ck / "trainer_state.json").read_text())
            log_history = st.get("log_history", [])
            if rank == 0:
                print(f"resumed from {ck} at global step {eng.micro // eng.ga}", flush=True)
    mbs, ga, max_steps = args["per_device_train_batch_size"], args["gradient_accumulation_steps"], args["max_steps"]
    sampler = EpochSampler(len(train_dataset), mbs, world, rank, seed)
    step = eng.micro // ga
    t0, tokens_seen, window = time.perf_counter(), 0, []
    while step < max_steps:
        for _ in range(ga):
            t = eng.micro
            batch = collator([train_dataset[int(i)] for i in sampler.rows(t)], step=t * world + rank)
            batch = {k: v.pin_memory() for k, v in batch.items()}
            window.append(trainer.manual_training_step(trainer.model, batch))
            tokens_seen += batch["input_ids"].numel() * world
        took = trainer.manual_optimization_step(trainer.model)
        step += 1
        if step % logging_steps == 0 or step == max_steps:
            loss = float(torch.stack(window).mean())
            window = []
            if world > 1:
                lt = torch.tensor([loss], device="cuda")
                dist.all_reduce(lt, op=dist.ReduceOp.AVG)
                loss = float(lt)
            rec = {"step": step, "loss": loss, "learning_rate": trainer.optimizer.param_groups[0]["lr"],
                   "grad_norm": float(eng.last_grad_norm) if eng.last_grad_norm is not None else None,
                   "tokens_per_s": tokens_seen / (time.perf_counter() - t0), "optimizer_step_taken": bool(took),
                   "loss_scale": float(eng.loss_scaler.scale) if eng.loss_scaler is not None else None}
            log_history.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
        if step % save_steps == 0 or step == max_steps or (stop_after is not None and step == stop_after):
            ck = out / f"checkpoint-{step}"
            eng.save_checkpoint(ck)
            if rank == 0:
                st = json.loads((ck / "trainer_state.json").read_text())
                st["log_history"] = log_history
                (ck / "trainer_state.json").write_text(json.dumps(st))
        if stop_after is not None and step == stop_after:
            break
    return {"global_step": step, "log_history": log_history, "trainer": trainer}


def run(output_dir: str, model_type: str, training_arguments: Path, data_path: Path, data_split: str):  # scripts/training.py:107-125
    res = train(output_dir=output_dir, model_type=model_type, training_arguments=json.load(open(training_arguments)),
                data_path=data_path, data_split=data_split)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--output-dir", required=True)
    ap.add_argument("--model-type", required=True)
    ap.add_argument("--training-arguments", required=True, type=Path, help="JSON written by scripts/to_training_arguments.py")
    ap.add_argument("--data-path", required=True, type=Path)
    ap.add_argument("--data-split", default="train")
    a = ap.parse_args()
    run(a.output_dir, a.model_type, a.training_arguments, a.data_path, a.data_split)
