"""torchrun --nproc-per-node W scripts/dev/phase_bench.py [--model pythia-1b] [--no-overlap]
Device-timed phase breakdown of the data-parallel step: plain micro-batches, the boundary micro-batch (the one whose
backward overlaps the bucket collectives), and the pieces of manual_optimization_step. Max over ranks per phase."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402
from multimodal_llm_pretraining_b200.engine import TrainEngine  # noqa: E402
from multimodal_llm_pretraining_b200.models import get_model_class  # noqa: E402
from multimodal_llm_pretraining_b200.optim import get_scheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="pythia-1b")
    ap.add_argument("--mbs", type=int, default=16)
    ap.add_argument("--grad-acc", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--strategy", default=None)
    ap.add_argument("--no-overlap", action="store_true")
    a = ap.parse_args()
    rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    strategy = a.strategy or ("none" if world == 1 else "zero1")
    mc = get_model_class(a.model)
    torch.manual_seed(0)
    model = mc.build_model(use_custom_kernels=True).to(dev).train()
    okw = dict(mc.optimizer_kwargs)
    okw["weight_decay"] = 0.0
    opt = mc.optimizer(model.parameters(), **okw)
    skw = dict(mc.scheduler_kwargs)
    warm = skw.pop("num_warmup_steps", 0)
    sched = get_scheduler(mc.scheduler_type, opt, warm, mc.training_steps, skw)
    eng = TrainEngine(model, opt, sched, max_grad_norm=mc.max_grad_norm, gradient_accumulation_steps=a.grad_acc, strategy=strategy,
                      profile_phases=True, **({"overlap": False} if a.no_overlap else {}))
    S = mc.sequence_length
    g = torch.Generator().manual_seed(7 + rank)
    batches = [torch.randint(0, mc.vocab_size, (a.mbs, S), generator=g).to(dev) for _ in range(a.grad_acc)]

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    rows = []
    for s in range(a.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        marks = [ev()]
        for m in range(a.grad_acc):
            eng.manual_training_step({"input_ids": batches[m], "labels": batches[m]})
            marks.append(ev())
        eng.manual_optimization_step()
        marks.append(ev())
        torch.cuda.synchronize()
        d = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
        if s > 0:
            rows.append(d)
    t = torch.tensor(rows, dtype=torch.float64, device=dev).mean(0)
    per_rank = None
    if world > 1:
        mine = torch.stack([t[:-2].mean(), t[-2], t[-1]])
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 2) for v in a] for a in allr]  # [plain micro, boundary micro, optim] per rank
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.tolist()
    if rank == 0:
        print(json.dumps({"world": world, "strategy": strategy, "overlap": not a.no_overlap, "model": a.model,
                          "micro_ms_plain_mean": sum(t[:-2]) / max(1, len(t) - 2), "micro_ms_each": t[:-1],
                          "micro_ms_boundary": t[-2], "optim_ms": t[-1], "optim_phases_ms": getattr(eng, "last_phase_ms", None),
                          "per_rank_plain_boundary_optim_ms": per_rank}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
