#!/usr/bin/env python
"""bench.py — the hot path's headline metric: train tokens/s (+ MFU, training-days estimate) of the data-parallel
pretraining step (forward + backward + optimizer) of Pythia-1b bf16, micro-batch 16 x 2049 tokens, grad-acc 16
(README's benchmark config; BASELINE.json configs[2]), device-timed, on N B200s of one box.

    python bench.py --gpus N --steps K --warmup W              # this repo (libb200pt kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path on the host cores

For N > 1 launch with torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
A "step" = grad-acc micro-batches (fwd+bwd) + clip + fused Adam + scheduler + zero_grad, i.e.
manual_training_step x ga + manual_optimization_step of src/benchmarking/utils.py:61-80.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="pythia-1b")
    ap.add_argument("--mbs", type=int, default=None, help="micro-batch size (default 16 for Pythia, 64 for RoBERTa)")
    ap.add_argument("--grad-acc", type=int, default=16)
    ap.add_argument("--seq-len", type=int, default=None, help="RoBERTa only: 512 (default) or 128")
    ap.add_argument("--strategy", default=None, choices=[None, "none", "ddp", "zero1", "zero2", "zero3"])
    ap.add_argument("--checkpointing", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"],
                    help="bf16 (every BASELINE.json config) or fp16 + dynamic loss scaling (the reference's own precision for Pythia != 1b / RoBERTa)")
    ap.add_argument("--batch-preserving", action="store_true",
                    help="grad-acc = 1024 / (gpus * mbs): the reference's rule W * mbs * ga = batch_size (src/models/__init__.py:99-102)")
    ap.add_argument("--shard-master", action="store_true",
                    help="ZeRO-1/2 only: shard the fp32 master too (each rank keeps the fp32 weights of the slices it owns; parameters become "
                         "views of the 16-bit copy), i.e. DeepSpeed's 12 B/param/W of optimizer state per rank")
    ap.add_argument("--phases", action="store_true",
                    help="after the timed region, one more instrumented step: per-rank micro-batch and optimizer-phase device times in the JSON line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-tokens", type=int, default=1024,
                    help="predicted tokens of the bounded CPU sample (one sequence): ~10 s per step on 16 cores, two steps")
    ap.add_argument("--no-config1", action="store_true", help="skip the BASELINE.json configs[0] CPU figures (Pythia-160m, OMP=1 and all cores)")
    ap.add_argument("--config1-tokens", default="256,2048", help="predicted tokens of the bounded config-1 samples: OMP=1, all cores")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_step_factory(model_name: str, sample_tokens: int):
    """The reference's own CPU path for this step, on a bounded sample: the model src/models/pythia.py:15-22 builds
    (transformers.GPTNeoXForCausalLM, fp32) + torch.optim.Adam (src/models/pythia.py:43-67) + clip (utils.py:66-70),
    plain loop = Trainer.training_step + optimizer.step arithmetic. Falls back to the oracle port if transformers is
    not importable."""
    from multimodal_llm_pretraining_b200.models.configs import pythia_config_dict

    cfg = pythia_config_dict(model_name)
    torch.manual_seed(0)
    ids = torch.randint(0, cfg["vocab_size"], (1, sample_tokens + 1))
    try:
        from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

        model = GPTNeoXForCausalLM(GPTNeoXConfig(**cfg, attn_implementation="sdpa")).float().train()
        opt = torch.optim.Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.95), eps=1e-8)

        def step():
            loss = model(input_ids=ids, labels=ids).loss
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            model.zero_grad()
            return float(loss.detach())

        import transformers

        return step, "reference", f"transformers {transformers.__version__} GPTNeoXForCausalLM fp32 + torch.optim.Adam"
    except Exception:  # pragma: no cover - transformers is in the image
        from oracle import neox_oracle as O

        P = {}
        from multimodal_llm_pretraining_b200.modeling_gpt_neox import neox_param_shapes
        from types import SimpleNamespace

        for n, s in neox_param_shapes(SimpleNamespace(**cfg)):
            P[n] = torch.ones(s) if "norm.weight" in n else (torch.zeros(s) if n.endswith("bias") else torch.randn(s) * 0.02)
        state = {}

        def step():
            loss, grads = O.neox_loss_and_grads(P, ids, ids, cfg)
            _, coef = O.clip_coef(grads, 1.0)
            O.adam_step(P, {k: g * coef for k, g in grads.items()}, state, lr=3e-4)
            return float(loss)

        return step, "port", "oracle/neox_oracle.py fp32"


def _all_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_baseline(model_name: str, sample_tokens: int, steps: int, warmup: int, threads: int | None = None) -> dict:
    # default: all the host cores this process may use — torchrun exports OMP_NUM_THREADS=1 to every rank, which would otherwise
    # make the reference arm at N > 1 a single-threaded (16x slower) run; threads=1 is the reference-faithful setting (.env:4)
    try:
        torch.set_num_threads(threads or _all_cores())
    except RuntimeError:
        pass
    step, kind, what = cpu_reference_step_factory(model_name, sample_tokens)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_tokens / dt, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{model_name} 1 sequence x {sample_tokens + 1} tokens ({sample_tokens} predicted) fwd+bwd+clip+Adam, {what}, "
                      f"{steps} timed step(s) after {warmup} warm-up, {dt:.2f} s/step, host cpu_count={os.cpu_count()}",
            "s_per_step": dt}


def run_config1_cpu(tokens: str = "256,2048", steps: int = 2, warmup: int = 1) -> dict:
    """BASELINE.json configs[0], the reference's own CPU-runnable case: Pythia-160m causal-LM step, micro-batch 4 x 2049 tokens,
    fp32 (SURVEY §8d "CPU baseline"), at the reference-faithful OMP_NUM_THREADS=1 (/root/reference/.env:4) and at all cores.
    Bounded: a full 4 x 2049 micro-batch takes minutes on one core, so each setting times one sequence of the micro-batch,
    truncated (the per-token cost of a causal LM GROWS with context, so the truncated sample flatters the CPU)."""
    out = {"workload": "pythia-160m, micro-batch 4 x 2049 tokens, fp32, HF GPTNeoXForCausalLM + torch.optim.Adam + clip (BASELINE.json configs[0])"}
    t1, tall = (int(x) for x in tokens.split(","))
    for key, threads, toks in (("omp_num_threads_1", 1, t1), ("all_cores", None, tall)):
        r = run_cpu_baseline("pythia-160m", toks, steps, warmup, threads=threads)
        out[key] = {"tokens_per_s": r["value"], "cores": r["cores"], "s_per_step": r["s_per_step"],
                    "sample": f"1 of the 4 sequences, {toks} predicted tokens, {steps} timed steps after {warmup} warm-up"}
    return out


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if a.mbs is None:
        a.mbs = 16
    cb = run_cpu_baseline(a.model, a.cpu_sample_tokens, a.steps, a.warmup)
    config1 = None if a.no_config1 else run_config1_cpu(a.config1_tokens)
    line = {
        "impl": "reference", "metric": "train_tokens_per_s", "value": cb["value"], "unit": "tokens/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{a.model} pretraining step (fwd+bwd+Adam), mbs {a.mbs} x 2049 tokens, grad-acc {a.grad_acc}; "
                               f"CPU arm runs a bounded sample of it: {cb['sample']}"},
        "cpu_baseline": {**{k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}, "config1": config1},
        "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def main_b200(a):
    import torch.distributed as dist

    from multimodal_llm_pretraining_b200 import kernels as K
    from multimodal_llm_pretraining_b200.engine import TrainEngine
    from multimodal_llm_pretraining_b200.models import get_model_class
    from multimodal_llm_pretraining_b200.models.configs import neox_train_flops_per_sequence, roberta_train_flops_per_sequence
    from multimodal_llm_pretraining_b200.optim import fused_optimizer_class, get_scheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on STDOUT; rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    strategy = a.strategy or ("none" if world == 1 else "zero1")

    mc = get_model_class(a.model)
    cfg = mc.config_dict()
    is_roberta = a.model == "roberta"
    if a.mbs is None:
        a.mbs = (256 if a.seq_len == 128 else 64) if is_roberta else 16  # 32 768 tokens per micro-batch either way
    if a.batch_preserving:
        a.grad_acc = max(1, mc.batch_size // (world * a.mbs))
    if is_roberta:
        S_in = S_pred = a.seq_len or mc.sequence_length  # masked LM: every position is predicted (labels = input_ids)
    else:
        S_in = mc.sequence_length  # 2049 tokens in, 2048 predicted
        S_pred = S_in - 1
    torch.manual_seed(0)
    model = mc.build_model(use_custom_kernels=True)
    if a.precision == "fp16":
        model.set_compute_dtype(torch.float16)
    model = model.to(dev).train()
    if a.checkpointing:
        model.gradient_checkpointing_enable()
    okw = dict(mc.optimizer_kwargs)
    okw["weight_decay"] = 0.0  # HF Trainer's param groups override the kwarg with TrainingArguments.weight_decay = 0
    opt = fused_optimizer_class(mc.optimizer)(model.parameters(), **okw)
    skw = dict(mc.scheduler_kwargs)
    warm = skw.pop("num_warmup_steps", 0)
    sched = get_scheduler(mc.scheduler_type, opt, warm, mc.training_steps, skw)
    eng = TrainEngine(model, opt, sched, max_grad_norm=mc.max_grad_norm, gradient_accumulation_steps=a.grad_acc, strategy=strategy,
                      shard_master=a.shard_master)

    ga, mbs = a.grad_acc, a.mbs
    total_steps = a.warmup + a.steps
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = ga * 2  # pinned host micro-batches, cycled
    host = [torch.randint(0, mc.vocab_size, (mbs, S_in), generator=g).pin_memory() for _ in range(n_host)]
    resident = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, from_host: bool):
        last = None
        for s in range(n):
            losses = []
            for m in range(ga):
                i = (s * ga + m) % n_host
                if from_host:
                    ids = host[i].to(dev, non_blocking=True)
                else:
                    ids = resident[i]
                losses.append(eng.manual_training_step({"input_ids": ids, "labels": ids}))
            eng.manual_optimization_step()
            if from_host:
                last = float(torch.stack(losses).mean())  # D2H read of the step's result
        return last

    def timed(n, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = K.LAUNCHES
        e0.record()
        last = run_steps(n, from_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), K.LAUNCHES - l0, last

    # warm-up (untimed)
    run_steps(a.warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, _ = timed(a.steps, False)
    ms_e2e, _, last_loss = timed(a.steps, True)
    clocks = sampler.stop() if rank == 0 else None

    # roofline pass: CUDA events around every tcgen05 GEMM launch of one more step (same stream, same workload)
    K.GEMM_PROFILE = []
    run_steps(1, False)
    torch.cuda.synchronize()
    prof, K.GEMM_PROFILE = K.GEMM_PROFILE, None
    gemm_flops = sum(p[0] for p in prof)
    gemm_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(1, False)
    e1.record()
    torch.cuda.synchronize()
    step_ms_plain = e0.elapsed_time(e1)

    tokens_per_step = world * ga * mbs * S_pred
    ms_per_step = ms_total / a.steps
    value = tokens_per_step / (ms_per_step / 1e3)
    e2e_value = tokens_per_step / (ms_e2e / a.steps / 1e3)
    f_tok = (roberta_train_flops_per_sequence(cfg, S_in) if is_roberta else neox_train_flops_per_sequence(cfg, S_in)) / S_pred
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    per_gpu = value / world
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    # opt-in: per-rank device times of one more instrumented step (micro-batches, boundary micro-batch, optimizer phases)
    phases = None
    if a.phases:
        eng.profile_phases = True
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(ga + 2)]
        evs[0].record()
        for m in range(ga):
            eng.manual_training_step({"input_ids": resident[m % n_host], "labels": resident[m % n_host]})
            evs[m + 1].record()
        eng.manual_optimization_step()
        evs[ga + 1].record()
        torch.cuda.synchronize()
        eng.profile_phases = False
        micro = [evs[i].elapsed_time(evs[i + 1]) for i in range(ga)]
        mine = {"rank": rank, "micro_ms_plain_mean": sum(micro[:-1]) / max(1, ga - 1) if ga > 1 else micro[0], "micro_ms_boundary": micro[-1],
                "micro_ms_min": min(micro), "micro_ms_max": max(micro), "optim_ms": evs[ga].elapsed_time(evs[ga + 1]),
                "optim_phases_ms": eng.last_phase_ms, "step_ms": evs[0].elapsed_time(evs[ga + 1])}
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, mine)
        else:
            allp = [mine]
        phases = {"per_rank": allp, "slowest_rank_step_ms": max(p_["step_ms"] for p_ in allp), "fastest_rank_step_ms": min(p_["step_ms"] for p_ in allp),
                  "note": "device times (CUDA events) of one step after the timed region; the step time of the job is the slowest rank's"}

    if rank == 0:
        cpu = None
        if not a.no_cpu_baseline and world == 1 and not is_roberta:  # the CPU arm is the Pythia reference path
            cpu = run_cpu_baseline(a.model, a.cpu_sample_tokens, 2, 1)  # ~12 s of CPU work in total
            cpu["config1"] = None if a.no_config1 else run_config1_cpu(a.config1_tokens)  # + ~35 s: Pythia-160m at OMP=1 and at all cores
        line = {
            "metric": "train_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {
                "workload": f"{a.model} pretraining step: {ga} micro-batches x ({mbs} x {S_in} tokens, {S_pred} predicted) fwd+bwd "
                            f"+ clip + fused Adam (random-init weights, uniform random tokens)",
                "model": a.model, "micro_batch": mbs, "grad_acc": ga, "global_batch_sequences": world * ga * mbs,
                "seq_len": S_in, "parallelism": f"{strategy}x{world}", "activation_checkpointing": bool(a.checkpointing), "precision": a.precision,
                "fp32_master": "sharded" if eng.shard_master else "replicated",
                **({"note": "hidden dropout 0.1 and attention-probability dropout 0.1 applied (roberta-large config)"} if is_roberta else {}),
                "l2": "working set per step (>= 2 GB of weights, > 30 GB of activations) far exceeds the 126 MB L2; no explicit flush",
            },
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": ga * mbs * S_in * 8, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / a.steps, "last_loss": last_loss},
            "gpu_launches": launches,
            "clocks": clocks,
            "memory": {"rank0_max_allocated_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
                       "rank0_parameter_state_gb": sum(t.numel() * t.element_size() for t in (model.flat.master, model.flat.grad, model.flat.shadow, model.flat.small,
                                                                                               opt._m, opt._v, opt._p32, getattr(eng, "_gshard", None), getattr(eng, "_w16", None)) if t is not None) / 1e9},
            "mfu": {
                "flops_per_token": f_tok, "tflops_per_gpu": per_gpu * f_tok / 1e12,
                "vs_datasheet_2250": per_gpu * f_tok / 1e12 / 2250.0, "vs_measured_sustained": per_gpu * f_tok / 1e12 / peak_sust,
                "definition": "F(S)=6*S*W_lin+12*L*h*S^2 per sequence (= FlopCounterMode, src/benchmarking/flops.py), per predicted token",
            },
            "training_days": mc.training_steps * (mc.batch_size / (world * ga * mbs)) * (ms_per_step / 1e3) / 86400.0,
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peak_sust, "unit": "TFLOP/s",
                "frac": (achieved / peak_sust) if achieved else None,
                # the aggregate spans every GEMM variant of the step, so there is no single per-launch DRAM figure for it; the
                # per-variant dram__bytes of this build are in profiles/ (ncu --set full), not re-stated here from a file
                "traffic": None,
                "kernel": "gemm_kernel (tcgen05 bf16 GEMM, all fwd/dgrad/wgrad launches of one step)",
                "how": f"sum of 2*M*N*K over {len(prof)} launches / sum of their CUDA-event durations in a separate instrumented step; "
                       f"GEMM share of step {gemm_ms / step_ms_plain:.3f}; peak = {peak_src}",
            },
        }
        if phases is not None:
            line["phases"] = phases
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "config1")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(main_reference(args) if args.impl == "reference" else main_b200(args))
