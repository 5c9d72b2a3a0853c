# 8-GPU measurement pass (one box, NVLink): DP parity at 8 ranks, then the BASELINE.json configs 2-5 through bench.py presets.
set -x
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 scripts/dev/dp_check.py > gpurun_out/r02_dp_check_n$N.log 2>&1; echo "dp_check rc=$?" | tee -a gpurun_out/r02_dp_check_n$N.log
B="bench.py --gpus $N --no-cpu-baseline"
# config 3: Pythia-1b ZeRO-1, mbs 16, ga 16 (+ per-rank phases), then the batch-1024-preserving ga
$TR --master-port 29612 $B --steps 3 --warmup 3 --phases > gpurun_out/r02_bench_n${N}_1b_zero1.json 2> gpurun_out/r02_bench_n${N}_1b_zero1.err
$TR --master-port 29613 $B --steps 3 --warmup 3 --batch-preserving > gpurun_out/r02_bench_n${N}_1b_zero1_batch1024.json 2> gpurun_out/r02_bench_n${N}_1b_zero1_batch1024.err
# config 2: Pythia-410m DDP
$TR --master-port 29614 $B --model pythia-410m --strategy ddp --steps 3 --warmup 3 --phases > gpurun_out/r02_bench_n${N}_410m_ddp.json 2> gpurun_out/r02_bench_n${N}_410m_ddp.err
# config 4: RoBERTa-large, S 512 and S 128 (DDP: the reference's default for it is no sharding)
$TR --master-port 29615 $B --model roberta --seq-len 512 --strategy ddp --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}_roberta_s512.json 2> gpurun_out/r02_bench_n${N}_roberta_s512.err
$TR --master-port 29616 $B --model roberta --seq-len 128 --strategy ddp --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}_roberta_s128.json 2> gpurun_out/r02_bench_n${N}_roberta_s128.err
# config 5: Pythia-2.8b ZeRO-1 + activation checkpointing
$TR --master-port 29617 $B --model pythia-2.8b --strategy zero1 --checkpointing --grad-acc 4 --steps 2 --warmup 3 --phases > gpurun_out/r02_bench_n${N}_2p8b_zero1_ckpt.json 2> gpurun_out/r02_bench_n${N}_2p8b_zero1_ckpt.err
# f3: ZeRO-2 on the 1b config (gradient sharding: reduce-scatter every micro-batch)
$TR --master-port 29618 $B --strategy zero2 --grad-acc 4 --steps 2 --warmup 3 > gpurun_out/r02_bench_n${N}_1b_zero2.json 2> gpurun_out/r02_bench_n${N}_1b_zero2.err
echo done
