# final 8-GPU pass at HEAD: the headline config with the replicated and with the sharded fp32 master, config 5 with the sharded master,
# config 4 (S 512) after the dropout / hash / delta changes
set -x
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --no-cpu-baseline --steps 2 --warmup 3"
$TR --master-port 29612 $B > gpurun_out/r02_final_n${N}_1b_zero1.json 2> gpurun_out/r02_final_n${N}_1b_zero1.err
$TR --master-port 29613 $B --shard-master > gpurun_out/r02_final_n${N}_1b_zero1_shardmaster.json 2> gpurun_out/r02_final_n${N}_1b_zero1_shardmaster.err
$TR --master-port 29614 $B --model pythia-2.8b --strategy zero1 --checkpointing --grad-acc 4 --shard-master > gpurun_out/r02_final_n${N}_2p8b_zero1_ckpt_shardmaster.json 2> gpurun_out/r02_final_n${N}_2p8b_zero1_ckpt_shardmaster.err
$TR --master-port 29615 $B --model roberta --seq-len 512 --strategy ddp > gpurun_out/r02_final_n${N}_roberta_s512.json 2> gpurun_out/r02_final_n${N}_roberta_s512.err
echo done
