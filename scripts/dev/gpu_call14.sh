set -x
for M in pythia-410m roberta; do
  python bench.py --model $M --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench14_sep_$M.json 2> gpurun_out/r02_bench14_sep_$M.err
  B200_FUSED_BIAS_GRAD=1 python bench.py --model $M --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench14_fused_$M.json 2> gpurun_out/r02_bench14_fused_$M.err
done
python bench.py --model pythia-2.8b --checkpointing --grad-acc 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench14_2p8b.json 2> gpurun_out/r02_bench14_2p8b.err
CMD="python bench.py --model pythia-2.8b --checkpointing --steps 1 --warmup 1 --grad-acc 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_2p8b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_pythia-2.8b.csv $CMD > gpurun_out/ncu_2p8b.log 2>&1
echo done
