# ncu launch lists of one short bench run per BASELINE model (shares, not absolutes), and one --set full capture of the top kernel
set -x
# programmatic dependent launch on the configs that are NOT power-bound (round 1 measured it neutral on the power-capped 1b step)
for M in pythia-410m roberta; do
  python bench.py --model $M --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_pdl0_$M.json 2> gpurun_out/r02_bench_pdl0_$M.err
  B200_PDL=1 python bench.py --model $M --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_pdl1_$M.json 2> gpurun_out/r02_bench_pdl1_$M.err
done
for M in pythia-1b pythia-410m roberta; do
  CMD="python bench.py --model $M --steps 1 --warmup 1 --grad-acc 1 --no-cpu-baseline"
  $CMD > gpurun_out/plain_$M.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_$M.csv $CMD > gpurun_out/ncu_$M.log 2>&1
done
python scripts/dev/prof_gemm.py > gpurun_out/plain_gemm_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 6 -o gpurun_out/r02_gemm_full python scripts/dev/prof_gemm.py > gpurun_out/ncu_gemm_full.log 2>&1
echo done
