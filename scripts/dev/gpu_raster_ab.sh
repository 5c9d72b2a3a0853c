set -x
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct"
B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0 python scripts/dev/prof_gemm.py 20 > gpurun_out/r02_gemm_time_old.txt 2>&1
B200_GEMM_HINTS=0 python scripts/dev/prof_gemm.py 20 > gpurun_out/r02_gemm_time_nohint.txt 2>&1
python scripts/dev/prof_gemm.py 20 > gpurun_out/r02_gemm_time_new.txt 2>&1
B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_raster_old.json 2> gpurun_out/r02_bench_raster_old.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_raster_new.json 2> gpurun_out/r02_bench_raster_new.err
B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_raster_old2.json 2> gpurun_out/r02_bench_raster_old2.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_raster_new2.json 2> gpurun_out/r02_bench_raster_new2.err
python scripts/dev/prof_gemm.py > gpurun_out/plain_gemm.log 2>&1 && ncu --metrics $M --clock-control none -k regex:gemm_kernel --csv --log-file gpurun_out/r02_ncu_gemm_new.csv python scripts/dev/prof_gemm.py > gpurun_out/ncu_gemm_new.log 2>&1
export B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0
python scripts/dev/prof_gemm.py > gpurun_out/plain_gemm_old.log 2>&1 && ncu --metrics $M --clock-control none -k regex:gemm_kernel --csv --log-file gpurun_out/r02_ncu_gemm_old.csv python scripts/dev/prof_gemm.py > gpurun_out/ncu_gemm_old.log 2>&1
echo done
