set -x
python -m pytest tests/test_training_script_gpu.py tests/test_fp16_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/r02_gputest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest4.log
B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab4_old.json 2> gpurun_out/r02_bench_ab4_old.err
B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab4_nohint.json 2> gpurun_out/r02_bench_ab4_nohint.err
B200_GEMM_GROUP_M=8 B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab4_old2.json 2> gpurun_out/r02_bench_ab4_old2.err
B200_GEMM_HINTS=0 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab4_nohint2.json 2> gpurun_out/r02_bench_ab4_nohint2.err
python -m pytest tests/test_attention_gpu.py -q -k "64" -p no:cacheprovider > gpurun_out/attn64_plain.log 2>&1 && timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_attention_gpu.py -q -k "64" -p no:cacheprovider > gpurun_out/r02_memcheck_attn_hd64.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck_attn_hd64.log
echo done
