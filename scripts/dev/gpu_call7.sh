set -x
python -m pytest tests/test_attention_gpu.py tests/test_engine_gpu.py tests/test_fullshape_gpu.py -q -x -p no:cacheprovider > gpurun_out/r02_gputest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest7.log
tail -3 gpurun_out/r02_gputest7.log
python scripts/dev/attn_scaling.py 256 > gpurun_out/r02_attn_scaling_256_split.txt 2>&1
B200_ATTN_TRACE=1 python scripts/dev/attn_trace_fwd.py 16 8 > gpurun_out/r02_fwd256s_trace_b16h8.txt 2>&1
B200_ATTN_FWD256_1WARP=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab7_1warp.json 2> gpurun_out/r02_bench_ab7_1warp.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab7_split.json 2> gpurun_out/r02_bench_ab7_split.err
B200_ATTN_FWD256_1WARP=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab7_1warp2.json 2> gpurun_out/r02_bench_ab7_1warp2.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_ab7_split2.json 2> gpurun_out/r02_bench_ab7_split2.err
echo done
