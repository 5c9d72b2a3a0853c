set -x
python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_roberta_gpu.py tests/test_fullshape_gpu.py -q -p no:cacheprovider > gpurun_out/r02_gputest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest13.log
tail -4 gpurun_out/r02_gputest13.log
B200_SEPARATE_BIAS_GRAD=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench13_sep.json 2> gpurun_out/r02_bench13_sep.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench13_fused.json 2> gpurun_out/r02_bench13_fused.err
B200_SEPARATE_BIAS_GRAD=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench13_sep2.json 2> gpurun_out/r02_bench13_sep2.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench13_fused2.json 2> gpurun_out/r02_bench13_fused2.err
echo done
