set -x
python -m pytest tests/test_attention_gpu.py tests/test_roberta_gpu.py tests/test_model_gpu.py -q -s -p no:cacheprovider > gpurun_out/r02_gputest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest12.log
tail -5 gpurun_out/r02_gputest12.log
python bench.py --model roberta --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench12_roberta.json 2> gpurun_out/r02_bench12_roberta.err
python bench.py --model roberta --seq-len 128 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench12_roberta128.json 2> gpurun_out/r02_bench12_roberta128.err
python bench.py --model pythia-410m --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench12_410m.json 2> gpurun_out/r02_bench12_410m.err
python scripts/dev/attn_scaling.py 64 > gpurun_out/r02_attn_scaling_64_b.txt 2>&1
echo done
