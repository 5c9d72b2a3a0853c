set -x
python -m pytest tests/test_training_script_gpu.py -q --tb=short -s -p no:cacheprovider > gpurun_out/r02_gputest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest5.log
python scripts/dev/attn_scaling.py 256 > gpurun_out/r02_attn_scaling_256.txt 2>&1
python scripts/dev/attn_scaling.py 64 > gpurun_out/r02_attn_scaling_64.txt 2>&1
python -m pytest tests/test_fullshape_gpu.py -q -s -p no:cacheprovider > gpurun_out/r02_fullshape_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_fullshape_parity.log
echo done
