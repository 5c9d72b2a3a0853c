set -x
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_gputest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest15.log
tail -4 gpurun_out/r02_gputest15.log
python bench.py --model pythia-2.8b --checkpointing --grad-acc 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench15_2p8b.json 2> gpurun_out/r02_bench15_2p8b.err
python bench.py --model roberta --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench15_roberta.json 2> gpurun_out/r02_bench15_roberta.err
echo done
