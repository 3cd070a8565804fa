set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest4.log
python bench.py --config features --steps 5 --warmup 3 > gpurun_out/r02_features_v2.json 2> gpurun_out/r02_features_v2.err
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
LD_TRAIN_GRAPH=0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02_bench4_nograph.json 2> gpurun_out/r02_bench4_nograph.err
tail -3 gpurun_out/r02_gputest4.log
