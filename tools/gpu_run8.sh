set -x
python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_gputest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest8.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err
LD_TRAIN_FUSE_BWD=0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02_bench8_nofuse.json 2> gpurun_out/r02_bench8_nofuse.err
LD_STEM_PX=4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --train-steps 0 > gpurun_out/r02_bench8_stem4.json 2> gpurun_out/r02_bench8_stem4.err
tail -3 gpurun_out/r02_gputest8.log
