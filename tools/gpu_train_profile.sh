set -x
python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r02_gputest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_train.log
python tools/profile_train.py 4 0 > gpurun_out/r02_train_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv python tools/profile_train.py 4 0 > gpurun_out/r02_train_ncu.log 2>&1
tail -3 gpurun_out/r02_gputest_train.log; tail -2 gpurun_out/r02_train_plain.log
