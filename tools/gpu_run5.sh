set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest5.log
python bench.py --config features --steps 5 --warmup 3 > gpurun_out/r02_features_v2.json 2> gpurun_out/r02_features_v2.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
ncu --set full --import-source on --clock-control none -k regex:fbank_kernel -c 1 -f -o gpurun_out/r02_fbank_v2 python tools/profile_fbank.py 1 > gpurun_out/r02_fbank_ncu2.log 2>&1
tail -3 gpurun_out/r02_gputest5.log
