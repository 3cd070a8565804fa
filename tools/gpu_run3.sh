set -x
python tools/profile_fbank.py 2 > gpurun_out/r02_fbank_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:fbank_kernel -c 1 -f -o gpurun_out/r02_fbank python tools/profile_fbank.py 1 > gpurun_out/r02_fbank_ncu.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest3.log
tail -3 gpurun_out/r02_gputest3.log
