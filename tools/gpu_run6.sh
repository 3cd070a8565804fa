set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gpu_cut_sampler" > gpurun_out/r02_gputest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest6.log
python bench.py --config corpus --gpus 1 > gpurun_out/r02_corpus_n1.json 2> gpurun_out/r02_corpus_n1.err
python bench.py --config cli --steps 3 --warmup 1 --cpu-sample-seconds 60 > gpurun_out/r02_cli.json 2> gpurun_out/r02_cli.err
tail -3 gpurun_out/r02_gputest6.log; tail -c 600 gpurun_out/r02_corpus_n1.err
