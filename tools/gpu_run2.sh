set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest2.log
for occ in 2 3 4; do LD_FBANK_OCC=$occ python bench.py --config features --steps 5 --warmup 3 > gpurun_out/r02_features_occ$occ.json 2> gpurun_out/r02_features_occ$occ.err; done
LD_GEMM_PDL=0 python bench.py --steps 3 --warmup 2 --train-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_pdl0.json 2> gpurun_out/r02_bench_pdl0.err
LD_GEMM_PDL=1 python bench.py --steps 3 --warmup 2 --train-steps 0 --no-cpu-baseline --no-parity > gpurun_out/r02_bench_pdl1.json 2> gpurun_out/r02_bench_pdl1.err
python bench.py --precision split --steps 3 --warmup 2 --train-steps 0 --no-cpu-baseline > gpurun_out/r02_bench_split.json 2> gpurun_out/r02_bench_split.err
tail -3 gpurun_out/r02_gputest2.log
