# End-of-round validation on one B200: smoke, the GPU test suite, the default bench line and the reference arm.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -2 gpurun_out/final_smoke.log; tail -3 gpurun_out/final_gputest.log
