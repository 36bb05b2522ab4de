set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/s5_rc.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s5_rc.log
( time python bench.py --impl reference ) > gpurun_out/s5_bench_ref.json 2> gpurun_out/s5_bench_ref.err; echo "ref rc=$?" >> gpurun_out/s5_rc.log
( time python bench.py ) > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?" >> gpurun_out/s5_rc.log
cat gpurun_out/s5_rc.log
