# Developer tool: what the driver runs at round end, in one gpurun call (outputs under gpurun_out/).
#   gpurun --timeout 1500 -- 'bash tools/validate_on_gpu.sh TAG'
tag=${1:-val}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/${tag}_rc.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_rc.log
( time python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?" >> gpurun_out/${tag}_rc.log
( time python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_rc.log
cat gpurun_out/${tag}_rc.log; tail -2 gpurun_out/${tag}_pytest.log
