# Developer tool (round 2): new fp32-mode test + full GPU suite, default bench, launch list, one full ncu capture of the fused tick
tag=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/${tag}_rc.log
( time python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_rc.log
python tools/time_pdipm.py > gpurun_out/${tag}_time.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 5 --warmup 3 --no-sub > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdipm_fast_kernel -s 3 -c 1 -o gpurun_out/${tag}_pdipm python bench.py --steps 3 --warmup 3 --no-sub > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_rc.log; tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_time.log
