# Developer tool: one full ncu capture of the fused tick's PDIPM kernel (outputs under gpurun_out/).
tag=${1:-prof}
python bench.py --steps 3 --warmup 3 --no-sub > gpurun_out/${tag}_bench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:pdipm_fast_kernel -s 3 -c 1 -o gpurun_out/${tag}_pdipm python bench.py --steps 3 --warmup 3 --no-sub > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
