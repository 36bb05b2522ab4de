# Developer tool: time several A/B builds of the library in one gpurun call.
#   gpurun -- 'bash tools/time_variants.sh base A B'   (exp_lib/<name>/libsrbd_mpc.so, tools/build_variant.sh)
for rep in 1 2; do
for v in "$@"; do
  SRBD_MPC_LIB=$PWD/exp_lib/$v/libsrbd_mpc.so python tools/time_pdipm.py 2>&1 | tail -1 | sed "s/^/$v: /"
done; done | tee -a gpurun_out/variants.log
