# Developer tool: A/B build of the CUDA library with extra -D flags on the PDIPM translation unit.
#   bash tools/build_variant.sh NAME [-DSRBD_X=1 ...]   ->  exp_lib/NAME/libsrbd_mpc.so (+ ptxas log)
# Time it with:  SRBD_MPC_LIB=exp_lib/NAME/libsrbd_mpc.so python tools/time_pdipm.py
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/exp_lib/$name; mkdir -p $out
csrc=$root/biped-pympc_b200/csrc; lib=$root/biped-pympc_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas=-v "$@" \
  -c $csrc/srbd_pdipm.cu -o $out/srbd_pdipm.o 2> $out/ptxas.log
objs=""; for f in srbd_abi srbd_qp_form srbd_misc srbd_lowlevel srbd_t1; do objs="$objs $lib/$f.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $objs $out/srbd_pdipm.o -o $out/libsrbd_mpc.so
grep -A2 "pdipm_fast_kernelILi14ELb0" $out/ptxas.log | grep -E "registers|spill" | head -3
