for w in 8 10 11 12 13 14; do
  B=$((148*w))
  SRBD_MPC_LIB=$PWD/exp_lib/W$w/libsrbd_mpc.so python tools/time_pdipm.py $B 5 2>&1 | tail -1 | sed "s/^/warps=$w B=$B: /"
done
