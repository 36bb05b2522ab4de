python tools/time_pdipm.py
for d in "$@"; do SRBD_MPC_LIB=/root/repo/exp_lib/$d/libsrbd_mpc.so python tools/time_pdipm.py; done
