/* Drop-in for the CusADi-generated libqp_former.so: exports the single symbol `evaluate`
 * that biped_pympc/cusadi/src/CusadiFunction.py:30-46 binds (signature from
 * biped_pympc/cusadi/src/generateCUDACode.py:160-163) and forwards to libsrbd_mpc.so. */
#include "../../../include/srbd_mpc.h"
float evaluate(const double** inputs, double* work, double** outputs, const int batch_size) {
  return srbd_qp_former_evaluate(inputs, work, outputs, batch_size);
}
