/* Drop-in for the CusADi-generated libsparse_pdipm_multiple_iterations.so (see
 * qp_former_evaluate.c): `evaluate` = 5 Newton iterations on the frozen 240/140/160 pattern. */
#include "../../../include/srbd_mpc.h"
float evaluate(const double** inputs, double* work, double** outputs, const int batch_size) {
  return srbd_sparse_pdipm_multiple_iterations_evaluate(inputs, work, outputs, batch_size);
}
