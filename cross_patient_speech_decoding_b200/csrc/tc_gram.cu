// tcgen05 / TMEM pooled-Gram kernel -- placeholder until the tensor-core path lands.
#include "common.cuh"
#include "descs.h"

extern "C" int cpsd_gram_nt_tc(const cpsd_gram_nt_desc* descs_dev, int nprob, int m_max, int n_max,
                               cudaStream_t stream) {
  (void)descs_dev; (void)nprob; (void)m_max; (void)n_max; (void)stream;
  cpsd_set_error("gram_nt_tc: tensor-core path not built in this revision");
  return CPSD_ERR_UNSUPPORTED;
}
