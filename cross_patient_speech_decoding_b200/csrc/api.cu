// Library-level C ABI: version, last error string, launch counter, device query.
#include "common.cuh"
#include <string.h>
#include "descs.h"

static thread_local char g_err[512] = "";
static long long g_launches = 0;

extern "C" void cpsd_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
extern "C" const char* cpsd_last_error(void) { return g_err; }
extern "C" void cpsd_count_launch(int n) { g_launches += n; }
extern "C" long long cpsd_launch_count(void) { return g_launches; }
extern "C" void cpsd_reset_launch_count(void) { g_launches = 0; }
extern "C" int cpsd_version(void) { return 100; }

// Returns the compute capability major*10+minor of the current device, or <0 on error.
extern "C" int cpsd_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}

// sizeof() of every descriptor record, for the host binding's layout check.
extern "C" int cpsd_desc_sizes(int* out) {
  out[0] = (int)sizeof(cpsd_gram_tn_desc);
  out[1] = (int)sizeof(cpsd_colsum_desc);
  out[2] = (int)sizeof(cpsd_proj_desc);
  out[3] = (int)sizeof(cpsd_gram_nt_desc);
  out[4] = (int)sizeof(cpsd_class_mean_desc);
  out[5] = (int)sizeof(cpsd_svm_desc);
  out[6] = (int)sizeof(cpsd_cca_desc);
  out[7] = 0;
  return CPSD_OK;
}
