// ABI bookkeeping: version, error strings, launch counter.
#include "common.cuh"

namespace irs { long long g_launches = 0; }

extern "C" int irs_abi_version(void) { return IRS_B200_ABI_VERSION; }

extern "C" const char* irs_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case IRS_E_BADARG: return "irs_b200: bad argument (null pointer or non-positive size)";
    case IRS_E_SHAPE: return "irs_b200: shape/alignment outside what the kernel supports";
    case IRS_E_WORKSPACE: return "irs_b200: workspace too small (use the *_workspace_bytes query)";
    case IRS_E_OVERFLOW: return "irs_b200: candidate buffer overflow";
    default: return "irs_b200: unknown error";
  }
}

extern "C" long long irs_launch_count(void) { return irs::g_launches; }
extern "C" void irs_launch_count_reset(void) { irs::g_launches = 0; }
