// libpqn_b200.so: version + error slot.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/pqn_b200.h"
#include "api_common.h"

namespace pqn {
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return PQN_OK;
  cudaGetLastError();  // clear the sticky launch error so later calls report their own
  return set_error(PQN_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace pqn

extern "C" {
const char* pqn_last_error(void) { return pqn::g_err; }
int pqn_version(void) { return 100; }  // 0.1.0
}
