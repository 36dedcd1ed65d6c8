// Error plumbing shared by the translation units of libpqn_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace pqn {
// records a printf-style message in the thread-local error slot and returns `code`
int set_error(int code, const char* fmt, ...);
// cudaPeekAtLastError() after a launch: 0 or PQN_E_CUDA (message recorded)
int check_launch(const char* what);
}  // namespace pqn
