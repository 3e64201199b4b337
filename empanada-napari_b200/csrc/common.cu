#include "common.cuh"

#include <cstdio>
#include <cstring>

static thread_local char g_err[512] = "";

extern "C" {

int be_set_error(const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "unknown error");
  return -1;
}

int be_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}

const char* be_last_error() { return g_err; }

int be_version() { return 100; }

}  // extern "C"
