// Error convention of the C-ABI: every entry point returns 0 on success or a negative code and
// leaves a thread-local message retrievable through be_last_error().
#pragma once
#include <cuda_runtime.h>

extern "C" {
int be_set_error(const char* msg);       // records msg, returns -1
int be_check_launch(const char* what);   // cudaGetLastError() -> 0 / -1 (message recorded)
const char* be_last_error();
}
