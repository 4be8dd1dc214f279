// Host-side helpers shared by every translation unit of libvglamm.
#pragma once
#include <cuda_runtime.h>

#include "../../include/vglamm.h"

namespace vg {
int set_error(int code, const char* fmt, ...);   // records a thread-local message, returns `code`
void count_launch();                             // one kernel launch issued (vg_launch_count)
int check_launch(const char* what);              // cudaGetLastError -> VG_OK / VG_ECUDA
}  // namespace vg
