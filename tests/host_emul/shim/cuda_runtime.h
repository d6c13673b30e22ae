// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when product sources are compiled against the CUDA
// execution-model emulation (tests/host_emul/cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
