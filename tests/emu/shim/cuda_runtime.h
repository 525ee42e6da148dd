// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime.h> in the FCX_EMU build (see ../cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
