// cuemu shim: stands in for <cuda_runtime.h> when kernels are compiled for the host emulator (tests only).
#pragma once
#include "cuemu.h"
