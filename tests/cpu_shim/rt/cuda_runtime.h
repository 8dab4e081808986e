// TEST INFRASTRUCTURE ONLY: what `#include <cuda_runtime.h>` resolves to when a product source is compiled with -DUA2_CPU_SHIM.
#pragma once
#include "../cuda_shim.h"
