// TEST INFRASTRUCTURE ONLY: empty stand-in so that `#include <cuda_runtime.h>`
// (reference cuda/modules/common.cu:92) resolves when the reference is compiled with g++.
#pragma once
