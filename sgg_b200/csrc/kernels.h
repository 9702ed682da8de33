// Internal launch helpers shared between translation units.
#pragma once
#include "common.cuh"
namespace sgg {
int launch_linear(const float *x, const float *w, const float *b, float *y, int M, int Nout, int K, int relu,
                  cudaStream_t st);
}
