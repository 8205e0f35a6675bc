// kernels.h -- launchers of the non-spectrum kernels (convert.cu, fir.cu, demod.cu).
#pragma once
#include <cuda_runtime.h>

namespace rfa {

cudaError_t convert_launch(int fmt, const void *iq, long long n, float *re, float *im, int num_sms,
                           cudaStream_t st);
cudaError_t mix_launch(int fmt, const void *iq, long long n, float *re, float *im, const float *cosT,
                       const float *sinT, int len, int idx, int num_sms, cudaStream_t st);

}  // namespace rfa
