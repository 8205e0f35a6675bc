// kernels.h -- launchers of the non-spectrum kernels (convert.cu, fir.cu, demod.cu).
#pragma once
#include <cuda_runtime.h>

namespace rfa {

cudaError_t convert_launch(int fmt, const void *iq, long long n, float *re, float *im, int num_sms,
                           cudaStream_t st);
cudaError_t mix_launch(int fmt, const void *iq, long long n, float *re, float *im, const float *cosT,
                       const float *sinT, int len, int idx, int num_sms, cudaStream_t st);

struct SynthComp {
    unsigned int step;
    int amp;
    unsigned int mod_step;
    int mod_k;
};
void synth_make_table(short *tab /*4096*/);
cudaError_t synth_launch(int fmt, unsigned int seed, const SynthComp *comps, int ncomp, int noise_shift,
                         unsigned long long first, long long nsamples, const short *tab_dev, void *out,
                         int num_sms, cudaStream_t st);

}  // namespace rfa
