// capi_core.h -- shared internals of the C ABI translation units (context, error state,
// device buffers).  Not installed; the public surface is include/rfa_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rfa_b200.h"
#include "rfa_fft_core.cuh"
#include "tuning.h"

namespace rfa {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define RFA_CK(call)                                        \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

#define RFA_REQUIRE(cond, ...)       \
    do {                             \
        if (!(cond)) {               \
            set_error(__VA_ARGS__);  \
            return RFA_ERR_INVALID;  \
        }                            \
    } while (0)

// grow-only device (or pinned host) buffer
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    int ensure(size_t bytes);
    void release();
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace rfa

struct rfa_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    long long launches = 0;
    rfa::Tuning tune;                                 // rfa_ctx_set_option
    // pipelined host-memory mode: copy engines run beside the compute stream
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    std::map<int, rfa::cf *> twiddles;                // per transform size
    std::map<std::pair<int, int>, float *> windows;   // per (kind, size)
    short *synth_table = nullptr;                     // generator's cosine table
    rfa::Buf stage[8];                                // staging for RFA_MEM_HOST calls
    rfa::Buf render[6];                               // rfa_render_waterfall: colour map, row means, host-mode outputs
    int get_twiddles(int n, const rfa::cf **out);
    int get_window(int kind, int n, const float **out);
    int use();  // cudaSetDevice
};
