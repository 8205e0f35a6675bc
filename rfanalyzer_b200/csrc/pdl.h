// pdl.h -- programmatic dependent launch for the small kernels of the IQ -> audio chain.
//
// Behind the resampler an Airspy chain call handles ~10^5 samples in four to nine kernels of 4-10 us each, most of that
// launch latency.  Every such kernel is launched with the programmatic-stream-serialization attribute and starts with
//     pdl_enter();   // griddepcontrol.launch_dependents; griddepcontrol.wait;
// so its blocks are scheduled (and ITS dependents' blocks after it) while the kernel before it still runs, and the kernels
// then execute back to back without launch gaps.  griddepcontrol.wait returns when the prerequisite grid has COMPLETED
// and its memory is visible, and every kernel waits before it touches anything, so the ordering is exactly stream order
// (transitively: a kernel that has not passed its wait has not finished).  After a kernel, copy or memset that knows nothing
// about this the attribute is inert.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace rfa {

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

template <class... KArgs, class... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace rfa
