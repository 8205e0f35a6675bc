// device_once.h -- "configure this kernel once per DEVICE".  cudaFuncSetAttribute applies to the current device only,
// and one process may hold contexts on several GPUs (the C ABI allows it even though the benchmark runs one process
// per GPU), so a plain function-local `static bool` would leave every device but the first unconfigured.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace rfa {

struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};  // bit d: device d is configured (device ordinals < 64)
    // true when the current device has not been configured through this object yet
    bool pending(int *dev_out) {
        int dev = 0;
        cudaGetDevice(&dev);
        *dev_out = dev;
        return !((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull);
    }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

}  // namespace rfa
