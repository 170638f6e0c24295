// Kernel launch helper with programmatic dependent launch (PDL).
//
// The hot loop of a model call is ~670 short-to-medium kernels on one stream.  With the
// cudaLaunchAttributeProgrammaticStreamSerialization attribute a kernel may be scheduled while its predecessor is still
// draining: its CTAs run their prologue (mbarrier init, TMEM allocation, tensor-map prefetch) on SMs the predecessor has
// already left, then block in griddepcontrol.wait until the predecessor has completed and its writes are visible.
// Rule for every kernel launched through launch_k: no global-memory access (read OR write) before pdl_wait().
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace ndit {

// 1: launch with the PDL attribute.  Thread-local: an engine call sets it from its own option ("pdl" / NDIT_PDL) for the
// duration of the call, so two engines in one process do not share the switch.
extern thread_local int g_pdl;

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cudaFuncSetAttribute is per (function, device): remember it per device, not per process (two engines on two devices in
// one process).  `flags` is a function-local static of the launcher, one slot per device ordinal.
struct PerDeviceFlag {
    bool done[64] = {};
    bool& here() {
        int dev = 0;
        cudaGetDevice(&dev);
        return done[dev & 63];
    }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace ndit
