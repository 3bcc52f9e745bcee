// pdl.cuh — programmatic dependent launch (PDL).
//
// A fused rollout is ~3 300 small dependent kernels; a decode pass at 4096 rows is ~90 kernels of 5-30 us each.  With
// cudaLaunchAttributeProgrammaticStreamSerialization the next kernel's CTAs are scheduled onto an SM as soon as the
// previous kernel's CTAs there have exited (every kernel triggers `launch_dependents` first thing), run their prologue
// (barrier init, TMEM allocation, tensor-map prefetch) and block in `griddepcontrol.wait` until the previous grid has
// completed and its memory is visible — so launch latency and prologues overlap the predecessor's tail.
// Rule: no global-memory access before pdl_wait().  Both instructions are no-ops for an ordinary launch.
// Measured on B200 (scripts/rollout_timing.py, 4096 episodes): PDL makes the rollout SLOWER (223.8 vs 208.2 ms inside the CUDA
// graph, 224.3 vs 212.5 ms eager), so the engine keeps it off by default (lmrl_set_use_pdl); the plumbing stays for re-measurement.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace lmrl {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

// host: launch `kernel` with (pdl != 0) or without the PDL attribute; `cluster` > 1 adds a cluster dimension
template <typename... P, typename... A>
inline cudaError_t launch_k(int pdl, int cluster, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}
}  // namespace lmrl
