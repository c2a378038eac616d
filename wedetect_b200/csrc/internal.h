// internal.h — host-side plumbing shared by the op implementations.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <vector>
#include <memory>
#include <atomic>
#include <utility>
#include "../../include/wedetect_b200.h"
#include "common.cuh"

namespace wd {

extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count();   // of the CURRENT device (cached per device ordinal)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the opt-in is per device, and one process may
// drive several GPUs.
cudaError_t ensure_smem_attr(const void* func, int bytes);

bool pdl_enabled();   // WD_NO_PDL=1 turns programmatic dependent launch off (A/B measurements)

// Launch with the programmatic-stream-serialization attribute (kernels launched this way call pdl_wait() before touching
// global memory), optionally as clusters of `cluster_x` CTAs.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// A compiled op: validated parameters + prebuilt TMA descriptors; `launch` enqueues its kernel(s).
struct CompiledOp {
    virtual ~CompiledOp() {}
    virtual int launch(cudaStream_t s) = 0;
    virtual int num_kernels() const { return 1; }
};

// Encode a tiled bf16/f32 tensor map (rank <= 5).  dims/box are innermost-first; strides are in BYTES
// for dims 1..rank-1.  swizzle128: inner box must span exactly 128 bytes.
// elem_strides (optional, innermost-first): traversal stride per dim; box[i] then spans box[i] source elements and
// delivers ceil(box[i] / elem_strides[i]) of them.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, const uint32_t* elem_strides = nullptr);

int compile_gemm(const wd_op& op, std::unique_ptr<CompiledOp>& out);
int compile_rowops(const wd_op& op, std::unique_ptr<CompiledOp>& out);  // LN / dwconv / stem / im2col / cast / text
int compile_postprocess(const wd_op& op, std::unique_ptr<CompiledOp>& out);
int compile_preprocess(const wd_op& op, std::unique_ptr<CompiledOp>& out);
int compile_mlp_fused(const wd_op& op, std::unique_ptr<CompiledOp>& out);    // ConvNeXt block MLP in one kernel (C = 128)   // letterbox (PIL-exact resize + paste)

}  // namespace wd
