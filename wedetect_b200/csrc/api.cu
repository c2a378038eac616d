// api.cu — C ABI: error plumbing, TMA descriptor encoding, program (op list) executor, CUDA graphs.
#include "internal.h"
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <set>
#include <time.h>

namespace wd {

std::atomic<uint64_t> g_launch_count{0};
static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count() {
    static std::mutex mu;
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_last_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (sms[dev] > 0) return sms[dev];
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_last_error("cudaGetDeviceProperties failed");
        return -1;
    }
    if (prop.major != 10) {
        set_last_error("libwedetect_b200 requires an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
        return -1;
    }
    sms[dev] = prop.multiProcessorCount;
    return sms[dev];
}

cudaError_t ensure_smem_attr(const void* func, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({func, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.insert({func, dev});
    return e;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

bool pdl_enabled() {
    static const bool on = getenv("WD_NO_PDL") == nullptr;
    return on;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, const uint32_t* elem_strides) {
    EncodeTiledFn fn = get_encode_fn();
    WD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    WD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16-byte aligned", base);
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = elem_strides ? elem_strides[i] : 1;
        WD_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map box[%d]=%u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        WD_REQUIRE(gstr[i] % 16 == 0, "tensor map stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i]);
    }
    if (swizzle128) WD_REQUIRE(box[0] * (uint32_t)elem_bytes == 128, "swizzle-128B tensor map needs a 128-byte inner box");
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u] stride0=%llu", (int)r,
                       rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), (unsigned long long)(rank > 2 ? gdim[2] : 0),
                       (unsigned long long)(rank > 3 ? gdim[3] : 0), (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
                       rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, (unsigned long long)(rank > 1 ? gstr[0] : 0));
        return -3;
    }
    return 0;
}

static int compile_op(const wd_op& op, std::unique_ptr<CompiledOp>& out) {
    switch (op.kind) {
        case WD_OP_GEMM: return compile_gemm(op, out);
        case WD_OP_POSTPROCESS: return compile_postprocess(op, out);
        case WD_OP_LETTERBOX:
        case WD_OP_CV_RESIZE_PAD: return compile_preprocess(op, out);
        case WD_OP_MLP_FUSED: return compile_mlp_fused(op, out);
        case WD_OP_LN_ROWS:
        case WD_OP_DWCONV_LN:
        case WD_OP_STEM_PATCH:
        case WD_OP_IM2COL_S2:
        case WD_OP_CAST_BF16:
        case WD_OP_TEXT_EMBED:
        case WD_OP_ATTN_SMALL:
        case WD_OP_L2NORM_ROWS:
        case WD_OP_GATHER_ROWS:
        case WD_OP_FOLD_TEXT:
        case WD_OP_GATHER_EMBED:
        case WD_OP_SCALE_ROWS:
        case WD_OP_RETR_REDUCE: return compile_rowops(op, out);
        default: set_last_error("unknown op kind %d", op.kind); return -1;
    }
}

}  // namespace wd

struct wd_program {
    std::vector<std::unique_ptr<wd::CompiledOp>> ops;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
};

extern "C" {

const char* wd_last_error(void) { return wd::g_err; }
int wd_version(void) { return 200; }
float wd_act_plane_scale(void) { return WD_ACT_PLANE_SCALE; }
uint64_t wd_launch_count(void) { return wd::g_launch_count.load(); }

int wd_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        wd::set_last_error("wd_device_info: %s", cudaGetErrorString(e));
        return -2;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int wd_op_run(const wd_op* op, void* stream) {
    if (!op) {
        wd::set_last_error("wd_op_run: null op");
        return -1;
    }
    std::unique_ptr<wd::CompiledOp> c;
    int rc = wd::compile_op(*op, c);
    if (rc) return rc;
    return c->launch(reinterpret_cast<cudaStream_t>(stream));
}

int wd_program_create(const wd_op* ops, int n_ops, wd_program** out) {
    if (!ops || !out || n_ops <= 0) {
        wd::set_last_error("wd_program_create: bad arguments");
        return -1;
    }
    auto prog = new wd_program();
    for (int i = 0; i < n_ops; ++i) {
        std::unique_ptr<wd::CompiledOp> c;
        int rc = wd::compile_op(ops[i], c);
        if (rc) {
            char msg[900];
            snprintf(msg, sizeof(msg), "%s", wd::g_err);
            wd::set_last_error("op %d (kind %d): %s", i, ops[i].kind, msg);
            delete prog;
            return rc;
        }
        prog->kernels += c->num_kernels();
        prog->ops.push_back(std::move(c));
    }
    *out = prog;
    return 0;
}

int wd_program_run(wd_program* prog, void* stream) {
    if (!prog) {
        wd::set_last_error("wd_program_run: null program");
        return -1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    for (size_t i = 0; i < prog->ops.size(); ++i) {
        int rc = prog->ops[i]->launch(s);
        if (rc) {
            char msg[900];
            snprintf(msg, sizeof(msg), "%s", wd::g_err);
            wd::set_last_error("launch of op %zu failed: %s", i, msg);
            return rc;
        }
    }
    return 0;
}

int wd_program_num_ops(const wd_program* prog) { return prog ? (int)prog->ops.size() : 0; }

// Eager replay with a CUDA event between consecutive ops: ms_per_op[i] = device time of op i (which may be
// several kernels, e.g. the post-process).  Synchronises the stream at the end (measurement helper only).
int wd_program_run_timed(wd_program* prog, void* stream, float* ms_per_op) {
    if (!prog || !ms_per_op) {
        wd::set_last_error("wd_program_run_timed: bad arguments");
        return -1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = prog->ops.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) WD_CHECK_CUDA(cudaEventCreate(&e));
    WD_CHECK_CUDA(cudaEventRecord(ev[0], s));
    int rc = 0;
    for (size_t i = 0; i < n && rc == 0; ++i) {
        rc = prog->ops[i]->launch(s);
        cudaEventRecord(ev[i + 1], s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == 0 && e == cudaSuccess)
        for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&ms_per_op[i], ev[i], ev[i + 1]);
    for (auto& x : ev) cudaEventDestroy(x);
    if (rc) return rc;
    WD_CHECK_CUDA(e);
    return 0;
}

// Debug helper: run op by op, polling the stream after each launch; returns the index of the first op that does
// not finish within timeout_ms (-1 if all finish).  A stuck kernel is left running: the caller should exit.
int wd_program_find_stuck_op(wd_program* prog, void* stream, int timeout_ms, int* stuck_op) {
    if (!prog || !stuck_op) {
        wd::set_last_error("wd_program_find_stuck_op: bad arguments");
        return -1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    *stuck_op = -1;
    for (size_t i = 0; i < prog->ops.size(); ++i) {
        int rc = prog->ops[i]->launch(s);
        if (rc) return rc;
        cudaEvent_t ev;
        WD_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        WD_CHECK_CUDA(cudaEventRecord(ev, s));
        int waited = 0;
        cudaError_t q;
        while ((q = cudaEventQuery(ev)) == cudaErrorNotReady && waited < timeout_ms) {
            struct timespec ts = {0, 1000000};
            nanosleep(&ts, nullptr);
            ++waited;
        }
        cudaEventDestroy(ev);
        if (q == cudaErrorNotReady) {
            *stuck_op = (int)i;
            return 0;
        }
        if (q != cudaSuccess) {
            wd::set_last_error("op %zu failed: %s", i, cudaGetErrorString(q));
            *stuck_op = (int)i;
            return -2;
        }
    }
    return 0;
}

int wd_program_capture(wd_program* prog, void* stream) {
    if (!prog) {
        wd::set_last_error("wd_program_capture: null program");
        return -1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (prog->exec) {
        cudaGraphExecDestroy(prog->exec);
        prog->exec = nullptr;
    }
    if (prog->graph) {
        cudaGraphDestroy(prog->graph);
        prog->graph = nullptr;
    }
    WD_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = wd_program_run(prog, stream);
    cudaError_t e = cudaStreamEndCapture(s, &prog->graph);
    if (rc) return rc;
    WD_CHECK_CUDA(e);
    WD_CHECK_CUDA(cudaGraphInstantiate(&prog->exec, prog->graph, 0));
    return 0;
}

int wd_program_replay(wd_program* prog, void* stream) {
    if (!prog || !prog->exec) {
        wd::set_last_error("wd_program_replay: program not captured");
        return -1;
    }
    WD_CHECK_CUDA(cudaGraphLaunch(prog->exec, reinterpret_cast<cudaStream_t>(stream)));
    wd::count_launch(prog->kernels);
    return 0;
}

int wd_program_num_launches(const wd_program* prog) { return prog ? prog->kernels : 0; }

void wd_program_destroy(wd_program* prog) {
    if (!prog) return;
    if (prog->exec) cudaGraphExecDestroy(prog->exec);
    if (prog->graph) cudaGraphDestroy(prog->graph);
    delete prog;
}

}  // extern "C"
