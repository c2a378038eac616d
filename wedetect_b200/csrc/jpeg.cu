// jpeg.cu — JPEG decode on the device for the two image entry points: the reference decodes on the CPU, one
// image at a time (LoadImageFromFile -> mmcv.imfrombytes -> cv2.imdecode, config/wedetect_base.py:112, infer_wedetect.py:111;
// Image.open(...).convert("RGB"), generate_proposal.py:1089-1090).  Here nvJPEG (CUDA toolkit library, like cuBLAS a vendor
// library, not ours) writes interleaved BGR / RGB pixels straight into the source buffer of WD_OP_CV_RESIZE_PAD / WD_OP_LETTERBOX:
// only the compressed bytes cross PCIe.  libnvjpeg is opened lazily with dlopen, so the extension has no link-time dependency on
// it and every other entry point works where it is absent; wd_jpeg_open then fails loudly.
// nvJPEG's IDCT and chroma up-sampling are not bit-identical to libjpeg-turbo's (what cv2 / PIL call): the decoded pixels differ
// by a few grey levels on chroma-subsampled files (measured in tests/test_gpu_mm_pipeline.py), so this decode is opt-in and the
// parity statements of the library are made on host-decoded pixels.
#include "internal.h"
#include <dlfcn.h>
#include <nvjpeg.h>
#include <mutex>

namespace wd {
namespace {

struct NvjpegApi {
    void* lib = nullptr;
    decltype(&nvjpegCreateSimple) create = nullptr;
    decltype(&nvjpegDestroy) destroy = nullptr;
    decltype(&nvjpegJpegStateCreate) state_create = nullptr;
    decltype(&nvjpegJpegStateDestroy) state_destroy = nullptr;
    decltype(&nvjpegGetImageInfo) info = nullptr;
    decltype(&nvjpegDecode) decode = nullptr;
};

const NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.create = reinterpret_cast<decltype(api.create)>(dlsym(api.lib, "nvjpegCreateSimple"));
        api.destroy = reinterpret_cast<decltype(api.destroy)>(dlsym(api.lib, "nvjpegDestroy"));
        api.state_create = reinterpret_cast<decltype(api.state_create)>(dlsym(api.lib, "nvjpegJpegStateCreate"));
        api.state_destroy = reinterpret_cast<decltype(api.state_destroy)>(dlsym(api.lib, "nvjpegJpegStateDestroy"));
        api.info = reinterpret_cast<decltype(api.info)>(dlsym(api.lib, "nvjpegGetImageInfo"));
        api.decode = reinterpret_cast<decltype(api.decode)>(dlsym(api.lib, "nvjpegDecode"));
        if (!api.create || !api.destroy || !api.state_create || !api.state_destroy || !api.info || !api.decode) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api.lib ? &api : nullptr;
}

}  // namespace
}  // namespace wd

struct wd_jpeg {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
};

extern "C" int wd_jpeg_open(wd_jpeg** out) {
    using namespace wd;
    WD_REQUIRE(out, "wd_jpeg_open: null output");
    *out = nullptr;
    if (device_sm_count() <= 0) return -2;
    const NvjpegApi* api = nvjpeg_api();
    WD_REQUIRE(api, "wd_jpeg_open: libnvjpeg.so.12 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
    auto j = new wd_jpeg();
    nvjpegStatus_t st = api->create(&j->handle);
    if (st == NVJPEG_STATUS_SUCCESS) st = api->state_create(j->handle, &j->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        if (j->handle) api->destroy(j->handle);
        delete j;
        set_last_error("wd_jpeg_open: nvjpeg status %d", (int)st);
        return -1;
    }
    *out = j;
    return 0;
}

extern "C" int wd_jpeg_info(wd_jpeg* j, const uint8_t* data, uint64_t len, int* width, int* height, int* components, int* subsampling) {
    using namespace wd;
    WD_REQUIRE(j && data && len > 0 && width && height, "wd_jpeg_info: bad arguments");
    const NvjpegApi* api = nvjpeg_api();
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css = NVJPEG_CSS_UNKNOWN;
    const nvjpegStatus_t st = api->info(j->handle, data, (size_t)len, &nc, &css, ws, hs);
    WD_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "wd_jpeg_info: nvjpeg status %d (not a baseline / progressive JPEG?)", (int)st);
    *width = ws[0];
    *height = hs[0];
    if (components) *components = nc;
    if (subsampling) *subsampling = (int)css;
    return 0;
}

extern "C" int wd_jpeg_decode(wd_jpeg* j, const uint8_t* data, uint64_t len, uint8_t* dst, uint64_t pitch, int bgr, void* stream) {
    using namespace wd;
    WD_REQUIRE(j && data && len > 0 && dst && pitch > 0, "wd_jpeg_decode: bad arguments");
    const NvjpegApi* api = nvjpeg_api();
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) {
        img.channel[c] = nullptr;
        img.pitch[c] = 0;
    }
    img.channel[0] = dst;
    img.pitch[0] = (size_t)pitch;
    const nvjpegStatus_t st = api->decode(j->handle, j->state, data, (size_t)len, bgr ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream);
    WD_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "wd_jpeg_decode: nvjpeg status %d", (int)st);
    return 0;
}

extern "C" void wd_jpeg_close(wd_jpeg* j) {
    if (!j) return;
    const wd::NvjpegApi* api = wd::nvjpeg_api();
    if (api) {
        if (j->state) api->state_destroy(j->state);
        if (j->handle) api->destroy(j->handle);
    }
    delete j;
}
