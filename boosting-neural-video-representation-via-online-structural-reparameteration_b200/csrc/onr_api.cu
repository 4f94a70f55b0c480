// onr_api.cu — library-level entry points: ABI version, error string, device check, and the
// host-side TMA descriptor builders (cuTensorMapEncodeTiled fetched through the runtime so the
// library has no link-time dependency on libcuda).
#include "onr_common.cuh"

#include <string.h>

namespace onr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_act_tmap(CUtensorMap* map, const void* ptr, int B, int H, int W, int Cp, int s, int box_w,
                  int box_h, int inner) {
    EncodeTiledFn enc = get_encode();
    ONR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    ONR_REQUIRE(Cp % 32 == 0 && s >= 1 && box_w <= 256 && box_h <= 256 && (inner == 32 || inner == 64),
                "make_act_tmap: bad shape");
    const cuuint64_t Ws = (cuuint64_t)W * s;
    cuuint64_t dims[5] = {(cuuint64_t)s * Cp, (cuuint64_t)W, (cuuint64_t)s, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)2 * s * Cp, (cuuint64_t)2 * Ws * Cp, (cuuint64_t)2 * s * Ws * Cp,
                             (cuuint64_t)2 * s * H * Ws * Cp};
    cuuint32_t box[5] = {(cuuint32_t)inner, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ONR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: %d (B%d H%d W%d Cp%d s%d)", (int)r,
                B, H, W, Cp, s);
    return 0;
}

int make_weight_tmap(CUtensorMap* map, const void* ptr, int taps, int rows, int k, int box_rows, int inner) {
    EncodeTiledFn enc = get_encode();
    ONR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    ONR_REQUIRE(k % 32 == 0 && box_rows <= 256 && (inner == 32 || inner == 64), "make_weight_tmap: bad shape");
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)2 * k, (cuuint64_t)2 * k * rows};
    cuuint32_t box[3] = {(cuuint32_t)inner, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ONR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d", (int)r);
    return 0;
}

int make_f32_2d_tmap(CUtensorMap* map, const void* ptr, int rows, int k, int pitch, int box_rows) {
    EncodeTiledFn enc = get_encode();
    ONR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    ONR_REQUIRE(pitch % 4 == 0 && pitch >= k && box_rows >= 1 && box_rows <= 256 && ((uintptr_t)ptr & 15) == 0,
                "make_f32_2d_tmap: bad shape (rows %d k %d pitch %d box %d)", rows, k, pitch, box_rows);
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)4 * pitch};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ONR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32 2d) failed: %d (rows %d k %d pitch %d)", (int)r, rows, k,
                pitch);
    return 0;
}

}  // namespace onr

extern "C" {

int onr_abi_version(void) { return ONR_ABI_VERSION; }

const char* onr_last_error(void) { return onr::g_err; }

unsigned long long onr_launch_count(void) { return __atomic_load_n(&onr::g_launches, __ATOMIC_RELAXED); }

int onr_check_device(void) {
    int dev = 0;
    ONR_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    ONR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    ONR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    ONR_REQUIRE(major == 10, "liborepnerv is sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
}

}  // extern "C"
