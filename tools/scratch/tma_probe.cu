// Probe: which inner-coordinate / box-width combinations does a non-swizzled fp32 / uint8 TMA tile load accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../../tw_invoice_unet_ocr_llm_b200/csrc/ptx.cuh"
using namespace ub;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap m, int rank, int cx, int cy, uint32_t bytes, float* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const uint32_t bar = base, dst = base + 1024;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, bytes);
        if (rank == 4) tma_load_4d(dst, &m, bar, cx, cy, 0, 0); else tma_load_3d(dst, &m, bar, cx, cy, 0);
        while (!mbar_try_wait(bar, 0)) {}
        const float* s = reinterpret_cast<const float*>(sm + (dst - smem_u32(sm)));
        for (uint32_t i = 0; i < bytes / 4; ++i) out[i] = s[i];
    }
}
int main(int argc, char** argv) {
    const int only_bw = argc > 1 ? atoi(argv[1]) : 0, only_cx = argc > 2 ? atoi(argv[2]) : 99;
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int W = 48, H = 40, C = 3, N = 2;
    std::vector<float> h(size_t(N) * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 1000);
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1 << 16);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int bw : {12, 16}) for (int cx : {-1, 0, 4, 7, -4}) {
        if ((only_bw && bw != only_bw) || (only_cx != 99 && cx != only_cx)) continue;
        CUtensorMap m;
        cuuint64_t dims[4] = {W, H, C, N}, strides[3] = {W * 4ull, H * W * 4ull, C * H * W * 4ull};
        cuuint32_t box[4] = {(cuuint32_t)bw, 18, 3, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("f32 box %d: encode failed %d\n", bw, (int)r); continue; }
        probe<<<1, 32, 32768>>>(m, 4, cx, -1, bw * 18 * 3 * 4, out);
        cudaError_t e = cudaDeviceSynchronize();
        float o[40]; if (!e) cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
        printf("f32 box %2d cx %2d: %s  row1: %g %g %g %g\n", bw, cx, cudaGetErrorString(e), e ? 0 : o[bw], e ? 0 : o[bw + 1], e ? 0 : o[bw + 2], e ? 0 : o[bw + 3]);
        if (e) return 1;
    }
    printf("done\n");
    return 0;
}
