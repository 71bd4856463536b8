// Bilinear x2 upsampling with align_corners=True (nn.Upsample in UpBlock, utilities/UAPS_unet.py:74-75,84)
// and 2x2 max pooling (nn.MaxPool2d(2) in DownBlock, :56) on channels-last bf16 activations, forward and
// backward.  One thread owns a 16-byte chunk (8 channels) of one OUTPUT pixel (forward) or one INPUT
// pixel (backward: gather form, no atomics), so every access is a full 16-byte vector and a warp
// covers 512 contiguous bytes.
#include <cuda_bf16.h>
#include "common.cuh"

namespace uaps {
namespace {

constexpr int RT = 256;

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
}

// source coordinate of output index o: o * (in-1)/(out-1)  (align_corners=True), as torch computes it in fp32
__device__ __forceinline__ void src_coord(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
    const float s = scale * o;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

// Both resampling kernels use a 3-D grid -- x: 16-byte chunks along a row (pixel * G + g), y: rows in bands of
// RY, z: image -- so the index math is 32-bit (the 1-D version spent most of its time in 64-bit div/mod) and a CTA
// covers a 2-D patch: the 2 x 2 (forward) / up-to-4 x 4 (backward) stencils of neighbouring threads overlap in L1.
constexpr int RX = 32, RY = 8;            // CTA = 32 chunks x 8 rows = 256 threads

__global__ void __launch_bounds__(RT) upsample2x_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W,
                                                            int G, float sh, float sw) {
    grid_dep_launch();
    grid_dep_wait();
    const int OH = 2 * H, OW = 2 * W;
    const unsigned col = blockIdx.x * RX + threadIdx.x;
    const int oy = blockIdx.y * RY + threadIdx.y;
    if (col >= (unsigned)OW * G || oy >= OH) return;
    const int ox = col / G, g = col - ox * G;
    int y0, y1, x0, x1; float ly, lx;
    src_coord(oy, sh, H, y0, y1, ly);
    src_coord(ox, sw, W, x0, x1, lx);
    const uint4* xb = x + (size_t)blockIdx.z * H * W * G;
    float a[8], c[8], d[8], e[8], o[8];
    unpack8(__ldg(xb + ((size_t)y0 * W + x0) * G + g), a);
    unpack8(__ldg(xb + ((size_t)y0 * W + x1) * G + g), c);
    unpack8(__ldg(xb + ((size_t)y1 * W + x0) * G + g), d);
    unpack8(__ldg(xb + ((size_t)y1 * W + x1) * G + g), e);
    const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = hy * (hx * a[i] + lx * c[i]) + ly * (hx * d[i] + lx * e[i]);
    y[((size_t)blockIdx.z * OH + oy) * OW * G + col] = pack8(o);
}

// weight of output index o onto input index i along one axis (0 when o's two taps miss i)
__device__ __forceinline__ float tap_weight(int o, int i, float scale, int in_size) {
    int i0, i1; float l1;
    src_coord(o, scale, in_size, i0, i1, l1);
    return (i0 == i ? 1.f - l1 : 0.f) + (i1 == i ? l1 : 0.f);
}

// gather backward: input pixel (iy, ix) collects from the outputs whose stencil touches it.  Output o taps
// floor(o * r) and its successor, r = (in-1)/(2in-1) in [1/3, 1/2): the candidates of input i are o in
// [2i-2, 2i+3] (the fp32 evaluation of o * r is the forward's, so borderline taps land where the forward put them).
constexpr int NCAND = 6;
__global__ void __launch_bounds__(RT) upsample2x_bwd_kernel(const uint4* __restrict__ gy, uint4* __restrict__ gx, int H, int W,
                                                            int G, float sh, float sw) {
    grid_dep_launch();
    grid_dep_wait();
    const int OH = 2 * H, OW = 2 * W;
    const unsigned col = blockIdx.x * RX + threadIdx.x;
    const int iy = blockIdx.y * RY + threadIdx.y;
    if (col >= (unsigned)W * G || iy >= H) return;
    const int ix = col / G, g = col - ix * G;
    float wy[NCAND], wx[NCAND];
#pragma unroll
    for (int k = 0; k < NCAND; ++k) {
        const int oy = 2 * iy - 2 + k, ox = 2 * ix - 2 + k;
        wy[k] = (oy >= 0 && oy < OH) ? tap_weight(oy, iy, sh, H) : 0.f;
        wx[k] = (ox >= 0 && ox < OW) ? tap_weight(ox, ix, sw, W) : 0.f;
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const uint4* gb = gy + (size_t)blockIdx.z * OH * OW * G + g;
#pragma unroll
    for (int a = 0; a < NCAND; ++a) {
        if (wy[a] == 0.f) continue;
        const uint4* row = gb + (size_t)(2 * iy - 2 + a) * OW * G;
#pragma unroll
        for (int b = 0; b < NCAND; ++b) {
            if (wx[b] == 0.f) continue;
            float v[8];
            unpack8(__ldg(row + (size_t)(2 * ix - 2 + b) * G), v);
            const float w = wy[a] * wx[b];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
        }
    }
    gx[((size_t)blockIdx.z * H + iy) * W * G + col] = pack8(acc);
}

// [B,C,H,W] fp32 (NCHW) -> [B,H,W,Cp] bf16 (channels-last, channels C..Cp-1 zero): the network input and the
// logits gradient entering the bf16 path.  One thread per pixel: per channel a warp reads 128 contiguous bytes.
__global__ void __launch_bounds__(RT) nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, uint4* __restrict__ out, int C,
                                                                   int HW, int Gp) {
    grid_dep_launch();
    grid_dep_wait();
    const int p = blockIdx.x * RT + threadIdx.x;
    if (p >= HW) return;
    const float* xb = x + (size_t)blockIdx.y * C * HW + p;
    uint4* ob = out + ((size_t)blockIdx.y * HW + p) * Gp;
    for (int g = 0; g < Gp; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = g * 8 + i;
            v[i] = c < C ? __ldg(xb + (size_t)c * HW) : 0.f;
        }
        ob[g] = pack8(v);
    }
}

__global__ void __launch_bounds__(RT) maxpool2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int G) {
    grid_dep_launch();
    grid_dep_wait();
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)B * OH * OW * G;
    for (long long t = (long long)blockIdx.x * RT + threadIdx.x; t < total; t += (long long)gridDim.x * RT) {
        const int g = (int)(t % G);
        long long p = t / G;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH);
        const int b = (int)(p / OH);
        const uint4* xb = x + ((size_t)b * H * W + (size_t)(2 * oy) * W + 2 * ox) * G + g;
        float a[8], c[8], d[8], e[8];
        unpack8(__ldg(xb), a); unpack8(__ldg(xb + G), c);
        unpack8(__ldg(xb + (size_t)W * G), d); unpack8(__ldg(xb + (size_t)W * G + G), e);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(fmaxf(a[i], c[i]), fmaxf(d[i], e[i]));
        y[t] = pack8(a);
    }
}

// the gradient goes to the first maximum of the window in (row, column) order, as torch's max_pool2d backward does
__global__ void __launch_bounds__(RT) maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gy,
                                                          uint4* __restrict__ gx, int B, int H, int W, int G) {
    grid_dep_launch();
    grid_dep_wait();
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)B * OH * OW * G;
    for (long long t = (long long)blockIdx.x * RT + threadIdx.x; t < total; t += (long long)gridDim.x * RT) {
        const int g = (int)(t % G);
        long long p = t / G;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH);
        const int b = (int)(p / OH);
        const size_t base = ((size_t)b * H * W + (size_t)(2 * oy) * W + 2 * ox) * G + g;
        const size_t offs[4] = {0, (size_t)G, (size_t)W * G, (size_t)W * G + G};
        float v[4][8], gv[8], o[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8(__ldg(x + base + offs[k]), v[k]);
        unpack8(__ldg(gy + t), gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int best = 0;
            float m = v[0][i];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k][i] > m) { m = v[k][i]; best = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][i] = (k == best) ? gv[i] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) gx[base + offs[k]] = pack8(o[k]);
    }
}

inline int rgrid(long long n) {
    long long want = ceil_div<long long>(n, RT), cap = (long long)device_info().sm_count * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace
}  // namespace uaps

using namespace uaps;

UAPS_API int uaps_upsample2x_nhwc(const void* x, void* y, int B, int H, int W, int C, int backward, cudaStream_t stream) {
    if (x == nullptr || y == nullptr || B <= 0 || H <= 0 || W <= 0) return UAPS_EINVAL;
    if (C % 8 != 0 || C <= 0) return UAPS_ERANGE;
    if (!aligned_to(x, 16) || !aligned_to(y, 16)) return UAPS_EALIGN;
    const int G = C / 8;
    const float sh = H > 1 ? (float)(H - 1) / (float)(2 * H - 1) : 0.f, sw = W > 1 ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
    if (B > 65535 || ceil_div(2 * H, RY) > 65535) return UAPS_ERANGE;
    const dim3 block(RX, RY);
    if (!backward)      // x: [B,H,W,C] -> y: [B,2H,2W,C]
        UAPS_LAUNCH(upsample2x_fwd_kernel, dim3(ceil_div(2 * W * G, RX), ceil_div(2 * H, RY), B), block, 0, stream,
                    reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W, G, sh, sw);
    else                // x: upstream gradient [B,2H,2W,C] -> y: [B,H,W,C]
        UAPS_LAUNCH(upsample2x_bwd_kernel, dim3(ceil_div(W * G, RX), ceil_div(H, RY), B), block, 0, stream,
                    reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W, G, sh, sw);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_maxpool2_nhwc(const void* x, const void* gy, void* out, int B, int H, int W, int C, cudaStream_t stream) {
    if (x == nullptr || out == nullptr || B <= 0 || H <= 0 || W <= 0) return UAPS_EINVAL;
    if (C % 8 != 0 || C <= 0 || (H % 2) != 0 || (W % 2) != 0) return UAPS_ERANGE;
    if (!aligned_to(x, 16) || !aligned_to(out, 16) || !aligned_to(gy, 16)) return UAPS_EALIGN;
    const int G = C / 8;
    const long long n = (long long)B * (H / 2) * (W / 2) * G;
    if (gy == nullptr)  // forward: x [B,H,W,C] -> out [B,H/2,W/2,C]
        UAPS_LAUNCH(maxpool2_fwd_kernel, dim3(rgrid(n)), dim3(RT), 0, stream, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), B, H, W, G);
    else                // backward: out = d x [B,H,W,C] (every element written)
        UAPS_LAUNCH(maxpool2_bwd_kernel, dim3(rgrid(n)), dim3(RT), 0, stream, reinterpret_cast<const uint4*>(x),
                    reinterpret_cast<const uint4*>(gy), reinterpret_cast<uint4*>(out), B, H, W, G);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_nchw_f32_to_nhwc_bf16(const float* x, void* out, int B, int C, int H, int W, int Cp, cudaStream_t stream) {
    if (x == nullptr || out == nullptr || B <= 0 || C <= 0 || H <= 0 || W <= 0) return UAPS_EINVAL;
    if (Cp % 8 != 0 || Cp < C || B > 65535 || (long long)H * W >= 2147483647LL) return UAPS_ERANGE;
    if (!aligned_to(x, 4) || !aligned_to(out, 16)) return UAPS_EALIGN;
    const int HW = H * W;
    UAPS_LAUNCH(nchw_f32_to_nhwc_bf16_kernel, dim3(ceil_div(HW, RT), B), dim3(RT), 0, stream, x, reinterpret_cast<uint4*>(out), C, HW, Cp / 8);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
