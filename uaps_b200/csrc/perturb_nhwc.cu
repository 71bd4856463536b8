// Channels-last bf16 versions of the encoder-feature perturbations (utilities/UAPS_unet.py:156-185,
// applied at :227-231) for the bf16 / tcgen05 model path.  Same semantics as perturb.cu, different
// layout: x is [B, HW, C] bf16 with C a multiple of 8; a thread owns one 16-byte chunk (8 channels)
// of one pixel, G = C/8 consecutive threads cover a pixel, so a warp reads 512 contiguous bytes.
// Randomness is Philox4x32-10 only (one 128-bit draw -> eight 16-bit uniforms); injected-draw parity
// lives in the fp32 NCHW kernels.
#include <cuda_bf16.h>
#include "common.cuh"

namespace uaps {
namespace {

constexpr int PT = 256;
constexpr int BG = 8;                       // batch samples per work item of perturb3_nhwc_kernel
constexpr uint32_t kStreamNoise = 11, kStreamDrop = 12;

struct Bf8 { uint4 raw; };
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
// eight uniforms in [0,1) from one Philox block: 16 bits each
__device__ __forceinline__ void uniform8(uint64_t seed, uint64_t idx, uint32_t stream, float (&u)[8]) {
    uint32_t r[4];
    Philox::draw4(seed, idx, stream, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u[2 * i] = (float)(r[i] & 0xffffu) * (1.0f / 65536.0f);
        u[2 * i + 1] = (float)(r[i] >> 16) * (1.0f / 65536.0f);
    }
}

// attention[b,p] = mean_c x[b,p,c]; smax[b] = max_p attention (order-preserving atomicMax)
__global__ void __launch_bounds__(PT) fdrop_stats_nhwc_kernel(const uint4* __restrict__ x, int G, long long HW, int B,
                                                              float* __restrict__ attention, uint32_t* __restrict__ smax_enc) {
    grid_dep_launch();
    grid_dep_wait();
    const long long chunks_per_sample = HW * G;       // a multiple of 32 (checked by the caller): warps stay whole
    const float invC = 1.0f / (8 * G);
    __shared__ float s_m[PT / kWarp];
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        float mx = -INFINITY;
        const uint4* xs = x + (size_t)b * chunks_per_sample;
        const long long stride = (long long)gridDim.x * PT;
        auto use = [&](long long t, const uint4& r) {
            float v[8];
            unpack8(r, v);
            float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            for (int o = 1; o < G; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);     // G threads of a pixel are adjacent lanes
            s *= invC;
            if ((t % G) == 0) attention[(size_t)b * HW + t / G] = s;
            mx = fmaxf(mx, s);
        };
        // four independent 16-byte loads in flight per thread (one per iteration left HBM under-subscribed: 2.4 TB/s);
        // chunks_per_sample and the strides are multiples of 32, so every warp takes the same path and the shuffles stay whole
        long long t = (long long)blockIdx.x * PT + threadIdx.x;
        for (; t + 3 * stride < chunks_per_sample; t += 4 * stride) {
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = __ldg(xs + t + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) use(t + u * stride, r[u]);
        }
        for (; t < chunks_per_sample; t += stride) use(t, __ldg(xs + t));
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x < kWarp) {
            float m = (threadIdx.x < PT / kWarp) ? s_m[threadIdx.x] : -INFINITY;
            m = warp_max(m);
            if (threadIdx.x == 0) atomicMax(smax_enc + b, enc_ordered(m));
        }
        __syncthreads();
    }
}

// FWD: y_* = perturbed copies of x.  BWD (x = nullptr): dx = sum of the three upstream gradients through
// the same masks.  The masks/noise depend only on (seed, index, attention), never on a stored tensor.
template <bool BWD>
__global__ void __launch_bounds__(PT) perturb3_nhwc_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g_noise,
                                                           const uint4* __restrict__ g_drop, const uint4* __restrict__ g_fdrop,
                                                           uint64_t seed, float range, float p, float scale,
                                                           const float* __restrict__ attention, const uint32_t* __restrict__ smax_enc,
                                                           float u, uint4* __restrict__ y_noise, uint4* __restrict__ y_drop,
                                                           uint4* __restrict__ y_fdrop, uint4* __restrict__ dx, int G, long long HW, int B,
                                                           const uint64_t* __restrict__ seed_dev, const float* __restrict__ u_dev,
                                                           const uint4* __restrict__ g_extra1, const uint4* __restrict__ g_extra2) {
    grid_dep_launch();
    grid_dep_wait();
    if (seed_dev != nullptr) seed += *seed_dev;       // device-resident step state (uaps_step_begin)
    if (u_dev != nullptr) u = *u_dev;
    const long long chunks_per_sample = HW * G;
    // A work item = one in-sample position x a group of BG batch samples: the FeatureNoise draw is keyed by the in-sample
    // index (UAPS_unet.py:178-179 samples ONE noise tensor of shape x.shape[1:] for the whole batch), so it is drawn once per
    // item instead of once per sample -- the kernel is bound by the Philox arithmetic (ncu: 63 % issue-active at 56 % of the
    // HBM rate), and this removes 7/16 of it.  The BG loads of an item are independent.
    const int n_groups = (B + BG - 1) / BG;
    const long long total = chunks_per_sample * n_groups;
    for (long long w = (long long)blockIdx.x * PT + threadIdx.x; w < total; w += (long long)gridDim.x * PT) {
        const int bg = (int)(w / chunks_per_sample);
        const long long in_sample = w - (long long)bg * chunks_per_sample;
        const long long pix = in_sample / G;
        float nz[8];
        uniform8(seed, (uint64_t)in_sample, kStreamNoise, nz);          // shared by the batch: keyed by the in-sample index
#pragma unroll
        for (int i = 0; i < 8; ++i) nz[i] = (nz[i] * 2.f - 1.f) * range;
        const int b_end = min(B, (bg + 1) * BG);
#pragma unroll 2
        for (int b = bg * BG; b < b_end; ++b) {
        const long long t = (long long)b * chunks_per_sample + in_sample;
        float kp[8];
        uniform8(seed, (uint64_t)t, kStreamDrop, kp);
        float m = 0.f;                                                   // FeatureDropout mask; its inputs exist only when that branch is live
        if ((BWD ? (const void*)g_fdrop : (const void*)y_fdrop) != nullptr) {
            const float thr = __fmul_rn(dec_ordered(smax_enc[b]), u);
            m = (attention[(size_t)b * HW + pix] < thr) ? 1.f : 0.f;
        }
        float o[8];
        if constexpr (!BWD) {
            float v[8];
            unpack8(__ldg(x + t), v);
            if (y_noise != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(v[i], nz[i], v[i]);
                y_noise[t] = pack8(o);
            }
            if (y_drop != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = (kp[i] >= p) ? v[i] * scale : 0.f;
                y_drop[t] = pack8(o);
            }
            if (y_fdrop != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = v[i] * m;
                y_fdrop[t] = pack8(o);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
            float g[8];
            if (g_noise != nullptr) {
                unpack8(__ldg(g_noise + t), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(g[i], nz[i], g[i]);
            }
            if (g_drop != nullptr) {
                unpack8(__ldg(g_drop + t), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += (kp[i] >= p) ? g[i] * scale : 0.f;
            }
            if (g_fdrop != nullptr) {
                unpack8(__ldg(g_fdrop + t), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(g[i], m, o[i]);
            }
            // gradients of the UNPERTURBED uses of the same feature map (main decoder's skip connection, the next level's
            // max-pool): summed here instead of by two separate accumulation kernels (each a read-read-write pass)
            if (g_extra1 != nullptr) {
                unpack8(__ldg(g_extra1 + t), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += g[i];
            }
            if (g_extra2 != nullptr) {
                unpack8(__ldg(g_extra2 + t), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += g[i];
            }
            dx[t] = pack8(o);
        }
        }
    }
}

inline int grid1d(long long n) {
    long long want = ceil_div<long long>(n, PT), cap = (long long)device_info().sm_count * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}
inline bool valid_g(int C) { const int G = C / 8; return C % 8 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0; }

}  // namespace
}  // namespace uaps

using namespace uaps;

UAPS_API int uaps_fdrop_stats_nhwc(const void* x, int B, int C, int64_t HW, float* attention, uint32_t* smax_enc,
                                   cudaStream_t stream) {
    if (x == nullptr || attention == nullptr || smax_enc == nullptr || B <= 0 || HW <= 0) return UAPS_EINVAL;
    if (!valid_g(C)) return UAPS_ERANGE;                       // C in {8,16,32,...,256}: a pixel's chunks fit one warp
    if (!aligned_to(x, 16)) return UAPS_EALIGN;
    const int G = C / 8;
    if ((HW * G) % 32 != 0) return UAPS_ERANGE;                // whole warps per sample (true for every UNet_UAPS level)
    long long gx = ceil_div<long long>(HW * G, PT), cap = ceil_div<long long>((long long)device_info().sm_count * 8, B);
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)(gx < 1 ? 1 : gx), (unsigned)(B < 65535 ? B : 65535), 1);
    UAPS_LAUNCH(fdrop_stats_nhwc_kernel, grid, dim3(PT), 0, stream, reinterpret_cast<const uint4*>(x), G, (long long)HW, B, attention, smax_enc);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_perturb3_nhwc(const void* x, uint64_t seed, float noise_range, double p_drop, const float* attention,
                                const uint32_t* smax_enc, float u, void* y_noise, void* y_drop, void* y_fdrop, int B, int C,
                                int64_t HW, const uint64_t* seed_dev, const float* u_dev, cudaStream_t stream) {
    if (x == nullptr || B <= 0 || HW <= 0) return UAPS_EINVAL;
    if (y_fdrop != nullptr && (attention == nullptr || smax_enc == nullptr)) return UAPS_EINVAL;   // only FeatureDropout needs the statistics
    if (y_noise == nullptr && y_drop == nullptr && y_fdrop == nullptr) return UAPS_EINVAL;
    if (!valid_g(C) || !(p_drop >= 0.0 && p_drop < 1.0)) return UAPS_ERANGE;
    if (!aligned_to(x, 16) || !aligned_to(y_noise, 16) || !aligned_to(y_drop, 16) || !aligned_to(y_fdrop, 16)) return UAPS_EALIGN;
    const float pk = (float)(1.0 - p_drop);
    UAPS_LAUNCH(perturb3_nhwc_kernel<false>, dim3(grid1d(HW * (C / 8) * ((B + BG - 1) / BG))), dim3(PT), 0, stream,
        reinterpret_cast<const uint4*>(x), (const uint4*)nullptr, (const uint4*)nullptr, (const uint4*)nullptr, seed, noise_range,
        (float)p_drop, (float)(1.0 / (double)pk), attention, smax_enc, u, reinterpret_cast<uint4*>(y_noise),
        reinterpret_cast<uint4*>(y_drop), reinterpret_cast<uint4*>(y_fdrop), (uint4*)nullptr, C / 8, (long long)HW, B, seed_dev, u_dev,
        (const uint4*)nullptr, (const uint4*)nullptr);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_perturb3_nhwc_bwd(const void* g_noise, const void* g_drop, const void* g_fdrop, uint64_t seed,
                                    float noise_range, double p_drop, const float* attention, const uint32_t* smax_enc,
                                    float u, void* dx, int B, int C, int64_t HW, const uint64_t* seed_dev, const float* u_dev,
                                    const void* g_extra1, const void* g_extra2, cudaStream_t stream) {
    if (dx == nullptr || B <= 0 || HW <= 0) return UAPS_EINVAL;
    if (g_fdrop != nullptr && (attention == nullptr || smax_enc == nullptr)) return UAPS_EINVAL;
    if (g_noise == nullptr && g_drop == nullptr && g_fdrop == nullptr && g_extra1 == nullptr && g_extra2 == nullptr) return UAPS_EINVAL;
    if (!aligned_to(g_extra1, 16) || !aligned_to(g_extra2, 16)) return UAPS_EALIGN;
    if (!valid_g(C) || !(p_drop >= 0.0 && p_drop < 1.0)) return UAPS_ERANGE;
    if (!aligned_to(dx, 16) || !aligned_to(g_noise, 16) || !aligned_to(g_drop, 16) || !aligned_to(g_fdrop, 16)) return UAPS_EALIGN;
    const float pk = (float)(1.0 - p_drop);
    UAPS_LAUNCH(perturb3_nhwc_kernel<true>, dim3(grid1d(HW * (C / 8) * ((B + BG - 1) / BG))), dim3(PT), 0, stream,
        (const uint4*)nullptr, reinterpret_cast<const uint4*>(g_noise), reinterpret_cast<const uint4*>(g_drop),
        reinterpret_cast<const uint4*>(g_fdrop), seed, noise_range, (float)p_drop, (float)(1.0 / (double)pk), attention, smax_enc, u,
        (uint4*)nullptr, (uint4*)nullptr, (uint4*)nullptr, reinterpret_cast<uint4*>(dx), C / 8, (long long)HW, B, seed_dev, u_dev,
        reinterpret_cast<const uint4*>(g_extra1), reinterpret_cast<const uint4*>(g_extra2));
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
