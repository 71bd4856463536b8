// Encoder-feature perturbations of the auxiliary decoders (utilities/UAPS_unet.py:156-185, applied
// to all five feature levels at :227-231) and the encoder-block dropout (:40).
// Streaming elementwise kernels over NCHW fp32: 128-bit loads/stores, grid sized to the SM count.
// Randomness is injected (parity mode) or regenerated from Philox(seed, element index), so the
// backward pass never needs a stored mask.
#include "common.cuh"

namespace uaps {
namespace {

constexpr int PT = 256;                 // threads per block
constexpr int kWaves = 8;               // resident-CTA multiples of the SM count for grid-stride kernels

inline int stream_grid(long long nvec) {
    long long want = ceil_div<long long>(nvec, PT);
    long long cap = (long long)device_info().sm_count * kWaves;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// stream ids keep the three uses of one seed independent
constexpr uint32_t kStreamNoise = 1, kStreamDrop = 2;

__device__ __forceinline__ void philox_noise4(uint64_t seed, uint64_t vec_idx, float range, float (&n)[4]) {
    uint32_t r[4];
    Philox::draw4(seed, vec_idx, kStreamNoise, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] = (Philox::u01(r[i]) * 2.f - 1.f) * range;
}
__device__ __forceinline__ void philox_keep4(uint64_t seed, uint64_t vec_idx, float p, float (&k)[4]) {
    uint32_t r[4];
    Philox::draw4(seed, vec_idx, kStreamDrop, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (Philox::u01(r[i]) >= p) ? 1.f : 0.f;
}

// y = x * n + x with separately rounded multiply and add, as torch's x.mul(n) + x (:180)
template <int VEC>
__global__ void __launch_bounds__(PT) feature_noise_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                           uint64_t seed, float range, float* __restrict__ y,
                                                           int B, long long chw) {
    const long long nv = chw / VEC;                       // vectors per sample
    const long long stride = (long long)gridDim.x * PT;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += stride) {
        float n[VEC];
        if (noise != nullptr) {
            load_vec<VEC>(noise + v * VEC, n);
        } else {
            // the draw for element e is lane e%4 of Philox block e/4, whatever VEC is
            float n4[4];
            philox_noise4(seed, (uint64_t)(v * VEC) / 4, range, n4);
#pragma unroll
            for (int i = 0; i < VEC; ++i) n[i] = n4[(v * VEC + i) & 3];
        }
        for (int b = 0; b < B; ++b) {                      // the noise is shared by the batch (:178-179)
            float xv[VEC], yv[VEC];
            load_vec<VEC>(x + (size_t)b * chw + v * VEC, xv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) yv[i] = __fadd_rn(__fmul_rn(xv[i], n[i]), xv[i]);
            store_vec<VEC>(y + (size_t)b * chw + v * VEC, yv);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(PT) dropout_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep,
                                                     uint64_t seed, float p, float scale, float* __restrict__ y,
                                                     long long n) {
    const long long nv = n / VEC;
    const long long stride = (long long)gridDim.x * PT;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += stride) {
        float xv[VEC], yv[VEC], k[VEC];
        load_vec<VEC>(x + v * VEC, xv);
        if (keep != nullptr) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) k[i] = keep[v * VEC + i] ? 1.f : 0.f;
        } else {
            float k4[4];
            philox_keep4(seed, (uint64_t)(v * VEC) / 4, p, k4);
#pragma unroll
            for (int i = 0; i < VEC; ++i) k[i] = k4[(v * VEC + i) & 3];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) yv[i] = xv[i] * k[i] * scale;
        store_vec<VEC>(y + v * VEC, yv);
    }
}

// attention[b,hw] = mean_c x[b,c,hw]; per-sample max via order-preserving atomicMax.
// grid = (blocks over hw, B); one thread owns VEC adjacent pixels and walks the C planes (each
// plane read is a coalesced 128-bit access across the warp).
template <int VEC>
__global__ void __launch_bounds__(PT) fdrop_stats_kernel(const float* __restrict__ x, int C, long long HW,
                                                         float* __restrict__ attention, uint32_t* __restrict__ smax_enc) {
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * C * HW;
    const long long nv = HW / VEC;
    const float invC = 1.0f / C;
    float mx = -INFINITY;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += (long long)gridDim.x * PT) {
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            float xv[VEC];
            load_vec<VEC>(xb + (size_t)c * HW + v * VEC, xv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += xv[i];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[i] *= invC; mx = fmaxf(mx, acc[i]); }
        store_vec<VEC>(attention + (size_t)b * HW + v * VEC, acc);
    }
    mx = warp_max(mx);
    __shared__ float s_m[PT / kWarp];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < kWarp) {
        float m = (threadIdx.x < PT / kWarp) ? s_m[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) atomicMax(smax_enc + b, enc_ordered(m));
    }
}

template <int VEC>
__global__ void __launch_bounds__(PT) fdrop_apply_kernel(const float* __restrict__ x, const float* __restrict__ attention,
                                                         const uint32_t* __restrict__ smax_enc, float u,
                                                         float* __restrict__ y, int C, long long HW) {
    const int b = blockIdx.y;
    const float thr = __fmul_rn(dec_ordered(smax_enc[b]), u);          // threshold = max_val * u (:165)
    const size_t off = (size_t)b * C * HW;
    const long long nv = HW / VEC;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += (long long)gridDim.x * PT) {
        float a[VEC], m[VEC];
        load_vec<VEC>(attention + (size_t)b * HW + v * VEC, a);
#pragma unroll
        for (int i = 0; i < VEC; ++i) m[i] = (a[i] < thr) ? 1.f : 0.f; // drop_mask (:167)
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            float xv[VEC];
            load_vec<VEC>(x + off + (size_t)c * HW + v * VEC, xv);
#pragma unroll
            for (int i = 0; i < VEC; ++i) xv[i] *= m[i];
            store_vec<VEC>(y + off + (size_t)c * HW + v * VEC, xv);
        }
    }
}

// one read of x, up to three perturbed copies written
template <int VEC>
__global__ void __launch_bounds__(PT) perturb3_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                      const uint8_t* __restrict__ keep, uint64_t seed, float range,
                                                      float p, float scale, const float* __restrict__ attention,
                                                      const uint32_t* __restrict__ smax_enc, float u,
                                                      float* __restrict__ y_noise, float* __restrict__ y_drop,
                                                      float* __restrict__ y_fdrop, int C, long long HW) {
    const int b = blockIdx.y;
    const long long chw = (long long)C * HW;
    const size_t off = (size_t)b * chw;
    const float thr = (y_fdrop != nullptr) ? __fmul_rn(dec_ordered(smax_enc[b]), u) : 0.f;
    const long long nv = HW / VEC;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += (long long)gridDim.x * PT) {
        float m[VEC];
        if (y_fdrop != nullptr) {
            float a[VEC];
            load_vec<VEC>(attention + (size_t)b * HW + v * VEC, a);
#pragma unroll
            for (int i = 0; i < VEC; ++i) m[i] = (a[i] < thr) ? 1.f : 0.f;
        }
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const long long e = (long long)c * HW + v * VEC;           // element index inside the sample
            float xv[VEC], o[VEC];
            load_vec<VEC>(x + off + e, xv);
            if (y_noise != nullptr) {
                float n[VEC];
                if (noise != nullptr) load_vec<VEC>(noise + e, n);
                else {
                    float n4[4];
                    philox_noise4(seed, (uint64_t)e / 4, range, n4);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) n[i] = n4[(e + i) & 3];
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = __fadd_rn(__fmul_rn(xv[i], n[i]), xv[i]);
                store_vec<VEC>(y_noise + off + e, o);
            }
            if (y_drop != nullptr) {
                float k[VEC];
                if (keep != nullptr) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) k[i] = keep[off + e + i] ? 1.f : 0.f;
                } else {
                    float k4[4];
                    philox_keep4(seed, (uint64_t)(off + e) / 4, p, k4);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) k[i] = k4[(off + e + i) & 3];
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = xv[i] * k[i] * scale;
                store_vec<VEC>(y_drop + off + e, o);
            }
            if (y_fdrop != nullptr) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = xv[i] * m[i];
                store_vec<VEC>(y_fdrop + off + e, o);
            }
        }
    }
}

// 1/(1-p) the way torch's fused dropout forms it: p_keep rounded to fp32, reciprocal in double, result to fp32
inline float keep_scale(double p) { const float pk = (float)(1.0 - p); return (float)(1.0 / (double)pk); }

inline int vec_for(long long inner, std::initializer_list<const void*> ptrs) {
    int vec = (inner % 4 == 0) ? 4 : 1;
    for (const void* p : ptrs)
        if (p != nullptr && !aligned_to(p, 16)) vec = 1;
    return vec;
}

}  // namespace
}  // namespace uaps

using namespace uaps;

UAPS_API int uaps_feature_noise(const float* x, const float* noise, uint64_t seed, float range, float* y, int B,
                                int64_t chw, cudaStream_t stream) {
    if (x == nullptr || y == nullptr || B <= 0 || chw <= 0) return UAPS_EINVAL;
    if (!aligned_to(x, 4) || !aligned_to(y, 4) || (noise && !aligned_to(noise, 4))) return UAPS_EALIGN;
    const int vec = vec_for(chw, {x, y, noise});
    if (vec == 4)
        feature_noise_kernel<4><<<stream_grid(chw / 4), PT, 0, stream>>>(x, noise, seed, range, y, B, chw);
    else
        feature_noise_kernel<1><<<stream_grid(chw), PT, 0, stream>>>(x, noise, seed, range, y, B, chw);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_dropout(const float* x, const uint8_t* keep, uint64_t seed, double p, float* y, int64_t n,
                          cudaStream_t stream) {
    if (x == nullptr || y == nullptr || n <= 0) return UAPS_EINVAL;
    if (!(p >= 0.0 && p < 1.0)) return UAPS_ERANGE;
    if (!aligned_to(x, 4) || !aligned_to(y, 4)) return UAPS_EALIGN;
    const float scale = keep_scale(p);
    const int vec = vec_for(n, {x, y});
    if (vec == 4)
        dropout_kernel<4><<<stream_grid(n / 4), PT, 0, stream>>>(x, keep, seed, (float)p, scale, y, n);
    else
        dropout_kernel<1><<<stream_grid(n), PT, 0, stream>>>(x, keep, seed, (float)p, scale, y, n);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

namespace {
inline dim3 grid_hw_b(long long nvec, int B) {
    long long want = ceil_div<long long>(nvec, PT);
    long long cap = ceil_div<long long>((long long)device_info().sm_count * kWaves, B);
    if (cap < 1) cap = 1;
    return dim3((unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), (unsigned)B, 1);
}
}  // namespace

UAPS_API int uaps_fdrop_stats(const float* x, int B, int C, int64_t HW, float* attention, uint32_t* smax_enc,
                              cudaStream_t stream) {
    if (x == nullptr || attention == nullptr || smax_enc == nullptr || B <= 0 || C <= 0 || HW <= 0) return UAPS_EINVAL;
    if (B > 65535) return UAPS_ERANGE;
    if (!aligned_to(x, 4) || !aligned_to(attention, 4) || !aligned_to(smax_enc, 4)) return UAPS_EALIGN;
    const int vec = vec_for(HW, {x, attention});
    if (vec == 4) fdrop_stats_kernel<4><<<grid_hw_b(HW / 4, B), PT, 0, stream>>>(x, C, HW, attention, smax_enc);
    else fdrop_stats_kernel<1><<<grid_hw_b(HW, B), PT, 0, stream>>>(x, C, HW, attention, smax_enc);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_fdrop_apply(const float* x, const float* attention, const uint32_t* smax_enc, float u, float* y,
                              int B, int C, int64_t HW, cudaStream_t stream) {
    if (x == nullptr || attention == nullptr || smax_enc == nullptr || y == nullptr || B <= 0 || C <= 0 || HW <= 0)
        return UAPS_EINVAL;
    if (B > 65535) return UAPS_ERANGE;
    if (!aligned_to(x, 4) || !aligned_to(attention, 4) || !aligned_to(y, 4)) return UAPS_EALIGN;
    const int vec = vec_for(HW, {x, attention, y});
    if (vec == 4) fdrop_apply_kernel<4><<<grid_hw_b(HW / 4, B), PT, 0, stream>>>(x, attention, smax_enc, u, y, C, HW);
    else fdrop_apply_kernel<1><<<grid_hw_b(HW, B), PT, 0, stream>>>(x, attention, smax_enc, u, y, C, HW);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_perturb3(const float* x, const float* noise, const uint8_t* keep, uint64_t seed, float noise_range,
                           double p_drop, const float* attention, const uint32_t* smax_enc, float u, float* y_noise,
                           float* y_drop, float* y_fdrop, int B, int C, int64_t HW, cudaStream_t stream) {
    if (x == nullptr || B <= 0 || C <= 0 || HW <= 0) return UAPS_EINVAL;
    if (y_noise == nullptr && y_drop == nullptr && y_fdrop == nullptr) return UAPS_EINVAL;
    if (y_fdrop != nullptr && (attention == nullptr || smax_enc == nullptr)) return UAPS_EINVAL;
    if (B > 65535 || !(p_drop >= 0.0 && p_drop < 1.0)) return UAPS_ERANGE;
    if (!aligned_to(x, 4)) return UAPS_EALIGN;
    const float scale = keep_scale(p_drop);
    const int vec = vec_for(HW, {x, noise, attention, y_noise, y_drop, y_fdrop});
    if (vec == 4)
        perturb3_kernel<4><<<grid_hw_b(HW / 4, B), PT, 0, stream>>>(x, noise, keep, seed, noise_range, (float)p_drop, scale,
                                                                   attention, smax_enc, u, y_noise, y_drop, y_fdrop, C, HW);
    else
        perturb3_kernel<1><<<grid_hw_b(HW, B), PT, 0, stream>>>(x, noise, keep, seed, noise_range, (float)p_drop, scale,
                                                               attention, smax_enc, u, y_noise, y_drop, y_fdrop, C, HW);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

// Backward of uaps_perturb3: the three upstream gradients are read once and folded into one
// dx = g_noise * n + g_noise  +  g_drop * keep / (1 - p)  +  g_fdrop * mask   (any g_* may be NULL).
namespace uaps {
namespace {
template <int VEC>
__global__ void __launch_bounds__(PT) perturb3_bwd_kernel(const float* __restrict__ g_noise, const float* __restrict__ g_drop,
                                                          const float* __restrict__ g_fdrop, const float* __restrict__ noise,
                                                          const uint8_t* __restrict__ keep, uint64_t seed, float range,
                                                          float p, float scale, const float* __restrict__ attention,
                                                          const uint32_t* __restrict__ smax_enc, float u,
                                                          float* __restrict__ dx, int C, long long HW) {
    const int b = blockIdx.y;
    const long long chw = (long long)C * HW;
    const size_t off = (size_t)b * chw;
    const float thr = (g_fdrop != nullptr) ? __fmul_rn(dec_ordered(smax_enc[b]), u) : 0.f;
    const long long nv = HW / VEC;
    for (long long v = (long long)blockIdx.x * PT + threadIdx.x; v < nv; v += (long long)gridDim.x * PT) {
        float m[VEC];
        if (g_fdrop != nullptr) {
            float a[VEC];
            load_vec<VEC>(attention + (size_t)b * HW + v * VEC, a);
#pragma unroll
            for (int i = 0; i < VEC; ++i) m[i] = (a[i] < thr) ? 1.f : 0.f;
        }
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const long long e = (long long)c * HW + v * VEC;
            float o[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = 0.f;
            if (g_noise != nullptr) {
                float gv[VEC], n[VEC];
                load_vec<VEC>(g_noise + off + e, gv);
                if (noise != nullptr) load_vec<VEC>(noise + e, n);
                else {
                    float n4[4];
                    philox_noise4(seed, (uint64_t)e / 4, range, n4);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) n[i] = n4[(e + i) & 3];
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] = __fadd_rn(__fmul_rn(gv[i], n[i]), gv[i]);
            }
            if (g_drop != nullptr) {
                float gv[VEC], k[VEC];
                load_vec<VEC>(g_drop + off + e, gv);
                if (keep != nullptr) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) k[i] = keep[off + e + i] ? 1.f : 0.f;
                } else {
                    float k4[4];
                    philox_keep4(seed, (uint64_t)(off + e) / 4, p, k4);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) k[i] = k4[(off + e + i) & 3];
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] += gv[i] * k[i] * scale;
            }
            if (g_fdrop != nullptr) {
                float gv[VEC];
                load_vec<VEC>(g_fdrop + off + e, gv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] += gv[i] * m[i];
            }
            store_vec<VEC>(dx + off + e, o);
        }
    }
}
}  // namespace
}  // namespace uaps

UAPS_API int uaps_perturb3_bwd(const float* g_noise, const float* g_drop, const float* g_fdrop, const float* noise,
                               const uint8_t* keep, uint64_t seed, float noise_range, double p_drop,
                               const float* attention, const uint32_t* smax_enc, float u, float* dx, int B, int C,
                               int64_t HW, cudaStream_t stream) {
    if (dx == nullptr || B <= 0 || C <= 0 || HW <= 0) return UAPS_EINVAL;
    if (g_noise == nullptr && g_drop == nullptr && g_fdrop == nullptr) return UAPS_EINVAL;
    if (g_fdrop != nullptr && (attention == nullptr || smax_enc == nullptr)) return UAPS_EINVAL;
    if (B > 65535 || !(p_drop >= 0.0 && p_drop < 1.0)) return UAPS_ERANGE;
    const float scale = keep_scale(p_drop);
    const int vec = vec_for(HW, {g_noise, g_drop, g_fdrop, noise, attention, dx});
    if (vec == 4)
        perturb3_bwd_kernel<4><<<grid_hw_b(HW / 4, B), PT, 0, stream>>>(g_noise, g_drop, g_fdrop, noise, keep, seed,
                                                                       noise_range, (float)p_drop, scale, attention,
                                                                       smax_enc, u, dx, C, HW);
    else
        perturb3_bwd_kernel<1><<<grid_hw_b(HW, B), PT, 0, stream>>>(g_noise, g_drop, g_fdrop, noise, keep, seed,
                                                                   noise_range, (float)p_drop, scale, attention,
                                                                   smax_enc, u, dx, C, HW);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
