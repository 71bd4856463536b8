// Shared device/host helpers for libuaps_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "uaps_b200.h"

#define UAPS_API extern "C" __attribute__((visibility("default")))

#define UAPS_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

namespace uaps {

constexpr int kWarp = 32;

// Device properties are read once per process and are immutable afterwards.
struct DeviceInfo { int sm_count; int cc_major; int cc_minor; };
inline const DeviceInfo& device_info() {
    static const DeviceInfo info = [] {
        DeviceInfo d{148, 10, 0};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
            cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
        }
        return d;
    }();
    return info;
}

template <typename T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch for the ~1150-kernel training iteration ------------------------------------------
// Every kernel of the bf16 path is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// grid_dep_launch() (its successor's CTAs may be scheduled as soon as all of this grid's CTAs have started) followed --
// after any set-up that touches no global memory -- by grid_dep_wait() (block until the predecessor grid has completed
// and flushed).  Launch latency, CTA scheduling and per-CTA set-up of kernel i+1 then overlap the tail of kernel i; in a
// captured iteration the edges become programmatic graph edges.  MEASURED on B200 (round 2): the captured training iteration
// gets SLOWER with it (32.85 -> 33.76 ms: early CTAs of the successor hold shared memory / TMEM while they wait and the
// persistent kernels are sized to fill the machine), batch-1 inference latency improves (0.274 -> 0.255 ms).  So it is
// OFF by default and UAPS_PDL=1 turns it on; the two instructions are no-ops in a kernel launched without the attribute.
// (The three launches of the fused loss keep their own, always-on PDL chain: fused_loss_impl.cuh.)
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_all_enabled() {
    static const bool on = [] { const char* e = getenv("UAPS_PDL"); return e != nullptr && atoi(e) != 0; }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool cooperative,
                            Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl_all_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cooperative) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define UAPS_LAUNCH(kern, grid, block, smem, stream, ...)                                          \
    do {                                                                                           \
        cudaError_t le__ = ::uaps::launch_k(kern, grid, block, smem, stream, false, __VA_ARGS__);  \
        if (le__ != cudaSuccess) return (int)le__;                                                 \
    } while (0)

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- vector global loads/stores -----------------------------------------------------------
template <int VEC> struct VecF;
template <> struct VecF<4> { using type = float4; };
template <> struct VecF<2> { using type = float2; };
template <> struct VecF<1> { using type = float;  };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&out)[VEC]) {
    using V = typename VecF<VEC>::type;
    V v = __ldg(reinterpret_cast<const V*>(p));
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) out[i] = f[i];
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&in)[VEC]) {
    using V = typename VecF<VEC>::type;
    V v;
    float* f = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) f[i] = in[i];
    *reinterpret_cast<V*>(p) = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// order-preserving float <-> uint32 (so atomicMax on unsigned works for signed floats); 0 is below every float
__host__ __device__ __forceinline__ uint32_t enc_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
    uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(b);
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: draw i is a pure function of (seed, i) --
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
        uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
        uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // 4 x 32 random bits for 128-bit counter (idx, stream) under 64-bit key `seed`
    __host__ __device__ static inline void draw4(uint64_t seed, uint64_t idx, uint32_t stream, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream, 0u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) { round(c, k0, k1); k0 += W0; k1 += W1; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // uniform in [0,1) with 24 bits
    __host__ __device__ static inline float u01(uint32_t bits) { return (bits >> 8) * (1.0f / 16777216.0f); }
};

}  // namespace uaps
