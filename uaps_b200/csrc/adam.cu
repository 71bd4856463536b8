// Adam step of UAPS_train.py:112, 292 (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no
// amsgrad) over the FLAT parameter / gradient / moment buffers of one model replica: one streaming kernel for all
// 3.71 M parameters of UNet_UAPS (16 B read + 12 B written per parameter) instead of a multi-tensor launch sequence.
// Same update as torch:  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//                        p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "common.cuh"

namespace uaps {
namespace {

constexpr int AT = 256;

__global__ void __launch_bounds__(AT) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                  float4* __restrict__ v, long long n4, float* __restrict__ pt,
                                                  const float* __restrict__ gt, float* __restrict__ mt, float* __restrict__ vt,
                                                  int tail, float b1, float b2, float step_size, float inv_bc2_sqrt, float eps,
                                                  float grad_scale, UapsStepState* __restrict__ state,
                                                  const float* __restrict__ guard) {
    if (state != nullptr) {                       // device-resident step: uaps_step_begin computed the bias corrections
        step_size = state->adam_step_size;
        inv_bc2_sqrt = state->adam_inv_bc2_sqrt;
        if (guard != nullptr) {
            const float gv = *guard;
            if (!(fabsf(gv) <= 3.0e38f)) {        // NaN / inf loss (e.g. an exchange that timed out): leave p, m, v untouched
                if (blockIdx.x == 0 && threadIdx.x == 0) { state->skipped = 1; state->n_skipped += 1; }
                return;
            }
        }
    }
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = fmaf(b1, mm, (1.f - b1) * gg);                          // exp_avg.lerp_(grad, 1 - beta1)
        vv = fmaf(b2, vv, (1.f - b2) * gg * gg);                     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
        pp -= step_size * (mm / denom);                              // param.addcdiv_(exp_avg, denom, value=-step_size)
    };
    for (long long i = (long long)blockIdx.x * AT + threadIdx.x; i < n4; i += (long long)gridDim.x * AT) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        const float4 gg = __ldg(g + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) upd(pt[threadIdx.x], gt[threadIdx.x], mt[threadIdx.x], vt[threadIdx.x]);
}

}  // namespace
}  // namespace uaps

using namespace uaps;

// p, g, m, v: device fp32 arrays of n elements, 16-byte aligned.  step: 1, 2, 3, ... (bias correction).
// grad_scale multiplies the gradient first (1 for the SUM-all-reduced gradient of a globally normalised loss).
UAPS_API int uaps_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t step, float lr, float beta1,
                            float beta2, float eps, float grad_scale, UapsStepState* state, const float* guard,
                            cudaStream_t stream) {
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || n <= 0) return UAPS_EINVAL;
    if (state == nullptr && (step < 1 || guard != nullptr)) return UAPS_EINVAL;
    if (state != nullptr) step = 1;               // unused: the kernel reads the device state
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f)) return UAPS_ERANGE;
    if (!aligned_to(p, 16) || !aligned_to(g, 16) || !aligned_to(m, 16) || !aligned_to(v, 16)) return UAPS_EALIGN;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    const long long n4 = n / 4;
    const int tail = (int)(n - n4 * 4);
    long long want = ceil_div<long long>(n4 > 0 ? n4 : 1, AT), cap = (long long)device_info().sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    adam_kernel<<<grid, AT, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
                                         reinterpret_cast<float4*>(v), n4, p + n4 * 4, g + n4 * 4, m + n4 * 4, v + n4 * 4, tail, beta1,
                                         beta2, step_size, inv_bc2_sqrt, eps, grad_scale, state, guard);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
