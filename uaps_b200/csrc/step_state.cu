// uaps_step_begin: the per-iteration host work of the reference's loop moved onto the device, so that one training
// iteration is a static launch sequence (CUDA-graph capturable).  Replaces, per iteration,
//   UAPS_train.py:251        np.random.dirichlet(np.ones(K))                  -> state->mix_w
//   UAPS_unet.py:165         np.random.uniform(0.7, 0.9) per FeatureDropout    -> state->u[]
//   UAPS_train.py:279-280    get_current_consistency_weight(iter_num // 80)    -> state->cw1 / cw2 (utilities/ramps.py:19-26)
//   torch.optim.Adam's step count / bias corrections (:292)                    -> state->adam_*
// and the host-side seed and exchange-epoch counters of uaps_b200 itself.  One thread; everything in fp64 like the
// host code it replaces, rounded to fp32 where the host code's values are rounded when they meet a tensor.
#include "common.cuh"

namespace uaps {
namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr uint32_t kStreamMix = 31, kStreamU = 32;

// 53-bit uniform in (0, 1) from two Philox words
__device__ __forceinline__ double u01_open(uint32_t hi, uint32_t lo) {
    const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void step_begin_kernel(UapsStepState* st, uint64_t seed_rank, uint64_t seed_shared, int K, int n_u, double c1,
                                  double c2, double rampup, int ipe, int n_xchg, float beta1, float beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t it = st->iter;
    st->key_rank = splitmix64(seed_rank ^ splitmix64(it));
    st->key_shared = splitmix64(seed_shared ^ splitmix64(it + 0x51ED270B27ull));
    // Dirichlet(1, ..., 1): K standard exponentials, normalised (what numpy's generator does for alpha = 1)
    double e[8], tot = 0.0;
    for (int k = 0; k < K; k += 2) {
        uint32_t r[4];
        Philox::draw4(seed_shared, it * 8 + (uint64_t)(k / 2), kStreamMix, r);
        e[k] = -log(u01_open(r[0], r[1]));
        if (k + 1 < K) e[k + 1] = -log(u01_open(r[2], r[3]));
    }
    for (int k = 0; k < K; ++k) tot += e[k];
    for (int k = 0; k < 8; ++k) st->mix_w[k] = k < K ? (float)(e[k] / tot) : 0.f;
    for (int i = 0; i < n_u && i < UAPS_STEP_USLOTS; i += 2) {
        uint32_t r[4];
        Philox::draw4(seed_shared, it * 16 + (uint64_t)(i / 2), kStreamU, r);
        st->u[i] = (float)(0.7 + 0.2 * u01_open(r[0], r[1]));
        if (i + 1 < UAPS_STEP_USLOTS) st->u[i + 1] = (float)(0.7 + 0.2 * u01_open(r[2], r[3]));
    }
    // sigmoid_rampup(iter // ipe, rampup) (utilities/ramps.py:19-26)
    double ramp = 1.0;
    if (rampup != 0.0) {
        double cur = (double)(it / (uint64_t)(ipe > 0 ? ipe : 1));
        cur = cur < 0.0 ? 0.0 : (cur > rampup ? rampup : cur);
        const double phase = 1.0 - cur / rampup;
        ramp = exp(-5.0 * phase * phase);
    }
    st->cw1 = (float)(c1 * ramp);
    st->cw2 = (float)(c2 * ramp);
    // Adam: the step about to be taken.  A skipped update (non-finite loss) does not consume a step.
    if (st->skipped) st->skipped = 0; else st->adam_step += 1;
    const double t = (double)st->adam_step;
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    st->adam_step_size = (float)((double)st->lr / bc1);
    st->adam_inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    st->xchg_base = st->xchg_next;
    st->xchg_next += (uint32_t)n_xchg;
    st->iter = it + 1;
}

}  // namespace
}  // namespace uaps

using namespace uaps;

UAPS_API int uaps_step_begin(UapsStepState* state, uint64_t seed_rank, uint64_t seed_shared, int K, int n_u, double consistency1,
                             double consistency2, double rampup_length, int iters_per_ramp_epoch, int n_exchanges, float beta1,
                             float beta2, cudaStream_t stream) {
    if (state == nullptr) return UAPS_EINVAL;
    if (K < 1 || K > 8 || n_u < 0 || n_u > UAPS_STEP_USLOTS || n_exchanges < 0 || iters_per_ramp_epoch < 1) return UAPS_ERANGE;
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f)) return UAPS_ERANGE;
    if (!aligned_to(state, 16)) return UAPS_EALIGN;
    step_begin_kernel<<<1, 32, 0, stream>>>(state, seed_rank, seed_shared, K, n_u, consistency1, consistency2, rampup_length,
                                            iters_per_ramp_epoch, n_exchanges, beta1, beta2);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
