// Fused pseudo-label + KL-uncertainty + uncertainty-weighted CE/Dice loss, forward (pass 1),
// scalar finalize, and backward (pass 2).  Replaces UAPS_train.py:186-189, 223-282 (+ :194-218 in
// supervised mode) and their autograd graph.  See include/uaps_b200.h for the ABI contract and
// DESIGN.md §3 for the math.
//
// Data layout: K logits tensors, each contiguous NCHW fp32.  A thread owns VEC consecutive pixels
// of one image and issues K*C independent 128-bit loads (one per class plane) before any math,
// so a warp has K*C*512 B in flight per iteration; all per-pixel state lives in registers.
//
// Bit-exactness of the pseudo-label: the chain that decides argmax reproduces the op order and
// roundings of torch's CUDA kernels -- max, x-max, expf, sequential sum, IEEE divide
// (cunn_SpatialSoftMaxForward), then separately rounded w*p multiplies and left-associated adds
// (one ATen kernel each in the reference, so no FMA contraction), then first-maximum argmax.
#pragma once
#include "common.cuh"

namespace uaps {
namespace loss {

constexpr int KMAX = UAPS_KMAX;
constexpr int CMAX = UAPS_CMAX;
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 2048;
constexpr int WS_HEADER_BYTES = 256;          // ticket counter lives in the first 4 bytes

__host__ __device__ constexpr int sums_count(int K, int C) { return 3 * K + 2 * K * C + C; }
__host__ __device__ constexpr int scalars_count(int K, int C) { return UAPS_SC_BASE + 4 * K + 2 * K * C; }

struct LossArgs {
    const float* z[KMAX];
    float* out[KMAX];            // pass1: exp_var_out (nullable entries); pass2: dz
    float w[KMAX];
    const int64_t* labels;       // supervised mode
    int64_t* pseudo;             // pass1 optional
    long long HW;
    unsigned groups_per_image;   // HW / VEC
    unsigned ngroups;            // B * groups_per_image
    int write_ev;
};

// ---- per-pixel forward -------------------------------------------------------------------
template <int K, int C>
struct PixelState {
    float p[K][C];   // softmax
    float l[K][C];   // log-softmax
    float q[C];      // mean prediction
    float lq[C];     // log q (as computed; -inf when q == 0)
    float V[K];      // KL(q || p_k) summed over classes
    float E[K];      // exp(-V)
    int y;           // pseudo-label / label
};

template <int C>
__device__ __forceinline__ void softmax_exact(const float (&z)[C], float (&p)[C], float (&l)[C]) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float e[C];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        e[c] = expf(__fsub_rn(z[c], m));
        s = __fadd_rn(s, e[c]);
    }
    const float ls = logf(s);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        p[c] = __fdiv_rn(e[c], s);
        l[c] = __fsub_rn(__fsub_rn(z[c], m), ls);
    }
}

// ---- fast arithmetic (default mode) ---------------------------------------------------------
// One MUFU per transcendental instead of the ~10-20 instruction precise expansions; every value
// stays within ~1e-6 relative of the fp32 reference.  The pseudo-label is still bit-exact: see
// `pixel_forward` -- whenever the fast mix cannot separate the top two classes by more than
// kTieMargin (>> the ~1.5e-6 worst-case distance between the fast and the torch-order value),
// the argmax is re-decided with the exact torch-order chain.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kTieMargin = 2e-5f;

// PLOG: take log(s) with the precise logf.  Pass 1 sets it: the batch means of V_k (differences of
// logs, ~1e-3 when the decoders agree) must not inherit the mantissa-dependent bias of lg2.approx.
template <int C, bool PLOG>
__device__ __forceinline__ void softmax_fast(const float (&z)[C], float (&p)[C], float (&l)[C]) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float t[C], e[C];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        t[c] = (z[c] - m) * kLog2e;           // subtract first: no cancellation error for large |z|
        e[c] = ex2_approx(t[c]);
        s += e[c];
    }
    const float r = rcp_approx(s);
    const float nls = PLOG ? -logf(s) : -lg2_approx(s) * kLn2;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        p[c] = e[c] * r;
        l[c] = fmaf(t[c], kLn2, nls);
    }
}

// torch-order argmax of the Dirichlet mix (UAPS_train.py:251-255), every rounding reproduced
template <int K, int C>
__device__ __forceinline__ int argmax_exact(const float (&p)[K][C], const float (&w)[K]) {
    int y = 0;
    float best = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float mix = __fmul_rn(w[0], p[0][c]);
#pragma unroll
        for (int k = 1; k < K; ++k) mix = __fadd_rn(mix, __fmul_rn(w[k], p[k][c]));
        if (c == 0) best = mix;
        else if (mix > best) { best = mix; y = c; }
    }
    return y;
}

template <int K, int C>
__device__ __noinline__ int argmax_exact_from_logits(const float (&z)[K][C], const float (&w)[K]) {
    float p[K][C], l[C];
#pragma unroll
    for (int k = 0; k < K; ++k) softmax_exact<C>(z[k], p[k], l);
    return argmax_exact<K, C>(p, w);
}

template <int K, int C, bool SUP, bool EXACT, bool PLOG>
__device__ __forceinline__ void pixel_forward(const float (&z)[K][C], const float (&w)[K], int label,
                                              PixelState<K, C>& st) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if constexpr (EXACT) softmax_exact<C>(z[k], st.p[k], st.l[k]);
        else softmax_fast<C, PLOG>(z[k], st.p[k], st.l[k]);
    }
    if constexpr (SUP) {
        st.y = label;
#pragma unroll
        for (int k = 0; k < K; ++k) { st.V[k] = 0.f; st.E[k] = 1.f; }
        return;
    } else {
        if constexpr (EXACT) {
            st.y = argmax_exact<K, C>(st.p, w);
        } else {
            int y = 0;
            float best = 0.f, second = -1.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float mix = w[0] * st.p[0][c];
#pragma unroll
                for (int k = 1; k < K; ++k) mix = fmaf(w[k], st.p[k][c], mix);
                if (c == 0) best = mix;
                else if (mix > best) { second = best; best = mix; y = c; }
                else second = fmaxf(second, mix);
            }
            if (__builtin_expect(!(best - second > kTieMargin), 0)) y = argmax_exact_from_logits<K, C>(z, w);
            st.y = y;
        }
        // mean prediction (:223) and KL maps (:226-236): V_k = sum_c xlogy(q,q) - q * l_kc
        float h = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = st.p[0][c];
#pragma unroll
            for (int k = 1; k < K; ++k) acc = __fadd_rn(acc, st.p[k][c]);
            const float q = acc * (1.0f / K);
            st.q[c] = q;
            st.lq[c] = (EXACT || PLOG) ? logf(q) : lg2_approx(q) * kLn2;
            h += (q == 0.f) ? 0.f : q * st.lq[c];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) d = fmaf(st.q[c], st.l[k][c], d);
            st.V[k] = h - d;
            st.E[k] = EXACT ? expf(-st.V[k]) : ex2_approx(-kLog2e * st.V[k]);
        }
    }
}

// K*C independent vector loads of one pixel group (one per class plane of every decoder)
template <int K, int C, int VEC>
__device__ __forceinline__ void load_group(const LossArgs& a, unsigned g, float (&zv)[K][C][VEC]) {
    const unsigned b = g / a.groups_per_image;
    const unsigned hw = (g - b * a.groups_per_image) * VEC;
    const size_t base = (size_t)b * C * a.HW + hw;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) load_vec<VEC>(a.z[k] + base + (size_t)c * a.HW, zv[k][c]);
}

// ---- pass 1 --------------------------------------------------------------------------------
// running sums of one thread, flat in the order of `sums` (see uaps_loss_sums_count); every index
// below is a compile-time constant after unrolling, so the array lives in registers
template <int K, int C>
struct AccIdx {
    static constexpr int CE = 0, E = K, V = 2 * K, I = 3 * K, P = 3 * K + K * C, T = 3 * K + 2 * K * C;
    static constexpr int S = 3 * K + 2 * K * C + C;
};

template <int K, int C, int VEC, bool SUP, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS, 1)
loss_pass1_kernel(const LossArgs a, unsigned* __restrict__ ticket, float* __restrict__ partials,
                  double* __restrict__ sums) {
    constexpr int S = sums_count(K, C);
    __shared__ float s_red[LOSS_THREADS / kWarp][S];
    __shared__ bool s_last;

    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = a.w[k];

    using AI = AccIdx<K, C>;
    float acc[S];
#pragma unroll
    for (int i = 0; i < S; ++i) acc[i] = 0.f;

    // register double buffering: the next group's K*C loads are in flight while this one is computed
    // (the kernel runs at one CTA per SM, so latency is hidden by prefetch depth, not occupancy)
    const unsigned stride = gridDim.x * LOSS_THREADS;
    unsigned g = blockIdx.x * LOSS_THREADS + threadIdx.x;
    float zn[K][C][VEC];
    if (g < a.ngroups) load_group<K, C, VEC>(a, g, zn);
    for (; g < a.ngroups; g += stride) {
        const unsigned b = g / a.groups_per_image;
        const unsigned hw = (g - b * a.groups_per_image) * VEC;
        const size_t base = (size_t)b * C * a.HW + hw;
        float zv[K][C][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < VEC; ++j) zv[k][c][j] = zn[k][c][j];
        if (g + stride < a.ngroups) load_group<K, C, VEC>(a, g + stride, zn);
        long long lab[VEC];
        if constexpr (SUP) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) lab[j] = __ldg(a.labels + (size_t)b * a.HW + hw + j);
        }
        float ev[K][VEC];
        long long yv[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float z[K][C];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) z[k][c] = zv[k][c][j];
            PixelState<K, C> st;
            pixel_forward<K, C, SUP, EXACT, true>(z, w, SUP ? (int)lab[j] : 0, st);
            yv[j] = st.y;
            // one-hot of the label as arithmetic masks (keeps p/l in registers: no select chains
            // that the compiler would turn into a local-memory indexed load)
            float oh[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                oh[c] = (st.y == c) ? 1.f : 0.f;
                acc[AI::T + c] += oh[c];
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                ev[k][j] = st.E[k];
                acc[AI::E + k] += st.E[k];
                acc[AI::V + k] += st.V[k];
                float ly = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    ly = fmaf(oh[c], st.l[k][c], ly);
                    acc[AI::I + k * C + c] = fmaf(oh[c], st.p[k][c], acc[AI::I + k * C + c]);
                    acc[AI::P + k * C + c] += st.p[k][c];
                }
                acc[AI::CE + k] -= ly;
            }
        }
        if (a.pseudo != nullptr) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) a.pseudo[(size_t)b * a.HW + hw + j] = yv[j];
        }
        if (a.write_ev) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (a.out[k] != nullptr) store_vec<VEC>(a.out[k] + (size_t)b * a.HW + hw, ev[k]);
        }
    }

    // warp shuffle -> shared -> per-block partial
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const float r = warp_sum(acc[i]);
        if (lane == 0) s_red[warp][i] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += LOSS_THREADS) {
        float r = 0.f;
#pragma unroll
        for (int wv = 0; wv < LOSS_THREADS / kWarp; ++wv) r += s_red[wv][i];
        partials[(size_t)blockIdx.x * S + i] = r;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: fold all per-block partials in fp64, fixed order -> deterministic
    for (int i = warp; i < S; i += LOSS_THREADS / kWarp) {
        double r = 0.0;
        for (unsigned blk = lane; blk < gridDim.x; blk += kWarp)
            r += (double)__ldcg(partials + (size_t)blk * S + i);
        r = warp_sum(r);
        if (lane == 0) sums[i] = r;
    }
    if (threadIdx.x == 0) *ticket = 0u;      // ready for the next call
}

// ---- finalize (compiled only into the entry-point unit) -----------------------------------------
#ifdef UAPS_LOSS_ENTRY
__global__ void loss_finalize_kernel(const double* __restrict__ sums, int K, int C, double N,
                                     float cw1, float cw2, int supervised, float* __restrict__ sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* sCE = sums;
    const double* sE = sums + K;
    const double* sV = sums + 2 * K;
    const double* sI = sums + 3 * K;
    const double* sP = sums + 3 * K + K * C;
    const double* sT = sums + 3 * K + 2 * K * C;
    float* ps = sc + UAPS_SC_BASE;
    float* Eb = ps + K;
    float* CE = Eb + K;
    float* DI = CE + K;
    float* Ikc = DI + K;
    float* Card = Ikc + K * C;
    double ps_loss = 0.0, unc = 0.0;
    for (int k = 0; k < K; ++k) {
        const double ce = sCE[k] / N;
        double d = 0.0;
        for (int c = 0; c < C; ++c) {
            const double card = sP[k * C + c] + sT[c];
            d += 2.0 * sI[k * C + c] / (card + 1e-7);                 // pytorch_losses.py:88
            Ikc[k * C + c] = (float)sI[k * C + c];
            Card[k * C + c] = (float)card;
        }
        const double dice = 1.0 - d / C;
        const double psk = 0.5 * (ce + dice);                          // UAPS_train.py:259-262
        const double eb = supervised ? 1.0 : sE[k] / N;
        ps[k] = (float)psk; Eb[k] = (float)eb; CE[k] = (float)ce; DI[k] = (float)dice;
        ps_loss += psk * eb;                                           // :265-268 (scalar x mean(E))
        unc += sV[k] / N;
    }
    ps_loss /= K;                                                       // :277
    unc /= K;                                                           // :241-243
    sc[UAPS_SC_PS_LOSS] = (float)ps_loss;
    sc[UAPS_SC_L_UNCERT] = supervised ? 0.f : (float)unc;
    sc[UAPS_SC_LOSS_U] = supervised ? (float)ps_loss : (float)((double)cw1 * ps_loss + (double)cw2 * unc);
    sc[UAPS_SC_CW1] = supervised ? 1.f : cw1;
    sc[UAPS_SC_CW2] = supervised ? 0.f : cw2;
    sc[UAPS_SC_INV_N] = (float)(1.0 / N);
    double mce = 0.0, mdi = 0.0;
    for (int k = 0; k < K; ++k) { mce += CE[k]; mdi += DI[k]; }
    sc[UAPS_SC_MEAN_CE] = (float)(mce / K);                             // :216 total_loss_ce
    sc[UAPS_SC_MEAN_DICE] = (float)(mdi / K);                           // :217 total_loss_dice
}

#endif  // UAPS_LOSS_ENTRY

// ---- pass 2 --------------------------------------------------------------------------------
template <int K, int C>
struct GradConsts {
    float psk[K];        // ps_k
    float cce[K];        // (d loss / d CE_k) / N
    float A[K][C];       // -(d loss / d Dice_k) * (2/C) / (Card + eps)
    float Bc[K][C];      // +(d loss / d Dice_k) * (2/C) * I / (Card + eps)^2
    float lam1, lam2, invNK;
};

template <int K, int C, int VEC, bool SUP, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS, 1)
loss_pass2_kernel(const LossArgs a, const float* __restrict__ sc, const float* __restrict__ grad_out) {
    __shared__ GradConsts<K, C> gc;
    if (threadIdx.x == 0) {
        // upstream gradient vector, laid out like scalars[] (only the five differentiable slots are read)
        const float g_loss = grad_out[UAPS_SC_LOSS_U], g_ps = grad_out[UAPS_SC_PS_LOSS], g_unc = grad_out[UAPS_SC_L_UNCERT];
        const float g_ce = grad_out[UAPS_SC_MEAN_CE], g_dice = grad_out[UAPS_SC_MEAN_DICE];
        const float lam1 = g_loss * sc[UAPS_SC_CW1] + g_ps;
        const float lam2 = g_loss * sc[UAPS_SC_CW2] + g_unc;
        const float invN = sc[UAPS_SC_INV_N];
        const float* ps = sc + UAPS_SC_BASE;
        const float* Eb = ps + K;
        const float* Ikc = ps + 4 * K;
        const float* Card = Ikc + K * C;
        gc.lam1 = lam1; gc.lam2 = lam2; gc.invNK = invN / K;
        for (int k = 0; k < K; ++k) {
            gc.psk[k] = ps[k];
            const float half = lam1 * Eb[k] / (2.f * K);           // d loss / d ps_k * 0.5
            const float dce = half + g_ce / K;                     // d loss / d CE_k
            const float ddice = half + g_dice / K;                 // d loss / d Dice_k
            gc.cce[k] = dce * invN;
            for (int c = 0; c < C; ++c) {
                const float den = Card[k * C + c] + 1e-7f;
                gc.A[k][c] = -ddice * (2.f / C) / den;
                gc.Bc[k][c] = ddice * (2.f / C) * Ikc[k * C + c] / (den * den);
            }
        }
    }
    __syncthreads();

    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = a.w[k];

    // register double buffering: the next group's K*C loads are in flight while this one is computed
    // (the kernel runs at one CTA per SM, so latency is hidden by prefetch depth, not occupancy)
    const unsigned stride = gridDim.x * LOSS_THREADS;
    unsigned g = blockIdx.x * LOSS_THREADS + threadIdx.x;
    float zn[K][C][VEC];
    if (g < a.ngroups) load_group<K, C, VEC>(a, g, zn);
    for (; g < a.ngroups; g += stride) {
        const unsigned b = g / a.groups_per_image;
        const unsigned hw = (g - b * a.groups_per_image) * VEC;
        const size_t base = (size_t)b * C * a.HW + hw;
        float zv[K][C][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < VEC; ++j) zv[k][c][j] = zn[k][c][j];
        if (g + stride < a.ngroups) load_group<K, C, VEC>(a, g + stride, zn);
        long long lab[VEC];
        if constexpr (SUP) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) lab[j] = __ldg(a.labels + (size_t)b * a.HW + hw + j);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float z[K][C];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) z[k][c] = zv[k][c][j];
            PixelState<K, C> st;
            pixel_forward<K, C, SUP, EXACT, false>(z, w, SUP ? (int)lab[j] : 0, st);

            float gk[K];
            float Gq[C];
#pragma unroll
            for (int c = 0; c < C; ++c) Gq[c] = 0.f;
            if constexpr (!SUP) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    gk[k] = (gc.lam2 - gc.lam1 * gc.psk[k] * st.E[k]) * gc.invNK;
#pragma unroll
                    for (int c = 0; c < C; ++c) Gq[c] += gk[k] * (st.lq[c] + 1.f - st.l[k][c]);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) Gq[c] *= (1.0f / K);
            }
            float oh[C];
#pragma unroll
            for (int c = 0; c < C; ++c) oh[c] = (st.y == c) ? 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float Gl[C], Gp[C];
                float sGl = 0.f, dot = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    float gl = -gc.cce[k] * oh[c];
                    if constexpr (!SUP) gl = fmaf(-gk[k], st.q[c], gl);
                    const float gp = fmaf(gc.A[k][c], oh[c], Gq[c] + gc.Bc[k][c]);
                    Gl[c] = gl; Gp[c] = gp;
                    sGl += gl;
                    dot = fmaf(st.p[k][c], gp, dot);
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
                    zv[k][c][j] = fmaf(st.p[k][c], (Gp[c] - dot) - sGl, Gl[c]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) store_vec<VEC>(a.out[k] + base + (size_t)c * a.HW, zv[k][c]);
    }
}


// ---- per-K launchers (one translation unit per K keeps nvcc parallel and compile time bounded) ---
// widest per-thread pixel vector that keeps the K*C*VEC logits in registers without spilling
__host__ __device__ constexpr int max_vec(int K, int C) { return K * C <= 24 ? 4 : (K * C <= 48 ? 2 : 1); }

template <typename Kern>
inline int grid_for(Kern kern, unsigned ngroups) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LOSS_THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
    long long want = ceil_div<long long>(ngroups, LOSS_THREADS);
    long long cap = (long long)device_info().sm_count * occ;
    if (cap > LOSS_MAX_BLOCKS) cap = LOSS_MAX_BLOCKS;
    return (int)(want < cap ? want : cap);
}

template <int K, int C, int VEC, bool SUP, bool EXACT>
inline int launch_pass1(const LossArgs& a, unsigned* ticket, float* partials, double* sums, cudaStream_t st) {
    auto kern = loss_pass1_kernel<K, C, VEC, SUP, EXACT>;
    kern<<<grid_for(kern, a.ngroups), LOSS_THREADS, 0, st>>>(a, ticket, partials, sums);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
template <int K, int C, int VEC, bool SUP, bool EXACT>
inline int launch_pass2(const LossArgs& a, const float* sc, const float* go, cudaStream_t st) {
    auto kern = loss_pass2_kernel<K, C, VEC, SUP, EXACT>;
    kern<<<grid_for(kern, a.ngroups), LOSS_THREADS, 0, st>>>(a, sc, go);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

// variants per (K, C): fast math at the widest vector, fast math scalar (odd HW / unaligned), and
// the exact torch-order arithmetic (validation mode, scalar only)
#define UAPS_LOSS_CASE(CC)                                                                          \
    case CC: {                                                                                      \
        constexpr int VM = max_vec(K, CC);                                                          \
        if (exact) return sup ? CALL(K, CC, 1, true, true) : CALL(K, CC, 1, false, true);           \
        if (vec == VM && VM > 1)                                                                    \
            return sup ? CALL(K, CC, VM, true, false) : CALL(K, CC, VM, false, false);              \
        return sup ? CALL(K, CC, 1, true, false) : CALL(K, CC, 1, false, false);                    \
    }

template <int K>
int launch_pass1_k(int C, int vec, bool sup, bool exact, const LossArgs& a, unsigned* ticket, float* partials,
                   double* sums, cudaStream_t st) {
#define CALL(KK, CC, VV, SS, EE) launch_pass1<KK, CC, VV, SS, EE>(a, ticket, partials, sums, st)
    switch (C) {
        UAPS_LOSS_CASE(2) UAPS_LOSS_CASE(3) UAPS_LOSS_CASE(4) UAPS_LOSS_CASE(5)
        UAPS_LOSS_CASE(6) UAPS_LOSS_CASE(7) UAPS_LOSS_CASE(8)
    }
#undef CALL
    return UAPS_ERANGE;
}
template <int K>
int launch_pass2_k(int C, int vec, bool sup, bool exact, const LossArgs& a, const float* sc, const float* go,
                   cudaStream_t st) {
#define CALL(KK, CC, VV, SS, EE) launch_pass2<KK, CC, VV, SS, EE>(a, sc, go, st)
    switch (C) {
        UAPS_LOSS_CASE(2) UAPS_LOSS_CASE(3) UAPS_LOSS_CASE(4) UAPS_LOSS_CASE(5)
        UAPS_LOSS_CASE(6) UAPS_LOSS_CASE(7) UAPS_LOSS_CASE(8)
    }
#undef CALL
    return UAPS_ERANGE;
}
#undef UAPS_LOSS_CASE

}  // namespace loss
}  // namespace uaps
