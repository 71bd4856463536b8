// C-ABI entry points of the fused UAPS loss (kernels: fused_loss_impl.cuh, instantiated per K in
// fused_loss_k*.cu).  Host side only validates, packs the by-value kernel argument block and picks
// the vector width; it never allocates, synchronises or retains pointers.
#define UAPS_LOSS_ENTRY
#include <cstdlib>
#include <cstring>
#include "fused_loss_impl.cuh"

namespace uaps {
namespace loss {
#define UAPS_DECL_K(KK)                                                                                       \
    extern template int launch_loss_k<KK>(int, int, bool, bool, const LossArgs&, float*, const float*,       \
                                          const float*, int*, cudaStream_t);
UAPS_DECL_K(1) UAPS_DECL_K(2) UAPS_DECL_K(3) UAPS_DECL_K(4) UAPS_DECL_K(5) UAPS_DECL_K(6)
#undef UAPS_DECL_K
}  // namespace loss

namespace {
using namespace loss;

int check_common(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                 const int64_t* labels, const float* wcw_dev) {
    if (z == nullptr || B <= 0 || HW <= 0) return UAPS_EINVAL;
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    if (labels == nullptr && mix_w == nullptr && wcw_dev == nullptr) return UAPS_EINVAL;
    if ((double)B * (double)HW >= 2147483648.0) return UAPS_ERANGE;
    for (int k = 0; k < K; ++k) {
        if (z[k] == nullptr) return UAPS_EINVAL;
        if (!aligned_to(z[k], 4)) return UAPS_EALIGN;
    }
    if (labels != nullptr && !aligned_to(labels, 8)) return UAPS_EALIGN;
    return UAPS_OK;
}

// widest vector width every plane of every tensor supports
int pick_vec(const float* const* z, float* const* o, int K, int64_t HW) {
    int vec = (HW % 4 == 0) ? 4 : (HW % 2 == 0 ? 2 : 1);
    for (int k = 0; k < K; ++k) {
        while (vec > 1 && !aligned_to(z[k], 4 * vec)) vec >>= 1;
        if (o != nullptr && o[k] != nullptr)
            while (vec > 1 && !aligned_to(o[k], 4 * vec)) vec >>= 1;
    }
    return vec;
}


// Kernel variant for a call: pass 1 runs 4 pixels/thread, pass 2 two (both software-pipelined) when
// the planes are 16-byte aligned (HW % 4 == 0); odd shapes take the scalar kernels and UAPS_LOSS_EXACT
// the torch-order ones.  UAPS_LOSS_IMPL=<0..2> overrides the choice (tuning knob, not part of the ABI).
int pick_impl(int avail_vec, int flags, bool pass2, int kc) {
    static const int forced = [] { const char* e = getenv("UAPS_LOSS_IMPL"); return e ? atoi(e) : -1; }();
    static const int forced_p1 = [] { const char* e = getenv("UAPS_LOSS_P1_IMPL"); return e ? atoi(e) : -1; }();   // pass 1 only
    if (flags & UAPS_LOSS_EXACT) return IMPL_EXACT;
    if (avail_vec < 2) return IMPL_SCALAR;
    if (!pass2 && forced_p1 >= 0 && forced_p1 != IMPL_EXACT && forced_p1 <= IMPL_VEC4 && (forced_p1 != IMPL_VEC4 || avail_vec >= 4) &&
        (forced_p1 != IMPL_VEC4_PF || avail_vec >= 4))
        return forced_p1;
    if (forced >= 0 && forced <= IMPL_SCALAR) return forced;
    if (avail_vec < 4) return IMPL_VEC2_PF;
    // pass 2: two pixels per thread at K*C = 16 (registers), four when there are few class planes (C = 2 data sets): with
    // K*C <= 12 the 4-pixel kernel keeps as many loads in flight as the 2-pixel one does at K*C = 16 -- measured on B200:
    // K=4 C=2 32x512x512 pass 2 106.8 -> 96.3 us (0.77 -> 0.85 of HBM peak), K=5 C=2 32x240x640 79.4 -> 75.4 us, K=3 C=4 32x1024x1024 fwd+bwd 0.736 -> 0.783
    static const int p2_vec4_max_kc = [] { const char* e = getenv("UAPS_LOSS_P2_VEC4_MAXKC"); return e ? atoi(e) : 12; }();
    if (pass2) return kc <= p2_vec4_max_kc ? IMPL_VEC4_PF : IMPL_VEC2_PF;
    return IMPL_VEC4_PF;
}

int dispatch_k(int K, int C, int impl, bool sup, bool pass2, const LossArgs& a, float* partials, const float* sc,
               const float* go, int* nblocks, cudaStream_t st) {
    switch (K) {
        case 1: return launch_loss_k<1>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
        case 2: return launch_loss_k<2>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
        case 3: return launch_loss_k<3>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
        case 4: return launch_loss_k<4>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
        case 5: return launch_loss_k<5>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
        case 6: return launch_loss_k<6>(C, impl, sup, pass2, a, partials, sc, go, nblocks, st);
    }
    return UAPS_ERANGE;
}


// ---- multi-GPU: fold + exchange over NVLink peer memory + finalize, one launch ------------------------------
// The path's single exchange step (the <= 70-double partial-sum vector between the passes) done by
// the fold kernel itself instead of fold -> ncclAllReduce -> finalize: every rank stores its folded sums into
// slot [rank] of EVERY rank's mailbox (P2P stores through NVLink/NVSwitch), raises a per-source flag, waits for
// the world's flags in its own mailbox and adds the slots in rank order -- the same order on every rank, so all
// ranks finalize bit-identical scalars.  Flags carry the call's epoch (no resets); two phases alternate so a
// fast rank's next exchange never overwrites slots a slow rank is still reading (a rank can only be one
// exchange ahead: it cannot leave exchange e+1 before every peer has entered it).
constexpr int XCHG_WMAX = UAPS_XCHG_MAX_RANKS;
constexpr int XCHG_SLOT = 128;                                  // doubles per (phase, source) slot, >= sums_count max
// Every double travels as TWO 8-byte words {32 data bits, epoch}: an aligned 8-byte store is delivered atomically, so
// the flag arrives WITH the data (NCCL's "LL" idea) -- no fence + separate flag store + second NVLink hop.
constexpr size_t XCHG_WORDS = (size_t)2 * XCHG_WMAX * 2 * XCHG_SLOT;          // [phase][source][2 * slot] x 8 bytes
constexpr size_t XCHG_STATUS_OFF = XCHG_WORDS * sizeof(unsigned long long);
constexpr size_t XCHG_BYTES = XCHG_STATUS_OFF + 128;
static_assert(sums_count(KMAX, CMAX) <= XCHG_SLOT, "slot too small");

struct ExchangeArgs {
    char* box[XCHG_WMAX];        // mailbox of every rank as mapped into THIS process (own one at [rank])
    int rank, world;
    unsigned epoch;              // 1, 2, 3, ... identical on all ranks for one exchange
    const unsigned* epoch_dev;   // nullable: the epoch used is *epoch_dev + epoch (UapsStepState.xchg_base)
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_word_sys(unsigned long long* p, unsigned data, unsigned flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ void ld_word_sys(const unsigned long long* p, unsigned& data, unsigned& flag) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(data), "=r"(flag) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(1024) loss_fold_exchange_finalize_kernel(const float* __restrict__ partials, int S, unsigned nblocks,
                                                                            double* __restrict__ sums, const ExchangeArgs x, int K,
                                                                            int C, double N, float cw1, float cw2, int supervised,
                                                                            float* __restrict__ sc, int nsc,
                                                                            const float* __restrict__ wcw_dev) {
    __shared__ double s_sums[XCHG_SLOT];
    __shared__ double s_term[KMAX * CMAX];
    __shared__ int s_timed_out;
    pdl_trigger();
    pdl_wait();
    const unsigned epoch = x.epoch + (x.epoch_dev != nullptr ? *x.epoch_dev : 0u);
    if (wcw_dev != nullptr) { cw1 = wcw_dev[UAPS_WCW_CW1]; cw2 = wcw_dev[UAPS_WCW_CW2]; }
    if (threadIdx.x == 0) s_timed_out = 0;
    cta_fold_rows(partials, S, nblocks, s_sums);                 // local fold, as loss_fold_finalize_kernel
    __shared__ unsigned s_words[XCHG_WMAX][2 * XCHG_SLOT];
    const int ph = epoch & 1;
    const int nw = 2 * S;                                        // 32-bit halves of my S doubles
    // scatter: word idx of my sums -> slot [ph][rank][idx] of EVERY mailbox (peer stores travel over NVLink), tagged with the epoch
    for (int t = threadIdx.x; t < x.world * nw; t += blockDim.x) {
        const int p = t / nw, idx = t - p * nw;
        const unsigned bits = reinterpret_cast<const unsigned*>(s_sums)[idx];
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(x.box[p]) +
                                  ((size_t)(ph * XCHG_WMAX + x.rank) * 2 * XCHG_SLOT + idx);
        st_word_sys(dst, bits, epoch);
    }
    // gather: every word of every source in MY mailbox, as soon as its tag shows this epoch (the other phase's words
    // carry epoch - 1, this phase's stale ones epoch - 2).  Bounded spin: a dead peer must not hang the GPU -- on timeout
    // the scalars become NaN and the status word records the epoch.
    int timed_out = 0;
    for (int t = threadIdx.x; t < x.world * nw; t += blockDim.x) {
        const int p = t / nw, idx = t - p * nw;
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(x.box[x.rank]) +
                                        ((size_t)(ph * XCHG_WMAX + p) * 2 * XCHG_SLOT + idx);
        unsigned data, flag;
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            ld_word_sys(src, data, flag);
            if (flag == epoch) break;
            // one thread timing out ends every thread's wait (the time-outs must not add up word after word)
            if (*reinterpret_cast<volatile int*>(&s_timed_out) != 0 || global_timer_ns() - t0 > x.timeout_ns) {
                s_timed_out = 1; timed_out = 1; break;
            }
        }
        if (timed_out) break;
        s_words[p][idx] = data;
    }
    timed_out = __syncthreads_or(timed_out);
    if (timed_out) {
        for (int i = threadIdx.x; i < nsc; i += blockDim.x) sc[i] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(x.box[x.rank] + XCHG_STATUS_OFF) = epoch;
        return;
    }
    if (threadIdx.x < S) {
        double r = 0.0;
        for (int p = 0; p < x.world; ++p)                        // rank order: identical on all ranks
            r += __hiloint2double((int)s_words[p][2 * threadIdx.x + 1], (int)s_words[p][2 * threadIdx.x]);
        s_sums[threadIdx.x] = r;
        sums[threadIdx.x] = r;
    }
    __syncthreads();
    cta_finalize(s_sums, K, C, N, cw1, cw2, supervised, sc, s_term);
}

}  // namespace
}  // namespace uaps

using namespace uaps;
using namespace uaps::loss;

UAPS_API int uaps_loss_sums_count(int K, int C) {
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    return sums_count(K, C);
}
UAPS_API int uaps_loss_scalars_count(int K, int C) {
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    return scalars_count(K, C);
}
UAPS_API size_t uaps_loss_workspace_bytes(int K, int C) {
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return 0;
    return WS_HEADER_BYTES + (size_t)LOSS_MAX_BLOCKS * sums_count(K, C) * sizeof(float);   // partials[S][MAX_BLOCKS]
}

static int loss_pass1_impl(const float* const* z, int K, int B, int C, int64_t HW,
                               const float* mix_w, const int64_t* labels, void* workspace, double* sums,
                               int64_t* pseudo_out, float* const* exp_var_out, int flags, cudaStream_t stream,
                          int64_t N_global, float cw1, float cw2, float* scalars, const ExchangeArgs* xchg = nullptr,
                          const float* wcw_dev = nullptr) {
    int rc = check_common(z, K, B, C, HW, mix_w, labels, wcw_dev);
    if (rc != UAPS_OK) return rc;
    if (workspace == nullptr || sums == nullptr) return UAPS_EINVAL;
    if (!aligned_to(workspace, 16) || !aligned_to(sums, 8)) return UAPS_EALIGN;
    if (pseudo_out != nullptr && !aligned_to(pseudo_out, 8)) return UAPS_EALIGN;
    LossArgs a{};
    for (int k = 0; k < K; ++k) {
        a.z[k] = z[k];
        a.w[k] = mix_w ? mix_w[k] : 0.f;
        a.out[k] = exp_var_out ? exp_var_out[k] : nullptr;
        if (a.out[k] != nullptr) {
            if (!aligned_to(a.out[k], 4)) return UAPS_EALIGN;
            a.write_ev = 1;
        }
    }
    a.labels = labels; a.pseudo = pseudo_out; a.HW = HW; a.B = B; a.w_dev = labels == nullptr ? wcw_dev : nullptr;
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + WS_HEADER_BYTES);
    const int impl = pick_impl(pick_vec(z, exp_var_out, K, HW), flags, false, K * C);
    int nblocks = 0;
    rc = dispatch_k(K, C, impl, labels != nullptr, false, a, partials, nullptr, nullptr, &nblocks, stream);
    if (rc != UAPS_OK) return rc;
    const int S = sums_count(K, C);
    if (xchg != nullptr)
        rc = (int)launch_pdl(loss_fold_exchange_finalize_kernel, dim3(1), dim3(1024), stream, pdl_enabled(), (const float*)partials, S,
                             (unsigned)nblocks, sums, *xchg, K, C, (double)N_global, cw1, cw2, (int)(labels != nullptr), scalars,
                             scalars_count(K, C), wcw_dev);
    else if (scalars != nullptr)
        rc = (int)launch_pdl(loss_fold_finalize_kernel, dim3(1), dim3(1024), stream, pdl_enabled(), (const float*)partials, S,
                             (unsigned)nblocks, sums, K, C, (double)N_global, cw1, cw2, (int)(labels != nullptr), scalars, wcw_dev);
    else
        rc = (int)launch_pdl(loss_fold_kernel, dim3(ceil_div(S, 256 / kWarp)), dim3(256), stream, pdl_enabled(), (const float*)partials,
                             S, (unsigned)nblocks, sums);
    if (rc != 0) return rc;
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_loss_pass1(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                             const int64_t* labels, void* workspace, double* sums, int64_t* pseudo_out,
                             float* const* exp_var_out, int flags, cudaStream_t stream) {
    return loss_pass1_impl(z, K, B, C, HW, mix_w, labels, workspace, sums, pseudo_out, exp_var_out, flags, stream, 0, 0.f, 0.f,
                           nullptr);
}

// pass 1 + fused fold/finalize: the single-rank path (no exchange between the passes), one launch fewer
UAPS_API int uaps_loss_pass1_scalars(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                                     const int64_t* labels, void* workspace, double* sums, int64_t* pseudo_out,
                                     float* const* exp_var_out, int flags, float cw1, float cw2, float* scalars,
                                     const float* wcw_dev, cudaStream_t stream) {
    if (scalars == nullptr || !aligned_to(scalars, 4)) return UAPS_EINVAL;
    if (wcw_dev != nullptr && !aligned_to(wcw_dev, 4)) return UAPS_EALIGN;
    return loss_pass1_impl(z, K, B, C, HW, mix_w, labels, workspace, sums, pseudo_out, exp_var_out, flags, stream,
                           (int64_t)B * HW, cw1, cw2, scalars, nullptr, wcw_dev);
}

// ---- exchange mailboxes (multi-GPU, one process per GPU on one NVLink domain) ---------------------------------
UAPS_API size_t uaps_xchg_mailbox_bytes(void) { return XCHG_BYTES; }

UAPS_API int uaps_xchg_alloc(void** mailbox) {
    if (mailbox == nullptr) return UAPS_EINVAL;
    cudaError_t e = cudaMalloc(mailbox, XCHG_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*mailbox, 0, XCHG_BYTES);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return (int)e;
}
UAPS_API int uaps_xchg_free(void* mailbox) { return mailbox ? (int)cudaFree(mailbox) : UAPS_OK; }

UAPS_API int uaps_xchg_export(void* mailbox, void* handle64) {
    if (mailbox == nullptr || handle64 == nullptr) return UAPS_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), mailbox);
}
UAPS_API int uaps_xchg_import(const void* handle64, void** peer_mailbox) {
    if (handle64 == nullptr || peer_mailbox == nullptr) return UAPS_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) (void)cudaGetLastError();          // not sticky: the caller falls back to NCCL
    return (int)e;
}
UAPS_API int uaps_xchg_close(void* peer_mailbox) { return peer_mailbox ? (int)cudaIpcCloseMemHandle(peer_mailbox) : UAPS_OK; }

UAPS_API int uaps_xchg_status(const void* mailbox, unsigned* status_out, cudaStream_t stream) {
    if (mailbox == nullptr || status_out == nullptr) return UAPS_EINVAL;
    cudaError_t e = cudaMemcpyAsync(status_out, reinterpret_cast<const char*>(mailbox) + XCHG_STATUS_OFF, sizeof(unsigned),
                                    cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    return (int)e;
}

// pass 1 + fold + peer-memory exchange + finalize: the multi-rank counterpart of uaps_loss_pass1_scalars
UAPS_API int uaps_loss_pass1_exchange(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                                      const int64_t* labels, void* workspace, double* sums, int64_t* pseudo_out,
                                      float* const* exp_var_out, int flags, void* const* mailboxes, int rank, int world,
                                      unsigned epoch, int64_t N_global, float cw1, float cw2, float* scalars,
                                      const float* wcw_dev, const uint32_t* epoch_dev, cudaStream_t stream) {
    if (scalars == nullptr || mailboxes == nullptr || N_global <= 0 || epoch == 0) return UAPS_EINVAL;
    if ((wcw_dev != nullptr && !aligned_to(wcw_dev, 4)) || (epoch_dev != nullptr && !aligned_to(epoch_dev, 4))) return UAPS_EALIGN;
    if (world < 1 || world > XCHG_WMAX || rank < 0 || rank >= world) return UAPS_ERANGE;
    if (!aligned_to(scalars, 4)) return UAPS_EALIGN;
    ExchangeArgs x{};
    for (int p = 0; p < world; ++p) {
        if (mailboxes[p] == nullptr) return UAPS_EINVAL;
        if (!aligned_to(mailboxes[p], 128)) return UAPS_EALIGN;
        x.box[p] = reinterpret_cast<char*>(mailboxes[p]);
    }
    x.rank = rank; x.world = world; x.epoch = epoch; x.epoch_dev = epoch_dev;
    // A peer may legitimately be seconds late (first-iteration lazy initialisation, a dataloader stall, rank 0 writing a
    // checkpoint): the default wait is 30 s.  It stays bounded so that a dead peer can never hang the GPU.
    const char* tmo = getenv("UAPS_XCHG_TIMEOUT_MS");
    x.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 30000ull) * 1000000ull;
    return loss_pass1_impl(z, K, B, C, HW, mix_w, labels, workspace, sums, pseudo_out, exp_var_out, flags, stream, N_global, cw1,
                           cw2, scalars, &x, wcw_dev);
}


UAPS_API int uaps_loss_finalize(const double* sums_global, int K, int C, int64_t N_global, float cw1,
                                  float cw2, int supervised, float* scalars, cudaStream_t stream) {
    if (sums_global == nullptr || scalars == nullptr || N_global <= 0) return UAPS_EINVAL;
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    if (!aligned_to(sums_global, 8) || !aligned_to(scalars, 4)) return UAPS_EALIGN;
    const int rc = (int)launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), stream, pdl_enabled(), sums_global, K, C, (double)N_global,
                                   cw1, cw2, supervised, scalars);
    if (rc != 0) return rc;
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_loss_pass2(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                               const int64_t* labels, const float* scalars, const float* grad_out,
                               float* const* dz, int flags, const float* wcw_dev, cudaStream_t stream) {
    int rc = check_common(z, K, B, C, HW, mix_w, labels, wcw_dev);
    if (rc != UAPS_OK) return rc;
    if (scalars == nullptr || grad_out == nullptr || dz == nullptr) return UAPS_EINVAL;
    LossArgs a{};
    for (int k = 0; k < K; ++k) {
        if (dz[k] == nullptr) return UAPS_EINVAL;
        if (!aligned_to(dz[k], 4)) return UAPS_EALIGN;
        a.z[k] = z[k];
        a.w[k] = mix_w ? mix_w[k] : 0.f;
        a.out[k] = dz[k];
    }
    a.labels = labels; a.pseudo = nullptr; a.HW = HW; a.B = B; a.w_dev = labels == nullptr ? wcw_dev : nullptr;
    const int impl = pick_impl(pick_vec(z, dz, K, HW), flags, true, K * C);
    int nblocks = 0;
    return dispatch_k(K, C, impl, labels != nullptr, true, a, nullptr, scalars, grad_out, &nblocks, stream);
}
