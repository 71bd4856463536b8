// C-ABI entry points of the fused UAPS loss (kernels: fused_loss_impl.cuh, instantiated per K in
// fused_loss_k*.cu).  Host side only validates, packs the by-value kernel argument block and picks
// the vector width; it never allocates, synchronises or retains pointers.
#define UAPS_LOSS_ENTRY
#include <cstdlib>
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
                 const int64_t* labels) {
    if (z == nullptr || B <= 0 || HW <= 0) return UAPS_EINVAL;
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    if (labels == nullptr && mix_w == nullptr) return UAPS_EINVAL;
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
int pick_impl(int avail_vec, int flags, bool pass2) {
    static const int forced = [] { const char* e = getenv("UAPS_LOSS_IMPL"); return e ? atoi(e) : -1; }();
    if (flags & UAPS_LOSS_EXACT) return IMPL_EXACT;
    if (avail_vec < 2) return IMPL_SCALAR;
    if (forced >= 0 && forced <= IMPL_SCALAR) return forced;
    if (avail_vec < 4) return IMPL_VEC2_PF;
    return pass2 ? IMPL_VEC2_PF : IMPL_VEC4_PF;
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
                          int64_t N_global, float cw1, float cw2, float* scalars) {
    int rc = check_common(z, K, B, C, HW, mix_w, labels);
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
    a.labels = labels; a.pseudo = pseudo_out; a.HW = HW; a.B = B;
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + WS_HEADER_BYTES);
    const int impl = pick_impl(pick_vec(z, exp_var_out, K, HW), flags, false);
    int nblocks = 0;
    rc = dispatch_k(K, C, impl, labels != nullptr, false, a, partials, nullptr, nullptr, &nblocks, stream);
    if (rc != UAPS_OK) return rc;
    const int S = sums_count(K, C);
    if (scalars != nullptr)
        loss_fold_finalize_kernel<<<1, 1024, 0, stream>>>(partials, S, (unsigned)nblocks, sums, K, C, (double)N_global, cw1, cw2,
                                                          labels != nullptr, scalars);
    else
        loss_fold_kernel<<<ceil_div(S, 256 / kWarp), 256, 0, stream>>>(partials, S, (unsigned)nblocks, sums);
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
                                     cudaStream_t stream) {
    if (scalars == nullptr || !aligned_to(scalars, 4)) return UAPS_EINVAL;
    return loss_pass1_impl(z, K, B, C, HW, mix_w, labels, workspace, sums, pseudo_out, exp_var_out, flags, stream,
                           (int64_t)B * HW, cw1, cw2, scalars);
}


UAPS_API int uaps_loss_finalize(const double* sums_global, int K, int C, int64_t N_global, float cw1,
                                  float cw2, int supervised, float* scalars, cudaStream_t stream) {
    if (sums_global == nullptr || scalars == nullptr || N_global <= 0) return UAPS_EINVAL;
    if (K < 1 || K > KMAX || C < 2 || C > CMAX) return UAPS_ERANGE;
    if (!aligned_to(sums_global, 8) || !aligned_to(scalars, 4)) return UAPS_EALIGN;
    loss_finalize_kernel<<<1, 32, 0, stream>>>(sums_global, K, C, (double)N_global, cw1, cw2, supervised, scalars);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}

UAPS_API int uaps_loss_pass2(const float* const* z, int K, int B, int C, int64_t HW, const float* mix_w,
                               const int64_t* labels, const float* scalars, const float* grad_out,
                               float* const* dz, int flags, cudaStream_t stream) {
    int rc = check_common(z, K, B, C, HW, mix_w, labels);
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
    a.labels = labels; a.pseudo = nullptr; a.HW = HW; a.B = B;
    const int impl = pick_impl(pick_vec(z, dz, K, HW), flags, true);
    int nblocks = 0;
    return dispatch_k(K, C, impl, labels != nullptr, true, a, nullptr, scalars, grad_out, &nblocks, stream);
}
