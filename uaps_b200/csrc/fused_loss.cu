// C-ABI entry points of the fused UAPS loss (kernels: fused_loss_impl.cuh, instantiated per K in
// fused_loss_k*.cu).  Host side only validates, packs the by-value kernel argument block and picks
// the vector width; it never allocates, synchronises or retains pointers.
#define UAPS_LOSS_ENTRY
#include "fused_loss_impl.cuh"

namespace uaps {
namespace loss {
#define UAPS_DECL_K(KK)                                                                                   \
    extern template int launch_pass1_k<KK>(int, int, bool, bool, const LossArgs&, unsigned*, float*, double*,   \
                                           cudaStream_t);                                                 \
    extern template int launch_pass2_k<KK>(int, int, bool, bool, const LossArgs&, const float*, const float*,   \
                                           cudaStream_t);
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


// vector width actually used: the widest the (K,C) instantiation has, if shape/alignment allow it
int final_vec(int K, int C, int avail) {
    const int vm = max_vec(K, C);
    return (avail >= vm && vm > 1) ? vm : 1;
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
    return WS_HEADER_BYTES + (size_t)LOSS_MAX_BLOCKS * sums_count(K, C) * sizeof(float);
}

UAPS_API int uaps_loss_pass1(const float* const* z, int K, int B, int C, int64_t HW,
                               const float* mix_w, const int64_t* labels, void* workspace, double* sums,
                               int64_t* pseudo_out, float* const* exp_var_out, int flags, cudaStream_t stream) {
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
    const int vec = (flags & UAPS_LOSS_EXACT) ? 1 : final_vec(K, C, pick_vec(z, exp_var_out, K, HW));
    a.labels = labels; a.pseudo = pseudo_out; a.HW = HW;
    a.groups_per_image = (unsigned)(HW / vec);
    a.ngroups = (unsigned)B * a.groups_per_image;
    unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + WS_HEADER_BYTES);
    const bool sup = labels != nullptr;
    const bool exact = (flags & UAPS_LOSS_EXACT) != 0;
    switch (K) {
        case 1: return launch_pass1_k<1>(C, vec, sup, exact, a, ticket, partials, sums, stream);
        case 2: return launch_pass1_k<2>(C, vec, sup, exact, a, ticket, partials, sums, stream);
        case 3: return launch_pass1_k<3>(C, vec, sup, exact, a, ticket, partials, sums, stream);
        case 4: return launch_pass1_k<4>(C, vec, sup, exact, a, ticket, partials, sums, stream);
        case 5: return launch_pass1_k<5>(C, vec, sup, exact, a, ticket, partials, sums, stream);
        case 6: return launch_pass1_k<6>(C, vec, sup, exact, a, ticket, partials, sums, stream);
    }
    return UAPS_ERANGE;
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
    const int vec = (flags & UAPS_LOSS_EXACT) ? 1 : final_vec(K, C, pick_vec(z, dz, K, HW));
    a.labels = labels; a.pseudo = nullptr; a.HW = HW;
    a.groups_per_image = (unsigned)(HW / vec);
    a.ngroups = (unsigned)B * a.groups_per_image;
    const bool sup = labels != nullptr;
    const bool exact = (flags & UAPS_LOSS_EXACT) != 0;
    switch (K) {
        case 1: return launch_pass2_k<1>(C, vec, sup, exact, a, scalars, grad_out, stream);
        case 2: return launch_pass2_k<2>(C, vec, sup, exact, a, scalars, grad_out, stream);
        case 3: return launch_pass2_k<3>(C, vec, sup, exact, a, scalars, grad_out, stream);
        case 4: return launch_pass2_k<4>(C, vec, sup, exact, a, scalars, grad_out, stream);
        case 5: return launch_pass2_k<5>(C, vec, sup, exact, a, scalars, grad_out, stream);
        case 6: return launch_pass2_k<6>(C, vec, sup, exact, a, scalars, grad_out, stream);
    }
    return UAPS_ERANGE;
}
