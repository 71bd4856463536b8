// Gradient all-reduce FUSED with the Adam step, over NVLink peer memory (the reference's DataParallel sums the
// replica gradients onto GPU 0 and re-broadcasts the parameters every forward, UAPS_model.py:13; UAPS_train.py:287-292).
//
// One process per GPU, every rank holds a full replica.  Each rank's flat gradient buffer lives in memory that every
// other rank has mapped (CUDA IPC).  ONE kernel per rank per iteration:
//   barrier A : "my gradients of this epoch are complete" -> every peer's flag box; wait for every peer's flag
//   reduce    : every rank reads ALL ranks' gradients (own from HBM, peers' over NVLink) and adds them in rank order --
//               the same order everywhere, so every rank forms bit-identical sums and applies the identical Adam update
//               to its own replica: no parameter broadcast, replicas stay bit-identical by construction
//   barrier B : "I have finished reading" -> every peer; wait for every peer (nobody may overwrite its gradients -- the
//               next iteration's zeroing -- while a peer is still reading them)
// Versus ncclAllReduce -> Adam: no second pass over the gradients, no NCCL call inside the iteration (the whole
// iteration stays a plain kernel sequence that captures into a CUDA graph), one launch.  Traffic per rank: (W-1) x 14.9 MB
// over NVLink for UNet_UAPS -- 20 us at W = 2, ~150 us at W = 8 against a ~36 ms iteration.
//
// Flags carry a monotonically increasing epoch kept in the rank's own flag box (all ranks call this the same number
// of times, so the epochs agree without any host or NCCL involvement).  Every wait is bounded (UAPS_XCHG_TIMEOUT_MS):
// a dead peer yields a skipped update and a latched status word, never a hung GPU.
#include <cstdlib>
#include "common.cuh"

namespace uaps {
namespace {

constexpr int GT = 256;
constexpr int WMAX = UAPS_XCHG_MAX_RANKS;
// 32-bit words of the flag box (a uaps_xchg_alloc mailbox, zero-initialised)
constexpr int FB_EPOCH = 0, FB_READY = 16, FB_DONE = 48, FB_TICKET = 80, FB_STATUS = 81;

struct ReduceArgs {
    const float4* g[WMAX];       // every rank's gradient buffer as mapped here (own at [rank])
    unsigned* box[WMAX];         // every rank's flag box as mapped here
    int rank, world;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// threads 0..world-1 each wait for one source's flag to reach `epoch`; returns (block-wide) whether anybody timed out
__device__ __forceinline__ int wait_flags(const unsigned* own_box, int base, int world, unsigned epoch, unsigned long long timeout_ns) {
    int bad = 0;
    if ((int)threadIdx.x < world) {
        const unsigned long long t0 = timer_ns();
        while ((int)(ld_flag(own_box + base + threadIdx.x) - epoch) < 0) {       // wrap-safe "flag < epoch"
            if (timer_ns() - t0 > timeout_ns) { bad = 1; break; }
        }
    }
    return __syncthreads_or(bad);
}

__global__ void __launch_bounds__(GT) grad_reduce_adam_kernel(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                              long long n4, const ReduceArgs a, float b1, float b2, float step_size,
                                                              float inv_bc2_sqrt, float eps, float grad_scale,
                                                              UapsStepState* __restrict__ state, const float* __restrict__ guard) {
    unsigned* own = a.box[a.rank];
    const unsigned epoch = ld_flag(own + FB_EPOCH) + 1u;          // advanced by the last CTA at the very end
    bool skip = false;
    if (state != nullptr) {
        step_size = state->adam_step_size;
        inv_bc2_sqrt = state->adam_inv_bc2_sqrt;
        if (guard != nullptr) skip = !(fabsf(*guard) <= 3.0e38f);     // the (global) loss is the same on every rank: all skip together
    }
    // ---- barrier A: gradients complete everywhere ---------------------------------------------------------------
    if (blockIdx.x == 0 && (int)threadIdx.x < a.world) st_flag(a.box[threadIdx.x] + FB_READY + a.rank, epoch);
    const int late = wait_flags(own, FB_READY, a.world, epoch, a.timeout_ns);
    if (late) skip = true;
    // ---- reduce + Adam ----------------------------------------------------------------------------------------------
    if (!skip) {
        auto upd = [&](float& pp, float gg, float& mm, float& vv) {
            gg *= grad_scale;
            mm = fmaf(b1, mm, (1.f - b1) * gg);
            vv = fmaf(b2, vv, (1.f - b2) * gg * gg);
            const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
            pp -= step_size * (mm / denom);
        };
        for (long long i = (long long)blockIdx.x * GT + threadIdx.x; i < n4; i += (long long)gridDim.x * GT) {
            float4 gs[WMAX];
#pragma unroll
            for (int r = 0; r < WMAX; ++r)
                if (r < a.world) gs[r] = (r == a.rank) ? __ldcg(a.g[r] + i) : ld_peer(a.g[r] + i);      // all loads in flight together
            float4 g = gs[0];
#pragma unroll
            for (int r = 1; r < WMAX; ++r)
                if (r < a.world) { g.x += gs[r].x; g.y += gs[r].y; g.z += gs[r].z; g.w += gs[r].w; }  // rank order: identical everywhere
            float4 pp = p[i], mm = m[i], vv = v[i];
            upd(pp.x, g.x, mm.x, vv.x); upd(pp.y, g.y, mm.y, vv.y); upd(pp.z, g.z, mm.z, vv.z); upd(pp.w, g.w, mm.w, vv.w);
            p[i] = pp; m[i] = mm; v[i] = vv;
        }
    }
    // ---- barrier B: everybody has finished reading everybody's gradients ------------------------------------------------
    __shared__ unsigned s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(own + FB_TICKET, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < a.world) st_flag(a.box[threadIdx.x] + FB_DONE + a.rank, epoch);
    const int late2 = wait_flags(own, FB_DONE, a.world, epoch, a.timeout_ns);
    if (threadIdx.x == 0) {
        own[FB_TICKET] = 0;
        if (late || late2) own[FB_STATUS] = epoch;                  // latched: the epoch of the first exchange that timed out
        if (state != nullptr && skip) { state->skipped = 1; state->n_skipped += 1; }
        __threadfence();
        st_flag(own + FB_EPOCH, epoch);
    }
}

}  // namespace
}  // namespace uaps

using namespace uaps;

// A device allocation other processes can map (cudaMalloc, zero-filled): gradient buffers that peers read.  Exported /
// mapped / unmapped with uaps_xchg_export / uaps_xchg_import / uaps_xchg_close like the mailboxes.
UAPS_API int uaps_peer_alloc(void** ptr, size_t bytes) {
    if (ptr == nullptr || bytes == 0) return UAPS_EINVAL;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return (int)e;
}
UAPS_API int uaps_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : UAPS_OK; }

UAPS_API int uaps_grad_reduce_adam(float* p, const float* const* grads, float* m, float* v, int64_t n, void* const* flagboxes,
                                   int rank, int world, int64_t step, float lr, float beta1, float beta2, float eps,
                                   float grad_scale, UapsStepState* state, const float* guard, cudaStream_t stream) {
    if (p == nullptr || grads == nullptr || m == nullptr || v == nullptr || flagboxes == nullptr || n <= 0) return UAPS_EINVAL;
    if (world < 1 || world > WMAX || rank < 0 || rank >= world || (n % 4) != 0) return UAPS_ERANGE;
    if (state == nullptr && (step < 1 || guard != nullptr)) return UAPS_EINVAL;
    if (state != nullptr) step = 1;
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f)) return UAPS_ERANGE;
    if (!aligned_to(p, 16) || !aligned_to(m, 16) || !aligned_to(v, 16)) return UAPS_EALIGN;
    ReduceArgs a{};
    for (int r = 0; r < world; ++r) {
        if (grads[r] == nullptr || flagboxes[r] == nullptr) return UAPS_EINVAL;
        if (!aligned_to(grads[r], 16) || !aligned_to(flagboxes[r], 128)) return UAPS_EALIGN;
        a.g[r] = reinterpret_cast<const float4*>(grads[r]);
        a.box[r] = reinterpret_cast<unsigned*>(flagboxes[r]);
    }
    a.rank = rank; a.world = world;
    const char* tmo = getenv("UAPS_XCHG_TIMEOUT_MS");
    a.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 30000ull) * 1000000ull;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    const long long n4 = n / 4;
    // every CTA waits at barrier A while holding its SM slot: the grid must be co-resident (it is: <= 8 CTAs of 256 threads per SM)
    long long want = ceil_div<long long>(n4, GT), cap = (long long)device_info().sm_count * 4;
    const int grid = (int)(want < cap ? want : cap);
    grad_reduce_adam_kernel<<<grid, GT, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(m),
                                                    reinterpret_cast<float4*>(v), n4, a, beta1, beta2, step_size, inv_bc2_sqrt, eps,
                                                    grad_scale, state, guard);
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
