// Weight gradient of the 3x3 / 1x1 convolutions on tcgen05 (UAPS_train.py:287 `loss.backward()` through
// the nn.Conv2d layers of utilities/UAPS_unet.py:36,41,73,138):
//     dW[co][ci][r][s] = sum over pixels p of dY[p][co] * X[p + (r-1, s-1)][ci]
// GEMM per vertical tap r:  D_r[(s, ci)][co] = sum_pixels X[p + (r-1, s-1)][ci] * dY[p][co]
//   M = (horizontal tap s) x (16/32-wide chunk of input channels), N = output channels, K = pixels.
// Both operands are channels-last, so the reduction dimension (pixels) is the OUTER one: they are fed to the
// tensor core as MN-major operands straight from the TMA boxes, no transpose pass:
//   A = ONE X halo tile [(16+2) x (8+2) px][n_chunk ch]: M-major.  Its M-group stride (LBO) is ONE PIXEL, so M group
//       s is the same tile shifted by s pixels -- the three horizontal taps ride in the M dimension, which costs
//       nothing (an M=128 instruction takes N/2 cycles whatever M holds; groups beyond 3 alias further shifts and
//       their accumulator lanes are never read).  8-pixel K atoms are one image row (10 pixels) apart (SBO).
//   B = dY tile [128 px][min(Cout,64) ch] (x 2 boxes for Cout > 64): N-major, N = Cout of this CTA (<= 128).
// Putting Cout on N and the taps on M makes the tensor time 12 * Cout cycles per 128-pixel tile (192 for the
// 16-channel layers) instead of 576+ with the roles swapped.  What then bounds the small-channel layers is the SS-mode
// operand FETCH (every MMA reads all its M rows from shared memory), so:
//   * M = 64 instead of 128 whenever taps x channel-chunk rows fit (3 x 16 = 48): half the A bytes per instruction;
//   * for Cout <= 64 the three VERTICAL taps ride in N as well ("fused_r"): with p' = p + (r-1, 0),
//         dW[r][s][ci][co] = sum_p' X[p' + (0, s-1)][ci] * dY[p' - (r-1, 0)][co],
//     the X box keeps only its column halo, the dY box gets a row halo, and N group g = 2 - r of the N-major dY descriptor is
//     the dY tile shifted by g rows (LBO = one box row): one MMA per 16-pixel k-step (N = 3 * Cout) instead of three.
// The accumulators (ks x Cout fp32 columns) stay in TMEM across all pixel tiles of the CTA (split-K over CTAs), then are
// added to dW with fp32 reductions.  The TMA -> MMA ring is 8 / 6 / 3 stages deep for 16 / 32 / wider channel chunks.
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

namespace uaps {
namespace wgrad {
using namespace uaps::tc;

constexpr int TILE_H = 16, TILE_W = 8, TILE_M = 128;
constexpr int THREADS = 128;
constexpr int MAX_STAGES = 8;

struct WgradArgs {
    int B, H, W;
    int cout, cin_total, ci_offset;
    int ks, n_chunk;
    int tiles_x, tiles_y, tiles_total, tiles_per_cta;
    int a_ch;                                      // channels per dY box: 16 / 32 / 64
    int n_co;                                      // output channels per CTA (MMA N): multiple of 16, <= 128
    int two_boxes;                                 // cout > 64: second 64-channel box carries data
    float* dw;
    int m_rows;                                    // MMA M: 64 when the taps x channel-chunk rows fit (halves the A-operand fetch), else 128
    int stages;                                    // depth of the TMA -> MMA ring (3..8)
    int fused_r;                                   // 1: the vertical taps ride in N (dY box with a row halo), one MMA per k-step
    // deterministic split-K (workspace given): every CTA stores its partial accumulators, the grid meets at a barrier and
    // each CTA then folds a slice of the output over all splits in split order and adds it to dW -- no fp32 atomics
    unsigned* ws_counters;                         // [0] arrivals, [1] departures (self-resetting; zero before the first use)
    float* ws_partials;                            // [split][group][r block][row][n_co]
    int groups;                                    // n_chunks * m_tiles
};
constexpr size_t WS_HEADER = 256;

// MN-major descriptors (cute::UMMA canonical forms, units of 16 bytes):
//   SW128: ((8,n),(8,k)):((1,LBO),(8,SBO))   SW64: ((4,n),(8,k)):((1,LBO),(4,SBO))   SW32: ((2,n),(8,k)):((1,LBO),(2,SBO))
// k rows are `span` bytes apart, 8 of them form an atom; SBO = distance between K atoms, LBO = between MN groups.
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t span, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t layout = span == 128 ? 2 : (span == 64 ? 4 : 6);
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}
// D = F32, A = B = BF16, both MN-major (bits 15, 16), N >> 3, M >> 4
__device__ __forceinline__ uint32_t idesc_mn(int n, int m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ WgradArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], done_bar;
    const int STAGES = a.stages;
    __shared__ uint32_t tmem_base_smem;
    grid_dep_launch();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int halo = a.ks - 1;
    const int box_w = TILE_W + halo;
    const int row_a = a.a_ch * 2;                                     // 32 / 64 / 128 bytes per pixel row of A
    const int row_b = a.n_chunk * 2;                                  // 32 or 64 bytes per pixel row of B
    // fused_r: the dY box carries the row halo ([16+2][8] pixels) and the X box only the column halo ([16][8+2])
    const int a_box = (a.fused_r ? (TILE_H + halo) * TILE_W : TILE_M) * row_a;
    const int a_bytes = (a.two_boxes ? 2 : 1) * a_box;
    const int b_bytes = (TILE_H + (a.fused_r ? 0 : halo)) * box_w * row_b;
    const int n_issuers = a.fused_r ? 1 : a.ks;
    const int stage_bytes = (a_bytes + b_bytes + 1023) & ~1023;
    // stride between the 128 / a_ch M groups: the second real box, or (fewer channels than M) one atom of the
    // same tile -- in bounds, finite, and its accumulator lanes are ignored
    const uint32_t lbo_dy = a.two_boxes ? (uint32_t)a_box : (uint32_t)(8 * row_a);   // N groups of dY (unused when one group)
    const int n_co = a.n_co;                                          // output channels of this CTA (N of the MMA)

    const int split = blockIdx.x, nc = blockIdx.y, mt = blockIdx.z;
    const int tile_lo = split * a.tiles_per_cta;
    const int tile_hi = min(a.tiles_total, tile_lo + a.tiles_per_cta);
    const int ntiles = tile_hi - tile_lo;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.ks * n_co) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        // the ks vertical taps are issued by ks different warps (warps 1..ks): each commits once per stage / at the end
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, n_issuers); }
        mbar_init(&done_bar, n_issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    grid_dep_wait();                    // the set-up above touched no global memory: it overlapped the predecessor's tail

    if (ntiles > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int it = 0; it < ntiles; ++it) {
                    int t = tile_lo + it;
                    const int tx = t % a.tiles_x; t /= a.tiles_x;
                    const int ty = t % a.tiles_y; t /= a.tiles_y;
                    const int x0 = tx * TILE_W, y0 = ty * TILE_H, n_img = t;
                    const int st = it % STAGES;
                    mbar_wait(empty_bar + st, ((it / STAGES) & 1) ^ 1);
                    unsigned char* sa = smem + (size_t)st * stage_bytes;
                    mbar_expect_tx(full_bar + st, a_bytes + b_bytes);
                    const int ydy = a.fused_r ? y0 - halo / 2 : y0, yx = a.fused_r ? y0 : y0 - halo / 2;
                    tma_load_4d(sa, &map_dy, mt * 128, x0, ydy, n_img, full_bar + st);
                    if (a.two_boxes) tma_load_4d(sa + a_box, &map_dy, mt * 128 + 64, x0, y0, n_img, full_bar + st);
                    tma_load_4d(sa + a_bytes, &map_x, nc * a.n_chunk, x0 - halo / 2, yx, n_img, full_bar + st);
                }
            }
            __syncwarp();
        } else if (warp == 1 && a.fused_r) {
            if (lane == 0) {
                // All nine taps in ONE instruction per 16-pixel k-step.  Substituting p' = p + (r-1, 0):
                //   dW[r][s][ci][co] = sum_p' X[p' + (0, s-1)][ci] * dY[p' - (r-1, 0)][co]
                // so with K = the tile's pixels p', M group s is the X tile shifted by s pixels (LBO = one pixel) and N
                // group g = 2 - r is the dY tile shifted by g rows (LBO = one box row): N = 3 * Cout.  Versus one MMA per r
                // this reads a third of the A operand (SS-mode operand fetch from shared memory is what bounds the
                // small-channel layers, profiles/r01_conv_loader_experiments.txt).
                const uint32_t idesc = idesc_mn(a.ks * n_co, a.m_rows);
                for (int it = 0; it < ntiles; ++it) {
                    const int st = it % STAGES;
                    mbar_wait(full_bar + st, (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sdy = smem_u32(smem + (size_t)st * stage_bytes);
                    const uint32_t sx = sdy + a_bytes;
                    const uint64_t xd0 = desc_mn(sx, row_b, row_b, box_w * row_b);            // A: X, M groups one pixel apart
                    const uint64_t yd0 = desc_mn(sdy, row_a, TILE_W * row_a, TILE_W * row_a);  // B: dY, N groups one row apart
#pragma unroll
                    for (int kk = 0; kk < TILE_M / 16; ++kk) {                    // 16 pixels (two image rows of the tile) per MMA
                        const uint64_t xd = xd0 + (uint64_t)(((2 * kk * box_w) * row_b) >> 4);
                        const uint64_t yd = yd0 + (uint64_t)((kk * 16 * row_a) >> 4);
                        umma_bf16(tmem_d, xd, yd, idesc, (it | kk) != 0);
                    }
                    umma_commit(empty_bar + st);
                }
                umma_commit(&done_bar);
            }
            __syncwarp();
        } else if (warp <= a.ks && !a.fused_r) {
            if (lane == 0) {
                // The three horizontal taps of a row r are ONE instruction: the N-major B descriptor's group
                // stride (LBO) is one pixel, so N group s is the same tile shifted by s pixels (N = 3 * n_chunk,
                // accumulator columns [s][ci]).  And the three vertical taps r are issued by three different
                // warps, each into its own accumulator: a single issuing thread, not the tensor pipe, was the
                // limit (72 -> 24 -> 8 MMAs per issuing thread per pixel tile).
                const int r = warp - 1;
                const uint32_t idesc = idesc_mn(n_co, a.m_rows);
                const uint32_t d = tmem_d + (uint32_t)(r * n_co);
                for (int it = 0; it < ntiles; ++it) {
                    const int st = it % STAGES;
                    mbar_wait(full_bar + st, (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sdy = smem_u32(smem + (size_t)st * stage_bytes);
                    const uint32_t sx = sdy + a_bytes;
                    const uint64_t xd0 = desc_mn(sx, row_b, row_b, box_w * row_b);        // A: X, M groups one pixel apart
                    const uint64_t yd0 = desc_mn(sdy, row_a, lbo_dy, 8 * row_a);          // B: dY
#pragma unroll
                    for (int kk = 0; kk < TILE_M / 16; ++kk) {                    // 16 pixels (two image rows of the tile) per MMA
                        const uint64_t xd = xd0 + (uint64_t)((((r + 2 * kk) * box_w) * row_b) >> 4);
                        const uint64_t yd = yd0 + (uint64_t)((kk * 16 * row_a) >> 4);
                        umma_bf16(d, xd, yd, idesc, (it | kk) != 0);
                    }
                    umma_commit(empty_bar + st);
                }
                umma_commit(&done_bar);
            }
            __syncwarp();
        }
        // ---- epilogue: lane = (horizontal tap s, input channel ci); columns = output channels -> fp32 reductions into dW
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // accumulator row -> TMEM lane: M=128 uses all 128 lanes in order; M=64 puts rows 16w..16w+15 in the first 16 lanes
        // of warp w's lane quarter (cute::UMMA tmem_frg, "half subpartitions" atom)
        const int row = a.m_rows == 64 ? warp * 16 + lane : warp * 32 + lane;
        const int s_tap = row / a.n_chunk, ci = a.ci_offset + nc * a.n_chunk + row % a.n_chunk;
        const bool lane_ok = row < a.ks * a.n_chunk && (a.m_rows == 128 || lane < 16);
        if (a.ws_partials != nullptr) {
            // ---- deterministic split-K, part 1: this CTA's partial accumulators -> workspace (plain 16-byte stores) ----
            const int rows = a.ks * a.n_chunk;
            const size_t blk_elems = (size_t)a.ks * rows * n_co;
            float* P = a.ws_partials + ((size_t)split * a.groups + (size_t)nc * gridDim.z + mt) * blk_elems;
            for (int blk = 0; blk < a.ks; ++blk)
                for (int j = 0; j < n_co / 16; ++j) {
                    float v[16];
                    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + blk * n_co + j * 16, v);
                    if (lane_ok) {
                        float4* dst = reinterpret_cast<float4*>(P + ((size_t)blk * rows + row) * n_co + j * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                }
        } else
        for (int blk = 0; blk < a.ks; ++blk) {
            // accumulator column block: per-r accumulators sit in order r; the fused layout's block g holds r = ks - 1 - g
            const int r = a.fused_r ? a.ks - 1 - blk : blk;
            for (int j = 0; j < n_co / 16; ++j) {
                float v[16];
                tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + blk * n_co + j * 16, v);
                if (lane_ok) {
                    const int co0 = mt * 128 + j * 16;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (co0 + i < a.cout)
                            atomicAdd(a.dw + (((size_t)(co0 + i) * a.cin_total + ci) * a.ks + r) * a.ks + s_tap, v[i]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
    if (a.ws_partials == nullptr) return;
    // ---- deterministic split-K, part 2: grid barrier (the launch is one co-resident wave), then every CTA folds its
    // slice of dW over all splits, in split order, and adds it to the gradient: one plain read-modify-write per element
    const unsigned n_ctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.ws_counters, 1u);
        unsigned seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.ws_counters) : "memory"); } while (seen < n_ctas);
    }
    __syncthreads();
    {
        const int rows = a.ks * a.n_chunk;
        const unsigned blk_elems = (unsigned)(a.ks * rows * n_co);
        const unsigned total4 = blk_elems * (unsigned)a.groups / 4;        // n_co is a multiple of 16: groups of 4 output channels
        const unsigned nsplit = gridDim.x, mtiles = gridDim.z;
        const size_t split_stride = (size_t)a.groups * blk_elems;
        for (unsigned e4 = cta * THREADS + threadIdx.x; e4 < total4; e4 += n_ctas * THREADS) {
            const unsigned e = e4 * 4;
            const unsigned g = e / blk_elems, within = e - g * blk_elems;
            const unsigned co = within % (unsigned)n_co, t2 = within / (unsigned)n_co;
            const unsigned rw = t2 % (unsigned)rows, blk = t2 / (unsigned)rows;
            const float4* src = reinterpret_cast<const float4*>(a.ws_partials + (size_t)g * blk_elems + within);
            // split order is fixed (deterministic); four independent partial sums keep four 16-byte loads in flight
            float4 acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned sp = 0;
            for (; sp + 4 <= nsplit; sp += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldcg(src + (size_t)(sp + u) * split_stride / 4);
#pragma unroll
                for (int u = 0; u < 4; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
            }
            for (; sp < nsplit; ++sp) {
                const float4 v = __ldcg(src + (size_t)sp * split_stride / 4);
                acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
            }
            const float sum[4] = {(acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                                  (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w)};
            const unsigned gnc = g / mtiles, gmt = g - gnc * mtiles;
            const int r = a.fused_r ? a.ks - 1 - (int)blk : (int)blk;
            const int cig = a.ci_offset + (int)gnc * a.n_chunk + (int)(rw % (unsigned)a.n_chunk), st = (int)(rw / (unsigned)a.n_chunk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned co_g = gmt * 128 + co + q;
                // one writer per element and launch, so the result is still deterministic; a reduction (RED, no return value)
                // instead of load-add-store keeps the scattered updates off the thread's critical path
                if (co_g < (unsigned)a.cout) atomicAdd(a.dw + (((size_t)co_g * a.cin_total + cig) * a.ks + r) * a.ks + st, sum[q]);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(a.ws_counters + 1, 1u) == n_ctas - 1) {      // last CTA out: leave the counters ready for the next launch
            a.ws_counters[0] = 0; a.ws_counters[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace wgrad
}  // namespace uaps

using namespace uaps;
using namespace uaps::wgrad;

namespace {
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
int encode(CUtensorMap* map, const void* ptr, int B, int H, int W, int Cs, int box_c, int box_h, int box_w) {
    EncodeTiledFn f = encode_fn();
    if (f == nullptr) return UAPS_ENODEV;
    cuuint64_t dims[4] = {(cuuint64_t)Cs, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cs * 2, (cuuint64_t)W * Cs * 2, (cuuint64_t)H * W * Cs * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return f(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? UAPS_OK : UAPS_EINVAL;
}
}  // namespace

namespace {
struct WPlan {
    WgradArgs a;
    int splits, n_chunks, m_tiles, dy_rows, x_rows;
    size_t smem, ws_bytes;
};

// Everything uaps_conv_wgrad decides before the launch.  ws_mode: deterministic split-K through a workspace (the splits'
// partial sums cost two streaming passes over cout * cin * taps floats instead of that many fp32 atomics each, and the
// launch must be ONE co-resident wave because the CTAs meet at a grid barrier).
int plan_wgrad(int B, int H, int W, int cout, int cin, int cin_total, int ci_offset, int ks, bool ws_mode, WPlan* pl,
               int occ_cap = 0) {
    if (B <= 0 || H <= 0 || W <= 0) return UAPS_EINVAL;
    if (cout <= 0 || cin <= 0 || (ks != 1 && ks != 3) || ci_offset < 0 || ci_offset + cin > cin_total) return UAPS_EINVAL;
    const int cin_pad = (cin + 15) / 16 * 16;
    if (cin_pad != cin) return UAPS_ERANGE;                                  // callers pad Cin=3 layers on their side (see conv.py)
    WgradArgs& a = pl->a;
    a = WgradArgs{};
    a.B = B; a.H = H; a.W = W; a.cout = cout; a.cin_total = cin_total; a.ci_offset = ci_offset; a.ks = ks;
    a.n_chunk = (cin_pad % 32 == 0) ? 32 : 16;
    a.tiles_x = (W + TILE_W - 1) / TILE_W; a.tiles_y = (H + TILE_H - 1) / TILE_H;
    a.tiles_total = a.tiles_x * a.tiles_y * B;
    a.two_boxes = cout > 64;
    const int cout_pad = (cout + 15) / 16 * 16;
    a.a_ch = cout_pad >= 64 ? 64 : (cout_pad % 32 == 0 ? 32 : 16);
    a.n_co = cout_pad >= 128 ? 128 : cout_pad;
    if (a.n_co != 16 && a.n_co != 32 && a.n_co != 64 && a.n_co != 128) return UAPS_ERANGE;   // 48/80/96/112: not a UNet_UAPS shape
    static const bool no_m64 = getenv("UAPS_WGRAD_M128") != nullptr;           // A/B knob
    a.m_rows = (ks * a.n_chunk <= 64 && !no_m64) ? 64 : 128;
    static const bool no_fused = getenv("UAPS_WGRAD_PER_R") != nullptr;          // A/B knob
    a.fused_r = (ks == 3 && !a.two_boxes && a.n_co == a.a_ch && !no_fused) ? 1 : 0;
    pl->n_chunks = cin_pad / a.n_chunk; pl->m_tiles = (cout + 127) / 128;
    a.groups = pl->n_chunks * pl->m_tiles;
    const int row_a = a.a_ch * 2, row_b = a.n_chunk * 2;
    pl->dy_rows = TILE_H + (a.fused_r ? ks - 1 : 0); pl->x_rows = TILE_H + (a.fused_r ? 0 : ks - 1);
    const int a_bytes = (a.two_boxes ? 2 : 1) * pl->dy_rows * TILE_W * row_a, b_bytes = pl->x_rows * (TILE_W + ks - 1) * row_b;
    // ring depth: small-channel tiles are ~10 KB, so a deeper ring is cheap and hides the TMA latency better (measured)
    static const int stages_env = [] { const char* e = getenv("UAPS_WGRAD_STAGES"); return e ? atoi(e) : 0; }();
    const int stage_sz = (a_bytes + b_bytes + 1023) & ~1023;
    // measured (B=64, tools/gpu_probe_layers.py): 3x3 16-channel layers 100 -> 86 us with 8 stages, 32-channel 62.5 -> 58.4 us
    // with 6; the 1x1 and >= 64-channel layers are best at 3 (deeper rings only cost them resident CTAs)
    // (>= 64 channels, 3x3: these run one CTA per SM whatever the ring depth -- 4 stages: 38.7 -> 36.4 us on 64 -> 64 @64x64)
    const int auto_stages = ks == 3 && stage_sz <= 12 * 1024 ? 8 : (ks == 3 && stage_sz <= 24 * 1024 ? 6 : (ks == 3 ? 4 : 3));
    a.stages = stages_env >= 2 && stages_env <= MAX_STAGES ? stages_env : auto_stages;
    while (a.stages > 2 && (size_t)a.stages * stage_sz > 200 * 1024) --a.stages;
    pl->smem = (size_t)a.stages * stage_sz + 1024;
    // split-K: enough CTAs to fill the machine (as many as fit per SM by shared memory and the 512 TMEM columns),
    // but at least 4 pixel tiles per CTA so the 9 * n_chunk * Cout reductions of the epilogue stay amortised
    int tmem_cols = 32;
    while (tmem_cols < ks * a.n_co) tmem_cols <<= 1;
    int per_sm = (int)((227 * 1024) / (pl->smem + 2048));
    if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
    static const int cap_env = [] { const char* e = getenv("UAPS_WGRAD_CTAS_PER_SM"); return e ? atoi(e) : 4; }();   // 6 measured slower
    if (per_sm > cap_env) per_sm = cap_env;
    if (occ_cap > 0 && per_sm > occ_cap) per_sm = occ_cap;   // what the occupancy calculator says can really be co-resident
    if (per_sm < 1) per_sm = 1;
    // split-K factor: trade main-loop length against the reduction of the splits' partial sums.  Atomic mode: every split
    // adds cout * cin * taps fp32 atomics (measured ~125 reductions/ns chip-wide).  Workspace mode: every split adds one
    // streaming write and one streaming read of that many floats (L2-resident, ~4 TB/s), plus a fixed grid-barrier cost.
    const int slots = per_sm * device_info().sm_count;
    const double tile_us = (12.0 * a.n_co + 300.0) / 1900.0;                 // tensor time 12 * Cout cycles per pixel tile + issue
    const double elems = (double)cout * cin * ks * ks;
    const double red_per_split_us = ws_mode ? elems * 8.0 / 4.0e6 : elems / 125e3;
    const int max_ctas = ws_mode ? slots : 4 * slots;
    int best = 1;
    double best_t = 1e30;
    for (int sp = 1; sp <= a.tiles_total && sp * a.groups <= max_ctas; sp = sp < 8 ? sp + 1 : sp + sp / 4) {
        const int per_cta = (a.tiles_total + sp - 1) / sp;
        const int waves = (sp * a.groups + slots - 1) / slots;
        const double t = waves * per_cta * tile_us + sp * red_per_split_us + (ws_mode ? 6.0 : 3.0);
        if (t < best_t) { best_t = t; best = sp; }
    }
    // the largest single-wave split count is always a candidate (the geometric search may step over it: 121 -> 151 left
    // 27 of 148 SMs idle on the 32-channel layers)
    if (slots / a.groups >= 1) {
        const int sp = slots / a.groups < a.tiles_total ? slots / a.groups : a.tiles_total;
        const double t = ((a.tiles_total + sp - 1) / sp) * tile_us + sp * red_per_split_us + (ws_mode ? 6.0 : 3.0);
        if (t < best_t) { best_t = t; best = sp; }
    }
    a.tiles_per_cta = (a.tiles_total + best - 1) / best;
    pl->splits = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
    pl->ws_bytes = WS_HEADER + (size_t)pl->splits * a.groups * ks * (ks * a.n_chunk) * a.n_co * sizeof(float);
    return UAPS_OK;
}
}  // namespace

// Bytes of workspace that make uaps_conv_wgrad take the deterministic split-K path for this shape (0: unsupported shape).
UAPS_API size_t uaps_conv_wgrad_workspace_bytes(int B, int H, int W, int cout, int cin, int ks) {
    WPlan pl;
    if (plan_wgrad(B, H, W, cout, cin, cin, 0, ks, true, &pl) != UAPS_OK) return 0;
    return pl.ws_bytes;           // an upper bound: the launch may plan fewer splits once it knows the real occupancy
}

// dy: [B,H,W,dy_c_stride] bf16 (channels >= cout must be zero or absent), x: [B,H,W,x_c_stride] bf16 holding `cin`
// channels of one K segment; dw: fp32 [cout][cin_total][ks][ks], ACCUMULATED into (caller zeroes it), the
// segment's channels start at ci_offset.  workspace (nullable): >= uaps_conv_wgrad_workspace_bytes, 16-byte aligned, its
// first 256 bytes zero before the first use (every launch leaves them zero): deterministic split-K without atomics.
UAPS_API int uaps_conv_wgrad(const void* dy, int dy_c_stride, const void* x, int x_c_stride, float* dw, int B, int H, int W,
                             int cout, int cin, int cin_total, int ci_offset, int ks, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
    if (dy == nullptr || x == nullptr || dw == nullptr) return UAPS_EINVAL;
    if ((dy_c_stride % 8) != 0 || (x_c_stride % 8) != 0 || dy_c_stride < cout || x_c_stride < cin) return UAPS_ERANGE;
    if (!aligned_to(dy, 16) || !aligned_to(x, 16) || !aligned_to(dw, 4) || (workspace != nullptr && !aligned_to(workspace, 16)))
        return UAPS_EALIGN;
    static const bool no_ws = getenv("UAPS_WGRAD_ATOMIC") != nullptr;            // A/B knob: always the atomic epilogue
    WPlan pl;
    // Few weights (16/32-channel and most 1x1 layers): the atomics cost a few microseconds while the fold of hundreds of
    // splits would serialise -- those layers keep the atomic epilogue (measured: 85 us atomic vs 132 us folded, 16->16 @256x256).
    static const long long ws_min = [] { const char* e = getenv("UAPS_WGRAD_WS_MIN_WEIGHTS"); return e ? atoll(e) : 16384ll; }();
    bool ws_mode = workspace != nullptr && !no_ws && (long long)cout * cin * ks * ks >= ws_min;
    int rc = plan_wgrad(B, H, W, cout, cin, cin_total, ci_offset, ks, ws_mode, &pl);
    if (rc != UAPS_OK) return rc;
    if (ws_mode) {                  // re-plan against the co-residency the occupancy calculator reports for this smem size
        int occ = 0;
        cudaError_t eo = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (eo == cudaSuccess) eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_wgrad_kernel, THREADS, pl.smem);
        if (eo != cudaSuccess || occ < 1) { (void)cudaGetLastError(); ws_mode = false; }
        if (getenv("UAPS_CONV_DEBUG") != nullptr) fprintf(stderr, "wgrad occupancy API: %d CTAs/SM at %zu B smem\n", occ, pl.smem);
        rc = plan_wgrad(B, H, W, cout, cin, cin_total, ci_offset, ks, ws_mode, &pl, ws_mode ? occ : 0);
        if (rc != UAPS_OK) return rc;
    }
    if (ws_mode && pl.ws_bytes > workspace_bytes) {                              // too small: the atomic path still works
        ws_mode = false;
        rc = plan_wgrad(B, H, W, cout, cin, cin_total, ci_offset, ks, false, &pl);
        if (rc != UAPS_OK) return rc;
    }
    WgradArgs& a = pl.a;
    if (dy_c_stride < a.a_ch) return UAPS_ERANGE;                            // the dY box must lie inside the tensor's channels
    a.dw = dw;
    if (ws_mode) {
        a.ws_counters = reinterpret_cast<unsigned*>(workspace);
        a.ws_partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + WS_HEADER);
    }
    CUtensorMap mdy, mx;
    rc = encode(&mdy, dy, B, H, W, dy_c_stride, a.a_ch, pl.dy_rows, TILE_W);
    if (rc != UAPS_OK) return rc;
    rc = encode(&mx, x, B, H, W, x_c_stride, a.n_chunk, pl.x_rows, TILE_W + ks - 1);
    if (rc != UAPS_OK) return rc;
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)pl.splits, (unsigned)pl.n_chunks, (unsigned)pl.m_tiles);
    if (getenv("UAPS_CONV_DEBUG") != nullptr)
        fprintf(stderr, "wgrad %dx%d cout %d cin %d ks %d: ws %d splits %d groups %d tiles/cta %d stages %d smem %zu m_rows %d fused_r %d\n",
                H, W, cout, cin, ks, (int)ws_mode, pl.splits, a.groups, a.tiles_per_cta, a.stages, pl.smem, a.m_rows, a.fused_r);
    if (ws_mode) {
        // The grid barrier needs every CTA resident at once: it is launched as a COOPERATIVE kernel, which the driver
        // refuses (instead of deadlocking) when the grid cannot be co-resident.
        e = launch_k(conv_wgrad_kernel, grid, dim3(THREADS), pl.smem, stream, true, mdy, mx, a);
        if (e != cudaSuccess) return (int)e;
    } else {
        UAPS_LAUNCH(conv_wgrad_kernel, grid, dim3(THREADS), pl.smem, stream, mdy, mx, a);
    }
    UAPS_LAUNCH_CHECK();
    return UAPS_OK;
}
