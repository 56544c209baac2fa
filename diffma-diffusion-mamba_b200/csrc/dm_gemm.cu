// Batched bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hand-written for sm_100a:
//
//     C[g] (M x N, bf16) = epilogue( (sum_s A_s[g]) (M x K, bf16, K contiguous) . B[g] (N x K, bf16, K contiguous)^T )
//
// This is the shape of the dense projections around the scan (SURVEY.md section 8a rows a3/a6 and the out-projection
// of a4/a7): in_proj  A = [x_ssm ; x_ssm*w] (2, B*L, 512), B = in_proj.weight (2, 2048|2096, 512);
//            out_proj A = un-permuted gated scan output (2, B*L, K_dir, 1024), B = out_proj.weight (2, 512, 1024).
// What the library GEMM cannot do and this kernel does:
//   * summed-A producer: CrossMerge (reference block/mamba.py:60-82) sums the K_dir direction outputs AFTER their
//     out-projections; by linearity the sum moves in front of the GEMM.  n_sum > 1 loads the n_sum slices of an A tile
//     by TMA and a warp group adds them in shared memory (fp32 add, one bf16 rounding) before the MMA reads the tile:
//     the contraction runs over K = 1024 instead of 3 * 1024 (3x fewer flops than feeding [y0|y1|y2] . [W;W;W]^T to a
//     library GEMM) and nothing merged is ever written to HBM;
//   * epilogue hooks on the fp32 accumulator: per-row scale (soft mask / RMSNorm rstd), per-column bias, and SiLU on a
//     column range -- the in-projection emits silu(z) for the gate once per SOURCE token, so the MUFU-bound scan kernel
//     does not recompute it once per direction.
//
// Structure: persistent CTAs (one per SM) walk 128 x BN output tiles; warp-specialised:
//   warp 0 / lane 0   TMA producer: cp.async.bulk.tensor (128B swizzle) of n_sum 128x64 A tiles and a BNx64 B tile per
//                     stage into a shared-memory ring, completion on `full[s]` (expect_tx);
//   warps 8..11       (n_sum > 1 only) adders: A_0 += A_1 + A_2 in place, fence.proxy.async, arrive on `ready[s]`;
//   warp 1 / lane 0   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 N=BN K=16, 4 per stage, accumulator in TMEM,
//                     DOUBLE BUFFERED (2 x BN columns): tcgen05.commit releases the stage (`empty[s]`) and, after the
//                     last k-tile, hands the accumulator to the epilogue (`acc_full[a]`);
//   warps 4..7        epilogue: tcgen05.ld (32 lanes x 32 columns) -> scale / bias / SiLU -> bf16 -> global, then
//                     `acc_empty[a]`: the epilogue of tile i overlaps the main loop of tile i+1.
// Every mbarrier wait is bounded (trap instead of hanging the GPU if a descriptor is wrong).
#include <cuda.h>
#include <cstdlib>

#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int kATile = BM * BK * 2;                 // 16 KB

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 28); ++i)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate,
                                          uint32_t kIdesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tanh_approx_f(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct GemmParams {
    __nv_bfloat16* C;
    const float* row_scale;       // (G, M) or nullptr
    const float* bias;            // (G, N) or nullptr
    int64_t c_group_stride, c_row_stride;
    int M, N, K;
    int m_tiles, n_tiles, total_tiles;
    int silu_from;                // columns >= silu_from get SiLU (>= N: none); multiple of 32
};

// shared-memory layout: [STAGES][A_0 .. A_{NSUM-1} (16 KB each) | B (BN x 128 B)] then the mbarriers and the TMEM slot
template <int BN, int STAGES, int NSUM>
__global__ void __launch_bounds__(NSUM > 1 ? 384 : 256, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const GemmParams p) {
    constexpr uint32_t kBTile = BN * BK * 2;
    constexpr uint32_t kStageBytes = NSUM * kATile + kBTile;
    constexpr uint32_t kTmemCols = 2 * BN;                     // two accumulators (power of two: 256 or 512)
    constexpr uint32_t kIdesc = make_idesc(BN);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + STAGES * kStageBytes);
    // barriers: full[S], empty[S], ready[S] (NSUM > 1), acc_full[2], acc_empty[2]
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto ready = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto acc_full = [&](int a) { return bar0 + 8u * (3 * STAGES + a); };
    auto acc_empty = [&](int a) { return bar0 + 8u * (3 * STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_tiles = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
            mbar_init(ready(s), 128);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full(a), 1);
            mbar_init(acc_empty(a), 128);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (group, m tile, n tile): n fastest, so CTAs running side by side share their A tile in L2
    auto decode = [&](int tile, int& g, int& m0, int& n0) {
        const int nt = tile % p.n_tiles;
        const int r = tile / p.n_tiles;
        n0 = nt * BN;
        m0 = (r % p.m_tiles) * BM;
        g = r / p.m_tiles;
    };

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int g, m0, n0;
                decode(tile, g, m0, n0);
                for (int kt = 0; kt < k_tiles; ++kt, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_bounded(empty(s), ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(full(s), kStageBytes);
                    const uint32_t a_dst = smem_base + s * kStageBytes;
#pragma unroll
                    for (int q = 0; q < NSUM; ++q) tma_load_4d(a_dst + q * kATile, &map_a, full(s), kt * BK, q, m0, g);
                    tma_load_3d(a_dst + NSUM * kATile, &map_b, full(s), kt * BK, n0, g);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            uint32_t it = 0, t_local = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_local) {
                const uint32_t a = t_local & 1;
                mbar_wait_bounded(acc_empty(a), ((t_local >> 1) & 1) ^ 1);    // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + a * BN;
                for (int kt = 0; kt < k_tiles; ++kt, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_bounded(NSUM > 1 ? ready(s) : full(s), (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_base + s * kStageBytes, b_addr = a_addr + NSUM * kATile;
                    const uint64_t a_desc = smem_desc_sw128(a_addr), b_desc = smem_desc_sw128(b_addr);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)      // advance 16 elements = 32 B inside the swizzle atom
                        umma_bf16(tmem_acc, a_desc + 2 * k, b_desc + 2 * k, (kt | k) != 0, kIdesc);
                    umma_commit(empty(s));                 // stage reusable once these MMAs have read it
                }
                umma_commit(acc_full(a));                  // accumulator complete
            }
        }
    } else if (warp >= 4 && warp < 8) {                    // ===== epilogue =====
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
        uint32_t t_local = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t_local) {
            int g, m0, n0;
            decode(tile, g, m0, n0);
            const uint32_t a = t_local & 1;
            mbar_wait_bounded(acc_full(a), (t_local >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < p.M;
            const float rs = (p.row_scale && row_ok) ? __ldg(p.row_scale + static_cast<int64_t>(g) * p.M + row) : 1.0f;
            const float* bias = p.bias ? p.bias + static_cast<int64_t>(g) * p.N : nullptr;
            __nv_bfloat16* crow = p.C + static_cast<int64_t>(g) * p.c_group_stride + static_cast<int64_t>(row) * p.c_row_stride + n0;
            const uint32_t taddr = tmem_base + a * BN + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                if (n0 + c >= p.N) break;                  // ragged N: whole 32-column chunks beyond N (uniform per CTA)
                uint32_t v[32];
                tmem_ld_32x32(taddr + c, v);
                tmem_ld_wait();
                const bool act = n0 + c >= p.silu_from;
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        if (n0 + c + i < p.N) {            // N is a multiple of 8 (checked on the host)
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[i + j]) * rs;
                            if (bias) {
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + i));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + i + 4));
                                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                            }
                            if (act) {                      // silu(x) = h + h tanh(h), h = x/2 (one MUFU)
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float h = 0.5f * f[j];
                                    f[j] = fmaf(h, tanh_approx_f(h), h);
                                }
                            }
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                w[j] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            *reinterpret_cast<uint4*>(crow + c + i) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(acc_empty(a));                     // 128 arrivals: the MMA warp may overwrite this accumulator
        }
    } else if (NSUM > 1 && warp >= 8) {                    // ===== adders: A_0 += A_1 + ... (same swizzled layout) =====
        const int t = threadIdx.x - 256;                   // 0..127
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int kt = 0; kt < k_tiles; ++kt, ++it) {
                const int s = it % STAGES;
                mbar_wait_bounded(full(s), (it / STAGES) & 1);
                uint8_t* a0 = aligned + s * kStageBytes;
#pragma unroll
                for (int i = 0; i < kATile / 16 / 128; ++i) {
                    const int off = (t + 128 * i) * 16;
                    uint4 acc4 = *reinterpret_cast<const uint4*>(a0 + off);
                    float f[8];
                    {
                        const uint32_t w[4] = {acc4.x, acc4.y, acc4.z, acc4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            f[2 * j] = __uint_as_float(w[j] << 16);
                            f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
                        }
                    }
#pragma unroll
                    for (int q = 1; q < NSUM; ++q) {
                        const uint4 x = *reinterpret_cast<const uint4*>(a0 + q * kATile + off);
                        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            f[2 * j] += __uint_as_float(w[j] << 16);
                            f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                        }
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                        o[j] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    *reinterpret_cast<uint4*>(a0 + off) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async-proxy reads
                mbar_arrive(ready(s));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// (G, rows, K) bf16 row-major operand -> 3-D tensor map with a 64 x box_rows box and the 128-byte swizzle
int make_map3(CUtensorMap* map, const void* base, int G, int rows, int K, int64_t group_stride, int64_t row_stride,
              int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return DM_ERR_CUDA;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(G)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride) * 2, static_cast<cuuint64_t>(group_stride) * 2};
    const cuuint32_t box[3] = {BK, static_cast<cuuint32_t>(box_rows), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? DM_OK : DM_ERR_INVALID_ARG;
}
// A operand with its summed slices as a dimension: (G, rows, n_sum, K) -> 4-D map, box 64 x 1 x 128 x 1
int make_map_a(CUtensorMap* map, const void* base, int G, int rows, int n_sum, int K, int64_t group_stride,
               int64_t row_stride, int64_t sum_stride) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return DM_ERR_CUDA;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(n_sum), static_cast<cuuint64_t>(rows),
                                static_cast<cuuint64_t>(G)};
    // a size-1 dimension still needs a legal (16-byte multiple, non-zero) stride
    const cuuint64_t ss = n_sum > 1 ? static_cast<cuuint64_t>(sum_stride) * 2 : static_cast<cuuint64_t>(row_stride) * 2;
    const cuuint64_t strides[3] = {ss, static_cast<cuuint64_t>(row_stride) * 2, static_cast<cuuint64_t>(group_stride) * 2};
    const cuuint32_t box[4] = {BK, 1, BM, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? DM_OK : DM_ERR_INVALID_ARG;
}

template <int BN, int STAGES, int NSUM>
int launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, GemmParams p, int groups, int dev, int n_sm,
                cudaStream_t cs) {
    p.m_tiles = (p.M + BM - 1) / BM;
    p.n_tiles = (p.N + BN - 1) / BN;
    p.total_tiles = p.m_tiles * p.n_tiles * groups;
    const size_t smem = static_cast<size_t>(STAGES) * (NSUM * kATile + BN * BK * 2) + 1024 + 256;
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        DM_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, STAGES, NSUM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.set(dev);
    }
    const int grid = p.total_tiles < n_sm ? p.total_tiles : n_sm;
    gemm_bf16_tcgen05_kernel<BN, STAGES, NSUM><<<grid, NSUM > 1 ? 384 : 256, smem, cs>>>(map_a, map_b, p);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

}  // namespace
}  // namespace dm

extern "C" int dm_gemm_bf16_tn_ex(const dm_gemm_args* a, void* stream) {
    using namespace dm;
    if (!a || !a->A || !a->B || !a->C || a->groups <= 0 || a->M <= 0 || a->N <= 0 || a->K <= 0) return DM_ERR_INVALID_ARG;
    if (a->n_sum != 1 && a->n_sum != 3) return DM_ERR_UNSUPPORTED;
    if (a->K % 8 || a->N % 8 || a->a_row_stride % 8 || a->b_row_stride % 8 || a->a_group_stride % 8 ||
        a->b_group_stride % 8 || a->c_row_stride % 8 || (a->n_sum > 1 && a->a_sum_stride % 8))
        return DM_ERR_INVALID_ARG;
    if (!aligned16(a->A) || !aligned16(a->B) || !aligned16(a->C) || (a->bias && !aligned16(a->bias))) return DM_ERR_INVALID_ARG;
    if (a->silu_from < 0 || (a->silu_from < a->N && a->silu_from % 32)) return DM_ERR_INVALID_ARG;
    int dev = 0, n_sm = 0;
    if (int e = current_device(&dev, &n_sm); e != DM_OK) return e;
    // tile width: 256 columns when that still gives at least half the SMs a tile (half the re-reads of A), else 128.
    // DM_GEMM_BN (128 | 256) overrides for experiments.
    static const int forced_bn = env_int("DM_GEMM_BN", 0);
    const int m_tiles = (a->M + BM - 1) / BM;
    int bn = (a->N >= 256 && static_cast<long long>(m_tiles) * ((a->N + 255) / 256) * a->groups >= n_sm / 2) ? 256 : 128;
    if (forced_bn == 128 || forced_bn == 256) bn = forced_bn;
    CUtensorMap map_a, map_b;
    int st = make_map_a(&map_a, a->A, a->groups, a->M, a->n_sum, a->K, a->a_group_stride, a->a_row_stride, a->a_sum_stride);
    if (st != DM_OK) return st;
    st = make_map3(&map_b, a->B, a->groups, a->N, a->K, a->b_group_stride, a->b_row_stride, bn);
    if (st != DM_OK) return st;
    GemmParams p{};
    p.C = static_cast<__nv_bfloat16*>(a->C);
    p.row_scale = a->row_scale; p.bias = a->bias;
    p.c_group_stride = a->c_group_stride; p.c_row_stride = a->c_row_stride;
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.silu_from = a->silu_from >= a->N ? (1 << 30) : a->silu_from;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (a->n_sum == 1) {
        // stage = 16 KB (A) + 16 / 32 KB (B): 4 stages = 128 / 192 KB
        return bn == 256 ? launch_gemm<256, 4, 1>(map_a, map_b, p, a->groups, dev, n_sm, cs)
                         : launch_gemm<128, 4, 1>(map_a, map_b, p, a->groups, dev, n_sm, cs);
    }
    // stage = 48 KB (3 A slices) + 16 / 32 KB: 3 x 64 KB or 2 x 80 KB
    return bn == 256 ? launch_gemm<256, 2, 3>(map_a, map_b, p, a->groups, dev, n_sm, cs)
                     : launch_gemm<128, 3, 3>(map_a, map_b, p, a->groups, dev, n_sm, cs);
}

extern "C" int dm_gemm_bf16_tn(const void* A, int64_t a_group_stride, int64_t a_row_stride, const void* B,
                               int64_t b_group_stride, int64_t b_row_stride, void* C, int64_t c_group_stride,
                               int64_t c_row_stride, const float* row_scale, int32_t groups, int32_t M, int32_t N,
                               int32_t K, void* stream) {
    dm_gemm_args a{};
    a.A = A; a.a_group_stride = a_group_stride; a.a_row_stride = a_row_stride; a.a_sum_stride = 0; a.n_sum = 1;
    a.B = B; a.b_group_stride = b_group_stride; a.b_row_stride = b_row_stride;
    a.C = C; a.c_group_stride = c_group_stride; a.c_row_stride = c_row_stride;
    a.row_scale = row_scale; a.bias = nullptr; a.silu_from = N;
    a.groups = groups; a.M = M; a.N = N; a.K = K;
    return dm_gemm_bf16_tn_ex(&a, stream);
}
