// Mamba-1 backward for sm_100a (upstream MambaInnerFn.backward =
// selective_scan_cuda.bwd + causal_conv1d_cuda.causal_conv1d_bwd + einsum weight gradients; reached from the
// reference through train.py:259).  Everything here works in SCAN order; the host un-permutes and sums directions.
//
//   m1_scan_bwd_kernel   one warp per (sequence, 32 channels), lane = channel.
//        sweep 1 (forward): recompute the recurrence, store the state at every chunk boundary (DM_BWD_CH = 4 tokens) in a
//        workspace -- SKIPPED when the training forward already wrote these checkpoints (dm_mamba1_group.chunk_states,
//        states_valid = 1: the normal case under autograd_ops.Mamba1ScanFn)
//        sweep 2 (reverse, chunk by chunk): the chunk's checkpoint, x_dbl rows, u, z and dout are staged by cp.async (double
//        buffered, the scan-order lookup one chunk ahead); recompute the chunk's states into shared memory (packed pairs of
//        states), then run the adjoint recurrence  dh_{j-1} = a_j dh_j,  dh_j += dy_j C_j  backwards on packed fp32x2 math,
//        producing dz, du (scan part), d(delta_raw) per token, dB/dC (bf16 activations: column sums of a [channel][value]
//        matrix by an all-ones MMA read with ldmatrix.trans; fp32 activations: fp32 transposition; then one atomic per
//        value), and dA / dD / d(dt_bias) accumulated in registers.  dout may be broadcast over the directions (stride 0).
//   m1_conv_bwd_kernel   one thread per (sequence, channel): recompute the conv pre-activation, dc = du * silu'(c),
//        dx by the 4-tap correlation with a sliding window of dc, dw / dbias accumulated and added atomically.
//
// The GEMM-shaped pieces (d dt_low = d_delta . W_dt, dW_dt, du += d_x_dbl . W_x, dW_x) are plain library GEMMs on the
// host side (autograd_ops.py): they are <2 % of the backward's time and have no fusion partner.
#include "dm_common.cuh"

namespace dm {
namespace {

#ifndef DM_BWD_CH
#define DM_BWD_CH 4           // tokens per backward chunk: 4 keeps the per-warp state buffer at 8 KB => 12 warps per SM
#endif
#ifndef DM_BWD_MINB
#define DM_BWD_MINB 12
#endif
constexpr int kN = 16, kW = 4, kE = 64, kR = 32, kCH = DM_BWD_CH;
constexpr int kMR = 8;        // rows of the dt_proj MMA tile (m16n8k16, 8 token rows used); xd / ds are sized for it

struct B1G {
    const void* xz; int64_t xz_bs, xz_ts;
    const void* dout; int64_t do_bs, do_ds, do_ts;
    const void* u; const float* x_dbl;
    const void* wdt; const float* dt_bias; const float* A; const float* D;
    float* d_xz_scan; float* du; float* ddelta; float* d_x_dbl; float* dA; float* dD; float* d_dt_bias; float* hb;
    const float* conv_w; const float* conv_b; float* d_conv_w; float* d_conv_b; const float* du_total;
    int states_valid;
};
struct B1P {
    int B, K, L, D, out_order, n_groups;
    const int32_t* order;
    B1G g[DM_MAX_GROUPS];
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    hi = pack_bf16(__bfloat162float(ah), __bfloat162float(bh));
    lo = pack_bf16(a - __bfloat162float(ah), b - __bfloat162float(bh));
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float softplus_fast(float x) {       // same formula as the forward kernel
    const float e = ex2_approx(x * kLog2e);
    const float big = lg2_approx(1.0f + e) * 0.6931471805599453f;
    const float small = e * fmaf(e, fmaf(e, 0.33333333f, -0.5f), 1.0f);
    const float r = e < 0.0078125f ? small : big;
    return x > 20.0f ? x : r;
}
__device__ __forceinline__ const int32_t* dir_order(const B1P& p, int k) {
    if (p.order == nullptr) return nullptr;
    const int32_t* o = p.order + static_cast<int64_t>(k) * p.L;
    return (__ldg(o) < 0) ? nullptr : o;
}

constexpr int kXd = kE + 4;
constexpr int kCk = kN + 4;
constexpr int kTbhStride = 40;      // bf16 elements per channel row of the bf16 transposition matrix (80 bytes)
constexpr int kTbStride = 34;   // float2 units per pair row of the transposition buffer: 2 * 34 words = 4 (mod 32), so
                                // the eight lanes of a 128-bit load phase (pair rows p .. p + 7) cover all 32 banks
template <typename T> struct BwdSmem {
    uint64_t hs[kCH - 1][kN / 2][32];   // state BEFORE tokens 1 .. kCH-1 of the chunk, as packed pairs (n, n + 1)
    float ck[2][32][kCk];        // state before token 0 = the chunk's checkpoint, staged by cp.async with the chunk's other
                                 // inputs ([channel][16 states], 80-byte rows: 128-bit reads by lane are conflict free)
    float xd[2][kCH][kXd];       // x_dbl rows of the chunk, double buffered (cp.async); 272-byte row stride: the four
                                 // rows the dt_proj MMA gathers its A fragment from fall into different banks
    float ds[kCH][40];           // delta_raw of the chunk's tokens (row stride 40 words: the MMA epilogue's 8-byte stores
                                 // of rows 0..3 are conflict free)
    // dB / dC reduction over the warp's 32 channels.  fp32 activations: [pair of values][channel] as packed fp32 pairs,
    // summed by 128-bit row reads.  bf16 activations: the same memory holds a [channel][value] bf16 matrix (80-byte row
    // stride) that ldmatrix.trans turns into the B operand of an all-ones MMA: half the shared-memory wavefronts.
    uint64_t tb[sizeof(T) == 4 ? kN * kTbStride : 32 * kTbhStride / 4];
    T us[2][kCH][32], zs[2][kCH][32], dos[2][kCH][32];
};

template <typename T>
__global__ void __launch_bounds__(32, DM_BWD_MINB) m1_scan_bwd_kernel(const __grid_constant__ B1P p, int n_units) {
    constexpr bool kSplit = sizeof(T) == 4;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    BwdSmem<T>& S = *reinterpret_cast<BwdSmem<T>*>(smem_raw);
    const int lane = threadIdx.x, unit = blockIdx.x;
    if (unit >= n_units) return;
    const int D = p.D, L = p.L;
    const int slices = D >> 5;
    const int cs = unit % slices, seq = unit / slices;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const B1G& G = p.g[g];
    const int c0 = cs * 32, c = c0 + lane;
    const int32_t* ord = dir_order(p, k);
    const int64_t sg = static_cast<int64_t>(b) * p.K + k;                       // sequence index inside the group
    const T* u_seq = static_cast<const T*>(G.u) + sg * L * D + c;
    const T* z_base = static_cast<const T*>(G.xz) + static_cast<int64_t>(b) * G.xz_bs + D + c;
    const T* do_base = static_cast<const T*>(G.dout) + static_cast<int64_t>(b) * G.do_bs + static_cast<int64_t>(k) * G.do_ds + c;
    const float* xd_seq = G.x_dbl + sg * L * kE;
    const int n_chunks = (L + kCH - 1) / kCH;
    float* hb = G.hb + ((sg * n_chunks) * D + c) * kN;                           // [chunk][D][16], this lane's channel
    const bool token_order = p.out_order == DM_OUT_TOKEN_ORDER;

    uint64_t A2[kN / 2];                                                         // A * log2(e), pairs (n, n + 1)
#pragma unroll
    for (int n = 0; n < kN; n += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(G.A + static_cast<int64_t>(c) * kN + n));
        A2[n / 2] = pack2(t.x * kLog2e, t.y * kLog2e);
        A2[n / 2 + 1] = pack2(t.z * kLog2e, t.w * kLog2e);
    }
    // decay pair a = 2^(dt * A2) on the MUFU
    auto decay2 = [](uint64_t dt2, uint64_t a2) {
        float e0, e1;
        unpack2(mul2(dt2, a2), e0, e1);
        return pack2(ex2_approx(e0), ex2_approx(e1));
    };
    const float dtb = G.dt_bias ? __ldg(G.dt_bias + c) : 0.f;
    const float Dc = G.D ? __ldg(G.D + c) : 0.f;

    // W_dt fragments (as in the forward kernel, 32 channels -> 4 n-tiles), held in registers for the whole sequence
    // (re-fetching them per chunk to relieve the register allocator measured slower: 278 vs 270 us)
    const T* Wdt = static_cast<const T*>(G.wdt) + static_cast<int64_t>(c0 + (lane >> 2)) * kR + 2 * (lane & 3);
    auto load_w = [&](uint32_t (&bw_hi)[4][2][2], uint32_t (&bw_lo)[4][2][2]) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const T* wp = Wdt + nt * 8 * kR + ks * 16;
                if constexpr (kSplit) {
                    const float2 w0 = __ldg(reinterpret_cast<const float2*>(wp));
                    const float2 w1 = __ldg(reinterpret_cast<const float2*>(wp + 8));
                    split_bf16(w0.x, w0.y, bw_hi[nt][ks][0], bw_lo[nt][ks][0]);
                    split_bf16(w1.x, w1.y, bw_hi[nt][ks][1], bw_lo[nt][ks][1]);
                } else {
                    bw_hi[nt][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(wp));
                    bw_hi[nt][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(wp + 8));
                }
            }
    };

    // Chunk staging, double buffered: issue() starts the cp.async copies of chunk ci (x_dbl rows, u, and for the reverse
    // sweep z and dout, 16-byte segments straight to shared memory) and returns at once; finish() waits for the OLDEST
    // outstanding group and leaves delta_raw of that chunk in S.ds (same MMA as the forward).  The next chunk's loads are
    // always in flight while the current one is processed: the first version loaded synchronously and spent most of
    // its time waiting on L2 with 2-3 warps per sub-partition.
    constexpr int kEpS = 16 / static_cast<int>(sizeof(T));     // elements per 16-byte segment
    constexpr int kSegRow = 32 / kEpS;                         // segments per (token, 32 channels) row
    const T* u0 = u_seq - lane;
    const T* z0 = z_base - lane;
    const T* do0 = do_base - lane;
    // source token of this lane's (row, segment) in the chunk the reverse sweep issues NEXT: read one chunk ahead, so
    // the scan-order lookup never sits between issue() and its dependent copies
    // (returned by value and handed to issue() as an argument: captured by reference it lived in local memory, and the
    // store behind the load stalled every chunk on the lookup's latency)
    auto load_src = [&](int ci) -> int {
        int src = 0;
        if (ci >= 0 && lane < kCH * kSegRow) {
            const int j = min(ci * kCH + lane / kSegRow, L - 1);
            src = ord ? __ldg(ord + j) : j;
        }
        return src;
    };
    auto issue = [&](int ci, int buf, bool with_grad, int src) {
        const int j0 = ci * kCH;
        for (int sgm = lane; sgm < kCH * (kE / 4); sgm += 32) {
            const int r = sgm / (kE / 4), part = sgm % (kE / 4);
            const int j = min(j0 + r, L - 1);
            cp_async16(smem_u32(&S.xd[buf][r][part * 4]), xd_seq + static_cast<int64_t>(j) * kE + part * 4);
        }
        for (int sgm = lane; sgm < kCH * kSegRow; sgm += 32) {
            const int r = sgm / kSegRow, part = sgm % kSegRow;
            const int j = min(j0 + r, L - 1);
            cp_async16(smem_u32(&S.us[buf][r][part * kEpS]), u0 + static_cast<int64_t>(j) * D + part * kEpS);
            if (with_grad) {
                // (kCH * kSegRow <= 32: this loop runs once per lane, src is this lane's row)
                cp_async16(smem_u32(&S.zs[buf][r][part * kEpS]), z0 + static_cast<int64_t>(src) * G.xz_ts + part * kEpS);
                cp_async16(smem_u32(&S.dos[buf][r][part * kEpS]),
                           do0 + static_cast<int64_t>(token_order ? src : j) * G.do_ts + part * kEpS);
            }
        }
        if (with_grad) {
            // the chunk's checkpoint: 32 channels x 16 states = 2 KB contiguous, 16-byte segments, coalesced
            const float* ckg = G.hb + ((sg * n_chunks + ci) * D + c0) * kN;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = i * 32 + lane;
                cp_async16(smem_u32(&S.ck[buf][idx >> 2][(idx & 3) * 4]), ckg + idx * 4);
            }
        }
        cp_async_commit();
    };
    static_assert(kCH * kSegRow <= 32, "one (row, segment) of u / z / dout per lane");
    uint32_t bw_hi[4][2][2], bw_lo[4][2][2];
    load_w(bw_hi, bw_lo);
    auto finish = [&](int buf, bool more_in_flight) {
        if (more_in_flight) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        float dacc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) dacc[nt][i] = 0.f;
        const int r = lane >> 2, q = lane & 3;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(&S.xd[buf][r < kCH ? r : kCH - 1][0]);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_hi[4] = {row[ks * 8 + q], 0u, row[ks * 8 + 4 + q], 0u};
            const uint32_t a_lo[4] = {row[16 + ks * 8 + q], 0u, row[16 + ks * 8 + 4 + q], 0u};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                mma_bf16_16816(dacc[nt], a_hi, bw_hi[nt][ks][0], bw_hi[nt][ks][1]);
                mma_bf16_16816(dacc[nt], a_lo, bw_hi[nt][ks][0], bw_hi[nt][ks][1]);
                if constexpr (kSplit) mma_bf16_16816(dacc[nt], a_hi, bw_lo[nt][ks][0], bw_lo[nt][ks][1]);
            }
        }
        if (r < kCH) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<float2*>(&S.ds[r][nt * 8 + 2 * q]) = make_float2(dacc[nt][0], dacc[nt][1]);
        }
        __syncwarp();
    };

    // ---- sweep 1: forward, boundary states -> workspace ----
    uint64_t h[kN / 2];
#pragma unroll
    for (int q = 0; q < kN / 2; ++q) h[q] = 0ull;
    // (skipped when the training forward already stored the checkpoints: G.states_valid)
    if (n_chunks > 1 && !G.states_valid) issue(0, 0, false, 0);
    for (int ci = 0; ci < n_chunks && !G.states_valid; ++ci) {
        float* hbc = hb + static_cast<int64_t>(ci) * D * kN;
#pragma unroll
        for (int q = 0; q < kN / 2; q += 2) *reinterpret_cast<ulonglong2*>(hbc + 2 * q) = make_ulonglong2(h[q], h[q + 1]);
        if (ci == n_chunks - 1) break;                       // the last chunk's end state is never needed
        const int buf = ci & 1;
        const bool more = ci + 1 < n_chunks - 1;             // sweep 1 visits chunks 0 .. n_chunks - 2
        __syncwarp();                                        // every lane is done with the buffer the prefetch overwrites
        if (more) issue(ci + 1, buf ^ 1, false, 0);
        finish(buf, more);
#pragma unroll
        for (int jj = 0; jj < kCH; ++jj) {
            const float dt = softplus_fast(S.ds[jj][lane] + dtb);
            const float dtu = dt * to_f32<T>(S.us[buf][jj][lane]);
            const uint64_t dt2 = pack2(dt, dt), dtu2 = pack2(dtu, dtu);
            const ulonglong2* Bv = reinterpret_cast<const ulonglong2*>(&S.xd[buf][jj][kR]);
#pragma unroll
            for (int q = 0; q < kN / 2; q += 2) {
                const ulonglong2 b2 = Bv[q / 2];
                h[q] = fma2(decay2(dt2, A2[q]), h[q], mul2(dtu2, b2.x));
                h[q + 1] = fma2(decay2(dt2, A2[q + 1]), h[q + 1], mul2(dtu2, b2.y));
            }
        }
    }

    // ---- sweep 2: reverse ----
    if (!G.states_valid) __threadfence_block();                // sweep 1's checkpoints are re-read through cp.async below
    // All the per-state arithmetic runs on packed pairs (fma / mul .f32x2): the kernel was bound by instruction issue and
    // the shared-memory pipe together (r02 ncu: 627 warp instructions and 153 shared wavefronts per token, issue 62 %,
    // LSU wavefronts 69 %), and 52 % of those instructions were scalar FMUL / FFMA / FADD of this loop.
    uint64_t dh[kN / 2], dA[kN / 2];
#pragma unroll
    for (int q = 0; q < kN / 2; ++q) { dh[q] = 0ull; dA[q] = 0ull; }
    float dD = 0.f, ddtb = 0.f;
    float* dz_out = G.d_xz_scan + sg * L * 2 * D + D + c;
    float* du_out = G.du + sg * L * D + c;
    float* dd_out = G.ddelta + sg * L * D + c;
    float* dxd_out = G.d_x_dbl + sg * L * kE;
    // reduction roles: lane (pr, half) sums pair row pr of the transposition buffer over channels half * 16 .. + 15,
    // then the two halves swap one component: the lane ends with value 2 * pr + half (0..15 = dB, 16..31 = dC)
    const int pr = lane & 15, half = lane >> 4;
    __syncwarp();
    issue(n_chunks - 1, (n_chunks - 1) & 1, true, load_src(n_chunks - 1));
    int src_next = load_src(n_chunks - 2);
    for (int ci = n_chunks - 1; ci >= 0; --ci) {
        const int j0 = ci * kCH, nrows = min(kCH, L - j0);
        const int buf = ci & 1;
        __syncwarp();                                        // every lane is done with the buffer the prefetch overwrites
        if (ci > 0) {
            issue(ci - 1, buf ^ 1, true, src_next);
            src_next = load_src(ci - 2);
        }
        finish(buf, ci > 0);
        const ulonglong2* ckl = reinterpret_cast<const ulonglong2*>(&S.ck[buf][lane][0]);
#pragma unroll
        for (int q = 0; q < kN / 2; q += 2) {
            const ulonglong2 t = ckl[q / 2];
            h[q] = t.x; h[q + 1] = t.y;
        }
        float dtv[kCH];
        // forward through the chunk: keep the state before each token in shared memory (y is rebuilt by the adjoint pass
        // from the state after the token, which it holds anyway: one broadcast read of C per token instead of two)
#pragma unroll
        for (int jj = 0; jj < kCH; ++jj) {
            const float dt = softplus_fast(S.ds[jj][lane] + dtb);
            const float uu = to_f32<T>(S.us[buf][jj][lane]);
            const float dtu = dt * uu;
            const uint64_t dt2 = pack2(dt, dt), dtu2 = pack2(dtu, dtu);
            const ulonglong2* Bv = reinterpret_cast<const ulonglong2*>(&S.xd[buf][jj][kR]);
#pragma unroll
            for (int q = 0; q < kN / 2; q += 2) {
                const ulonglong2 b2 = Bv[q / 2];
                if (jj > 0) {
                    S.hs[jj - 1][q][lane] = h[q];
                    S.hs[jj - 1][q + 1][lane] = h[q + 1];
                }
                h[q] = fma2(decay2(dt2, A2[q]), h[q], mul2(dtu2, b2.x));
                h[q + 1] = fma2(decay2(dt2, A2[q + 1]), h[q + 1], mul2(dtu2, b2.y));
            }
            dtv[jj] = dt;
        }
        // adjoint recurrence, last token of the chunk first.  h = state AFTER the token being processed: the chunk's end
        // state first, then the "state before" of the token handled one iteration earlier.
#pragma unroll
        for (int jj = kCH - 1; jj >= 0; --jj) {
            if (jj < nrows) {
                const int j = j0 + jj;
                const float dt = dtv[jj];
                const float uu = to_f32<T>(S.us[buf][jj][lane]);
                const float zz = to_f32<T>(S.zs[buf][jj][lane]);
                const float go = to_f32<T>(S.dos[buf][jj][lane]);
                const float sg_z = sigmoid_fast(zz);
                const float dy = go * zz * sg_z;                                   // d out / d y = silu(z)
                dD = fmaf(dy, uu, dD);
                const float w = dt * uu;
                const uint64_t dt2 = pack2(dt, dt), dy2 = pack2(dy, dy), w2 = pack2(w, w);
                const ulonglong2* Bv = reinterpret_cast<const ulonglong2*>(&S.xd[buf][jj][kR]);
                const ulonglong2* Cv = reinterpret_cast<const ulonglong2*>(&S.xd[buf][jj][kR + kN]);
                uint64_t ddt2 = 0ull, dbu2 = 0ull, y2 = 0ull;   // d delta (before the ln2 scale), sum_n dh_n B_n, y
                uint32_t pw[kSplit ? 1 : kN];          // bf16 path: this channel's row of the [channel][value] matrix
                __syncwarp();                          // the previous token's reduction has read the buffer
#pragma unroll
                for (int q = 0; q < kN / 2; q += 2) {
                    const ulonglong2 bq = Bv[q / 2], cq = Cv[q / 2];
                    ulonglong2 hq;                                                  // state before the token, pairs q, q + 1
                    if (jj > 0) hq = make_ulonglong2(S.hs[jj - 1][q][lane], S.hs[jj - 1][q + 1][lane]);
                    else hq = ckl[q / 2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint64_t b2 = e ? bq.y : bq.x, c2 = e ? cq.y : cq.x;
                        const int qq = q + e;
                        const uint64_t hp = e ? hq.y : hq.x;
                        const uint64_t a = decay2(dt2, A2[qq]);
                        y2 = fma2(h[qq], c2, y2);
                        const uint64_t pc = mul2(dy2, h[qq]);                       // dC_j contribution of this channel
                        dh[qq] = fma2(dy2, c2, dh[qq]);                             // dL/dh_j
                        dbu2 = fma2(dh[qq], b2, dbu2);
                        const uint64_t pb = mul2(dh[qq], w2);                       // dB_j contribution
                        if constexpr (kSplit) {
                            S.tb[(kN / 2 + qq) * kTbStride + lane] = pc;
                            S.tb[qq * kTbStride + lane] = pb;
                        } else {
                            float x0, x1;
                            unpack2(pb, x0, x1);
                            pw[qq] = pack_bf16(x0, x1);
                            unpack2(pc, x0, x1);
                            pw[kN / 2 + qq] = pack_bf16(x0, x1);
                        }
                        dh[qq] = mul2(dh[qq], a);                                   // -> dL/dh_{j-1} through the decay
                        const uint64_t t = mul2(dh[qq], hp);                        // dh_j * a * h_{j-1}
                        dA[qq] = fma2(t, dt2, dA[qq]);
                        ddt2 = fma2(t, A2[qq], ddt2);                               // (scaled by ln2 below)
                        h[qq] = hp;
                    }
                    if constexpr (!kSplit) {
                        if ((q & 2) != 0) {            // four pairs of dB and of dC are complete: two 16-byte row segments
                            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(&S.tb[0]) + lane * kTbhStride);
                            const int i = q >> 2;
                            dst[i] = make_uint4(pw[4 * i], pw[4 * i + 1], pw[4 * i + 2], pw[4 * i + 3]);
                            dst[2 + i] = make_uint4(pw[8 + 4 * i], pw[9 + 4 * i], pw[10 + 4 * i], pw[11 + 4 * i]);
                        }
                    }
                }
                float s0, s1, t0, t1, ya, yb;
                unpack2(ddt2, s0, s1);
                unpack2(dbu2, t0, t1);
                unpack2(y2, ya, yb);
                const float dz = go * fmaf(Dc, uu, ya + yb) * sg_z * fmaf(zz, 1.0f - sg_z, 1.0f);   // silu'(z) = s (1 + z (1 - s))
                const float dbu = t0 + t1;
                const float ddt = fmaf(s0 + s1, 0.6931471805599453f, dbu * uu);
                const float du = fmaf(dy, Dc, dbu * dt);
                const float ddr = ddt * sigmoid_fast(S.ds[jj][lane] + dtb);          // softplus'(x) = sigmoid(x)
                ddtb += ddr;
                dz_out[static_cast<int64_t>(j) * 2 * D] = dz;
                du_out[static_cast<int64_t>(j) * D] = du;
                dd_out[static_cast<int64_t>(j) * D] = ddr;
                // reduce dB / dC over the warp's 32 channels
                if constexpr (kSplit) {
                    __syncwarp();
                    const ulonglong2* trow = reinterpret_cast<const ulonglong2*>(&S.tb[pr * kTbStride + half * 16]);
                    uint64_t r0 = 0ull, r1 = 0ull;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const ulonglong2 v = trow[i];
                        r0 = add2(r0, v.x);
                        r1 = add2(r1, v.y);
                    }
                    float lo, hi;
                    unpack2(add2(r0, r1), lo, hi);
                    const float mine = half ? hi : lo, other = half ? lo : hi;
                    const float s = mine + __shfl_xor_sync(0xffffffffu, other, 16);
                    atomicAdd(dxd_out + static_cast<int64_t>(j) * kE + kR + 2 * pr + half, s);
                } else {
                    // column sums of the [32 channels][32 values] bf16 matrix = ones[16 x 32] . M on the tensor core: the
                    // terms are rounded to bf16 (the activations they are built from already are), the sum is fp32
                    __nv_bfloat16* tbh = reinterpret_cast<__nv_bfloat16*>(&S.tb[0]);
                    __syncwarp();
                    const uint32_t ones[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
                    float acc[4][4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        uint32_t b0, b1, b2, b3;       // k = channels 0-7, 8-15, 16-23, 24-31 of value columns 8t .. 8t+7
                        ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(tbh + lane * kTbhStride + 8 * t));
                        acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
                        mma_bf16_16816(acc[t], ones, b0, b1);
                        mma_bf16_16816(acc[t], ones, b2, b3);
                    }
                    // every row of the product is the same: lane (row r, quad q) adds value 8 (r & 3) + 2 q + (r >> 2)
                    const int r = lane >> 2, q = lane & 3, e = (r >> 2) & 1, t = r & 3;
                    const float v0 = e ? acc[0][1] : acc[0][0], v1 = e ? acc[1][1] : acc[1][0];
                    const float v2 = e ? acc[2][1] : acc[2][0], v3 = e ? acc[3][1] : acc[3][0];
                    const float s = (t & 2) ? ((t & 1) ? v3 : v2) : ((t & 1) ? v1 : v0);
                    atomicAdd(dxd_out + static_cast<int64_t>(j) * kE + kR + 8 * t + 2 * q + e, s);
                }
            } else {
                if (jj > 0) {                                                        // tail chunk (jj >= nrows >= 1)
#pragma unroll
                    for (int q = 0; q < kN / 2; ++q) h[q] = S.hs[jj - 1][q][lane];  // state after token jj - 1
                }
            }
        }
    }
    float* dAo = G.dA + static_cast<int64_t>(c) * kN;
#pragma unroll
    for (int q = 0; q < kN / 2; ++q) {
        float x0, x1;
        unpack2(dA[q], x0, x1);
        atomicAdd(dAo + 2 * q, x0);
        atomicAdd(dAo + 2 * q + 1, x1);
    }
    if (G.dD) atomicAdd(G.dD + c, dD);
    if (G.d_dt_bias) atomicAdd(G.d_dt_bias + c, ddtb);
}

// conv backward: thread = (sequence, channel)
template <typename T>
__global__ void __launch_bounds__(128) m1_conv_bwd_kernel(const __grid_constant__ B1P p) {
    const int D = p.D, L = p.L;
    const int cblocks = D / 128;
    const int seq = blockIdx.x / cblocks, c = (blockIdx.x % cblocks) * 128 + threadIdx.x;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const B1G& G = p.g[g];
    const int32_t* ord = dir_order(p, k);
    const int64_t sg = static_cast<int64_t>(b) * p.K + k;
    const T* x_base = static_cast<const T*>(G.xz) + static_cast<int64_t>(b) * G.xz_bs + c;
    const float* du = G.du_total + sg * L * D + c;
    float* dx_out = G.d_xz_scan + sg * L * 2 * D + c;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c) * kW));
    const float w[kW] = {wv.x, wv.y, wv.z, wv.w};
    const float bias = G.conv_b ? __ldg(G.conv_b + c) : 0.f;
    float xw[3] = {0.f, 0.f, 0.f};          // x_{j-3}, x_{j-2}, x_{j-1}
    float dcw[3] = {0.f, 0.f, 0.f};         // dc_{j-3}, dc_{j-2}, dc_{j-1}
    float dw[kW] = {0.f, 0.f, 0.f, 0.f}, db = 0.f;
    for (int j = 0; j < L + 3; ++j) {
        float xn = 0.f, dc = 0.f;
        if (j < L) {
            const int src = ord ? __ldg(ord + j) : j;
            xn = to_f32<T>(x_base[static_cast<int64_t>(src) * G.xz_ts]);
            float pre = bias;
            pre = fmaf(w[0], xw[0], pre); pre = fmaf(w[1], xw[1], pre); pre = fmaf(w[2], xw[2], pre); pre = fmaf(w[3], xn, pre);
            const float s = sigmoid_fast(pre);
            dc = du[static_cast<int64_t>(j) * D] * s * fmaf(pre, 1.0f - s, 1.0f);
            dw[0] = fmaf(dc, xw[0], dw[0]); dw[1] = fmaf(dc, xw[1], dw[1]); dw[2] = fmaf(dc, xw[2], dw[2]);
            dw[3] = fmaf(dc, xn, dw[3]);
            db += dc;
        }
        // dx_{j-3} = dc_{j-3} w3 + dc_{j-2} w2 + dc_{j-1} w1 + dc_j w0
        if (j >= 3) dx_out[static_cast<int64_t>(j - 3) * 2 * D] = fmaf(dcw[0], w[3], fmaf(dcw[1], w[2], fmaf(dcw[2], w[1], dc * w[0])));
        xw[0] = xw[1]; xw[1] = xw[2]; xw[2] = xn;
        dcw[0] = dcw[1]; dcw[1] = dcw[2]; dcw[2] = dc;
    }
#pragma unroll
    for (int t = 0; t < kW; ++t) atomicAdd(G.d_conv_w + static_cast<int64_t>(c) * kW + t, dw[t]);
    if (G.d_conv_b) atomicAdd(G.d_conv_b + c, db);
}

}  // namespace
}  // namespace dm

extern "C" int dm_mamba1_bwd_chunk_tokens(void) { return dm::kCH; }

extern "C" int dm_mamba1_scan_bwd(const dm_mamba1_args* a, const dm_mamba1_bwd_group* gr, int phase, void* stream) {
    using namespace dm;
    if (a == nullptr || gr == nullptr) return DM_ERR_INVALID_ARG;
    if (phase != 1 && phase != 2) return DM_ERR_INVALID_ARG;
    if (a->batch <= 0 || a->n_dir <= 0 || a->seqlen <= 0 || a->n_groups <= 0 || a->n_groups > DM_MAX_GROUPS)
        return DM_ERR_INVALID_ARG;
    if (a->act_dtype != DM_F32 && a->act_dtype != DM_BF16) return DM_ERR_UNSUPPORTED;
    if (a->d_state != kN || a->d_conv != kW || a->dt_rank != kR || a->d_inner % 128 != 0) return DM_ERR_UNSUPPORTED;
    B1P p{};
    p.B = a->batch; p.K = a->n_dir; p.L = a->seqlen; p.D = a->d_inner;
    p.out_order = a->out_order; p.n_groups = a->n_groups; p.order = a->order;
    for (int g = 0; g < a->n_groups; ++g) {
        const dm_mamba1_group& s = a->group[g];
        const dm_mamba1_bwd_group& r = gr[g];
        if (!s.xz || !s.u || !s.x_dbl || !s.dt_proj_weight || !s.A || !s.conv_weight || !r.d_xz_scan) return DM_ERR_INVALID_ARG;
        B1G& d = p.g[g];
        d.xz = s.xz; d.xz_bs = s.xz_batch_stride; d.xz_ts = s.xz_token_stride;
        d.dout = r.dout; d.do_bs = s.out_batch_stride; d.do_ds = s.out_dir_stride; d.do_ts = s.out_token_stride;
        d.u = s.u; d.x_dbl = s.x_dbl; d.wdt = s.dt_proj_weight; d.dt_bias = s.dt_bias; d.A = s.A; d.D = s.D;
        d.d_xz_scan = r.d_xz_scan; d.du = r.du; d.ddelta = r.ddelta; d.d_x_dbl = r.d_x_dbl; d.dA = r.dA; d.dD = r.dD;
        d.d_dt_bias = r.d_dt_bias; d.hb = r.state_workspace; d.states_valid = r.states_valid != 0;
        d.conv_w = s.conv_weight; d.conv_b = s.conv_bias; d.d_conv_w = r.d_conv_weight; d.d_conv_b = r.d_conv_bias;
        d.du_total = r.du;
        if (phase == 1 && (!r.dout || !r.du || !r.ddelta || !r.d_x_dbl || !r.dA || !r.state_workspace)) return DM_ERR_INVALID_ARG;
        if (phase == 2 && (!r.du || !r.d_conv_weight)) return DM_ERR_INVALID_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_seq = p.n_groups * p.B * p.K;
    if (phase == 1) {
        const int n_units = n_seq * (p.D / 32);
        int dev = 0, n_sm = 0;
        if (int e = current_device(&dev, &n_sm); e != DM_OK) return e;
        if (a->act_dtype == DM_F32) {
            const size_t bytes = sizeof(BwdSmem<float>);
            static PerDeviceOnce cfg;
            if (!cfg.done(dev)) {
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_bwd_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                cfg.set(dev);
            }
            m1_scan_bwd_kernel<float><<<n_units, 32, bytes, st>>>(p, n_units);
        } else {
            const size_t bytes = sizeof(BwdSmem<__nv_bfloat16>);
            static PerDeviceOnce cfg;
            if (!cfg.done(dev)) {
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                DM_CUDA_TRY(cudaFuncSetAttribute(m1_scan_bwd_kernel<__nv_bfloat16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                cfg.set(dev);
            }
            m1_scan_bwd_kernel<__nv_bfloat16><<<n_units, 32, bytes, st>>>(p, n_units);
        }
    } else {
        const int grid = n_seq * (p.D / 128);
        if (a->act_dtype == DM_F32) m1_conv_bwd_kernel<float><<<grid, 128, 0, st>>>(p);
        else m1_conv_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(p);
    }
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}
