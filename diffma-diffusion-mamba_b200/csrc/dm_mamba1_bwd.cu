// Mamba-1 backward for sm_100a (upstream MambaInnerFn.backward =
// selective_scan_cuda.bwd + causal_conv1d_cuda.causal_conv1d_bwd + einsum weight gradients; reached from the
// reference through train.py:259).  Everything here works in SCAN order; the host un-permutes and sums directions.
//
//   m1_scan_bwd_kernel   one warp per (sequence, 32 channels), lane = channel.
//        sweep 1 (forward): recompute the recurrence, store the state at every chunk boundary (DM_BWD_CH = 4 tokens) in a
//        workspace -- SKIPPED when the training forward already wrote these checkpoints (dm_mamba1_group.chunk_states,
//        states_valid = 1: the normal case under autograd_ops.Mamba1ScanFn)
//        sweep 2 (reverse, chunk by chunk): reload the boundary state, recompute the chunk's states into shared
//        memory, then run the adjoint recurrence  dh_{j-1} = a_j dh_j,  dh_j += dy_j C_j  backwards, producing
//        dz, du (scan part), d(delta_raw) per token, dB/dC (reduced over the warp's 32 channels through a shared
//        transposition, then one atomic per value), and dA / dD / d(dt_bias) accumulated in registers.  Chunk inputs
//        (x_dbl rows, u, z, dout) are staged by cp.async, double buffered, the scan-order lookup one chunk ahead.
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

template <typename T> struct BwdSmem {
    float hs[kCH][kN][32];       // state BEFORE each token of the chunk
    float xd[2][kCH][kE];        // x_dbl rows of the chunk, double buffered (cp.async)
    float ds[kMR][34];           // delta_raw tile (the MMA tile has 8 token rows; rows >= kCH are never read)
    float tb[32][33];            // transposition buffer for the dB / dC reduction over channels
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

    float A2[kN];
#pragma unroll
    for (int n = 0; n < kN; n += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(G.A + static_cast<int64_t>(c) * kN + n));
        A2[n] = t.x * kLog2e; A2[n + 1] = t.y * kLog2e; A2[n + 2] = t.z * kLog2e; A2[n + 3] = t.w * kLog2e;
    }
    const float dtb = G.dt_bias ? __ldg(G.dt_bias + c) : 0.f;
    const float Dc = G.D ? __ldg(G.D + c) : 0.f;

    // W_dt fragments (as in the forward kernel, 32 channels -> 4 n-tiles)
    uint32_t bw_hi[4][2][2], bw_lo[4][2][2];
    {
        const T* Wdt = static_cast<const T*>(G.wdt);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const T* wp = Wdt + static_cast<int64_t>(c0 + nt * 8 + (lane >> 2)) * kR + ks * 16 + 2 * (lane & 3);
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
    }

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
    int src_pref = 0;
    auto load_src = [&](int ci) {
        if (ci >= 0 && lane < kCH * kSegRow) {
            const int j = min(ci * kCH + lane / kSegRow, L - 1);
            src_pref = ord ? __ldg(ord + j) : j;
        }
    };
    auto issue = [&](int ci, int buf, bool with_grad) {
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
                const int src = src_pref;                     // kCH * kSegRow <= 32: this loop runs once per lane
                cp_async16(smem_u32(&S.zs[buf][r][part * kEpS]), z0 + static_cast<int64_t>(src) * G.xz_ts + part * kEpS);
                cp_async16(smem_u32(&S.dos[buf][r][part * kEpS]),
                           do0 + static_cast<int64_t>(token_order ? src : j) * G.do_ts + part * kEpS);
            }
        }
        cp_async_commit();
        if (with_grad) load_src(ci - 1);
    };
    static_assert(kCH * kSegRow <= 32, "one (row, segment) of u / z / dout per lane");
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
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<float2*>(&S.ds[r][nt * 8 + 2 * q]) = make_float2(dacc[nt][0], dacc[nt][1]);
        __syncwarp();
    };

    // ---- sweep 1: forward, boundary states -> workspace ----
    float h[kN];
#pragma unroll
    for (int n = 0; n < kN; ++n) h[n] = 0.f;
    // (skipped when the training forward already stored the checkpoints: G.states_valid)
    if (n_chunks > 1 && !G.states_valid) issue(0, 0, false);
    for (int ci = 0; ci < n_chunks && !G.states_valid; ++ci) {
        float* hbc = hb + static_cast<int64_t>(ci) * D * kN;
#pragma unroll
        for (int n = 0; n < kN; n += 4) *reinterpret_cast<float4*>(hbc + n) = make_float4(h[n], h[n + 1], h[n + 2], h[n + 3]);
        if (ci == n_chunks - 1) break;                       // the last chunk's end state is never needed
        const int buf = ci & 1;
        const bool more = ci + 1 < n_chunks - 1;             // sweep 1 visits chunks 0 .. n_chunks - 2
        __syncwarp();                                        // every lane is done with the buffer the prefetch overwrites
        if (more) issue(ci + 1, buf ^ 1, false);
        finish(buf, more);
#pragma unroll
        for (int jj = 0; jj < kCH; ++jj) {
            const float dt = softplus_fast(S.ds[jj][lane] + dtb);
            const float dtu = dt * to_f32<T>(S.us[buf][jj][lane]);
            const float* Bv = &S.xd[buf][jj][kR];
#pragma unroll
            for (int n = 0; n < kN; ++n) h[n] = fmaf(ex2_approx(dt * A2[n]), h[n], dtu * Bv[n]);
        }
    }

    // ---- sweep 2: reverse ----
    float dh[kN], dA[kN];
#pragma unroll
    for (int n = 0; n < kN; ++n) { dh[n] = 0.f; dA[n] = 0.f; }
    float dD = 0.f, ddtb = 0.f;
    float* dz_out = G.d_xz_scan + sg * L * 2 * D + D + c;
    float* du_out = G.du + sg * L * D + c;
    float* dd_out = G.ddelta + sg * L * D + c;
    float* dxd_out = G.d_x_dbl + sg * L * kE;
    __syncwarp();
    load_src(n_chunks - 1);
    issue(n_chunks - 1, (n_chunks - 1) & 1, true);
    for (int ci = n_chunks - 1; ci >= 0; --ci) {
        const int j0 = ci * kCH, nrows = min(kCH, L - j0);
        const int buf = ci & 1;
        __syncwarp();                                        // every lane is done with the buffer the prefetch overwrites
        if (ci > 0) issue(ci - 1, buf ^ 1, true);
        finish(buf, ci > 0);
        const float* hbc = hb + static_cast<int64_t>(ci) * D * kN;
#pragma unroll
        for (int n = 0; n < kN; n += 4) {
            const float4 t = *reinterpret_cast<const float4*>(hbc + n);
            h[n] = t.x; h[n + 1] = t.y; h[n + 2] = t.z; h[n + 3] = t.w;
        }
        float dtv[kCH], yv[kCH];
        // forward through the chunk: keep the state before each token in shared memory, y in registers
#pragma unroll
        for (int jj = 0; jj < kCH; ++jj) {
            const float dt = softplus_fast(S.ds[jj][lane] + dtb);
            const float uu = to_f32<T>(S.us[buf][jj][lane]);
            const float dtu = dt * uu;
            const float* Bv = &S.xd[buf][jj][kR];
            const float* Cv = &S.xd[buf][jj][kR + kN];
            float y = 0.f;
#pragma unroll
            for (int n = 0; n < kN; ++n) {
                S.hs[jj][n][lane] = h[n];
                h[n] = fmaf(ex2_approx(dt * A2[n]), h[n], dtu * Bv[n]);
                y = fmaf(h[n], Cv[n], y);
            }
            dtv[jj] = dt;
            yv[jj] = fmaf(Dc, uu, y);
        }
        // adjoint recurrence, last token of the chunk first
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
                const float dz = go * yv[jj] * sg_z * fmaf(zz, 1.0f - sg_z, 1.0f);    // silu'(z) = s (1 + z (1 - s))
                dD = fmaf(dy, uu, dD);
                const float* Bv = &S.xd[buf][jj][kR];
                const float* Cv = &S.xd[buf][jj][kR + kN];
                float ddt = 0.f, dbu = 0.f;     // d delta, sum_n dh_n B_n
                float pB[kN], pC[kN];
#pragma unroll
                for (int n = 0; n < kN; ++n) {
                    const float hp = S.hs[jj][n][lane];
                    const float a = ex2_approx(dt * A2[n]);
                    const float hcur = fmaf(a, hp, dt * uu * Bv[n]);               // state after token jj
                    pC[n] = dy * hcur;                                              // dC_j[n] contribution of this channel
                    dh[n] = fmaf(dy, Cv[n], dh[n]);                                 // dL/dh_j
                    const float t = dh[n] * hp * a;
                    dA[n] = fmaf(t, dt, dA[n]);                                     // d/dA: dh * hp * a * dt
                    ddt = fmaf(t, A2[n], ddt);                                      // (scaled by ln2 below) dh * hp * a * A
                    dbu = fmaf(dh[n], Bv[n], dbu);
                    pB[n] = dh[n] * dt * uu;                                        // dB_j[n] contribution
                    dh[n] *= a;                                                     // -> dL/dh_{j-1} through the decay
                }
                ddt = fmaf(ddt, 0.6931471805599453f, dbu * uu);
                const float du = fmaf(dy, Dc, dbu * dt);
                const float ddr = ddt * sigmoid_fast(S.ds[jj][lane] + dtb);          // softplus'(x) = sigmoid(x)
                ddtb += ddr;
                dz_out[static_cast<int64_t>(j) * 2 * D] = dz;
                du_out[static_cast<int64_t>(j) * D] = du;
                dd_out[static_cast<int64_t>(j) * D] = ddr;
                // reduce dB / dC over the warp's 32 channels: transpose through shared memory, lane v sums value v
                __syncwarp();
#pragma unroll
                for (int n = 0; n < kN; ++n) { S.tb[n][lane] = pB[n]; S.tb[kN + n][lane] = pC[n]; }
                __syncwarp();
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) s += S.tb[lane][i];
                atomicAdd(dxd_out + static_cast<int64_t>(j) * kE + kR + lane, s);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < kN; ++n) atomicAdd(G.dA + static_cast<int64_t>(c) * kN + n, dA[n]);
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
