// Mamba-2 forward hot path for sm_100a (reference call sites
// block/mamba2.py:392-696 and the CrossScan/CrossMerge gathers block/mamba2.py:31-81).
//
// One kernel does everything between the in-projection and the RMSNorm scale:
//   split [z | x | B | C | dt] -> causal conv1d + SiLU over x, B, C -> dt = softplus(dt + bias),
//   dA = exp(dt A_h) -> state recurrence S = dA S + dt x (x) B ; y = S C + D x -> v = y silu(z)
//   -> v (act dtype) to its un-permuted row, sum_c v^2 per token (for the gated RMSNorm) by one atomic per
//   warp and token.
// One WARP per (sequence, 32 channels of one head), lane = channel, the 16 states of the channel in
// registers.  Unlike Mamba-1 there is no cross-channel reduction before the recurrence (B, C, dt come
// straight out of the in-projection), so the whole op is one launch; the decay is one scalar per
// (head, token), so the MUFU load is ~6 per (b,d,l) instead of Mamba-1's 20 and the kernel is bound by
// the FP32 pipe.  That sequential kernel (m2_ssd_kernel) serves fp32 I/O and odd head sizes; bf16 with headdim 64 --
// every BASELINE config -- runs the chunked "SSD" tensor-core form m2_ssd_chunk_kernel (dm_mamba2_chunk.cuh).
#include <cstdlib>

#include "dm_common.cuh"
#include "dm_mamba2_chunk.cuh"

namespace dm {
namespace {

constexpr int kN = 16;
constexpr int kW = 4;
constexpr int kCH = 8;

struct M2G {
    const void* in;
    int64_t in_bs, in_ts;
    void* out;
    int64_t out_bs, out_ds, out_ts;
    float* sumsq;
    int64_t ss_bs, ss_ds;
    const float* conv_w;
    const float* conv_b;
    const float* dt_bias;
    const float* A;
    const float* D;
};
struct M2P {
    int B, K, L, D, H, P;
    int out_order, n_groups, gate;
    const int32_t* order;
    M2G g[DM_MAX_GROUPS];
};

constexpr int kSC = 64;       // channels per warp (one head when headdim = 64): lane owns channels c0+lane, c0+32+lane

template <typename T> struct SsdSmem {
    T xs[2][kCH][kSC];
    T zs[2][kCH][kSC];
    T bcs[2][kCH][32];      // raw [B | C] rows of the chunk
    float bc[kCH][32];      // conv + SiLU of them
    float dt[kCH], dA[kCH];
    int rows[2][kCH];       // output row index of each scanned token
};

__device__ __forceinline__ const int32_t* dir_order(const M2P& p, int k) {
    if (p.order == nullptr) return nullptr;
    const int32_t* o = p.order + static_cast<int64_t>(k) * p.L;
    return (__ldg(o) < 0) ? nullptr : o;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// One WARP per (sequence, 64 channels inside one head); lane owns 2 channels, 2 x 16 states as packed fp32 pairs.
template <typename T>
__global__ void __launch_bounds__(32, 12) m2_ssd_kernel(const __grid_constant__ M2P p, int n_units) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x;
    const int unit = blockIdx.x;
    if (unit >= n_units) return;
    SsdSmem<T>& S = *reinterpret_cast<SsdSmem<T>*>(smem_raw);

    const int D = p.D, L = p.L;
    const int slices = D / kSC;
    const int cs = unit % slices, seq = unit / slices;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M2G& G = p.g[g];
    const int c0 = cs * kSC, head = c0 / p.P;
    const int32_t* ord = dir_order(p, k);
    const T* in_base = static_cast<const T*>(G.in) + static_cast<int64_t>(b) * G.in_bs;
    T* out_base = static_cast<T*>(G.out) + static_cast<int64_t>(b) * G.out_bs + static_cast<int64_t>(k) * G.out_ds + c0 + lane;
    float* ssq = G.sumsq ? G.sumsq + static_cast<int64_t>(b) * G.ss_bs + static_cast<int64_t>(k) * G.ss_ds : nullptr;
    const bool token_order = p.out_order == DM_OUT_TOKEN_ORDER;

    // conv taps: own two x channels and own B|C channel (conv channel D + lane)
    float wx[2][kW], bx[2], wb[kW];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c0 + ch * 32 + lane) * kW));
        wx[ch][0] = t.x; wx[ch][1] = t.y; wx[ch][2] = t.z; wx[ch][3] = t.w;
        bx[ch] = G.conv_b ? __ldg(G.conv_b + c0 + ch * 32 + lane) : 0.f;
    }
    {
        const float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(D + lane) * kW));
        wb[0] = t.x; wb[1] = t.y; wb[2] = t.z; wb[3] = t.w;
    }
    const float bb = G.conv_b ? __ldg(G.conv_b + D + lane) : 0.f;
    const float A2 = __ldg(G.A + head) * kLog2e;
    const float dtb = G.dt_bias ? __ldg(G.dt_bias + head) : 0.f;
    const float Dh = G.D ? __ldg(G.D + head) : 0.f;
    const int dt_off = 2 * D + 2 * kN + head;

    constexpr int kSeg = kSC * sizeof(T) / 16;      // 16-byte segments per (token, 64 channels): 8 / 16
    constexpr int kSegB = 32 * sizeof(T) / 16;      // per (token, B|C row): 4 / 8
    float dt_next = 0.f;
    auto prefetch = [&](int ci) {
        const int buf = ci & 1, j0 = ci * kCH;
        const uint32_t xdst = smem_u32(&S.xs[buf][0][0]), zdst = smem_u32(&S.zs[buf][0][0]),
                       bdst = smem_u32(&S.bcs[buf][0][0]);
#pragma unroll
        for (int i = 0; i < kSeg * kCH / 32; ++i) {
            const int s = lane + 32 * i;
            const int r = s / kSeg, part = s % kSeg;
            const int j = min(j0 + r, L - 1);
            const int src = ord ? __ldg(ord + j) : j;
            if (part == 0) S.rows[buf][r] = token_order ? src : j;
            const char* row = reinterpret_cast<const char*>(in_base + static_cast<int64_t>(src) * G.in_ts);
            cp_async16(zdst + s * 16, row + static_cast<size_t>(c0) * sizeof(T) + part * 16);
            cp_async16(xdst + s * 16, row + static_cast<size_t>(D + c0) * sizeof(T) + part * 16);
        }
#pragma unroll
        for (int i = 0; i < (kSegB * kCH + 31) / 32; ++i) {
            const int s = lane + 32 * i;
            const int r = s / kSegB, part = s % kSegB;
            if (r < kCH) {
                const int j = min(j0 + r, L - 1);
                const int src = ord ? __ldg(ord + j) : j;
                const char* row = reinterpret_cast<const char*>(in_base + static_cast<int64_t>(src) * G.in_ts);
                cp_async16(bdst + s * 16, row + static_cast<size_t>(2 * D) * sizeof(T) + part * 16);
            }
        }
        cp_async_commit();
        if (lane < kCH) {                         // raw dt of (token j0+lane, own head): one scalar per lane
            const int j = min(j0 + lane, L - 1);
            const int src = ord ? __ldg(ord + j) : j;
            dt_next = to_f32<T>(in_base[static_cast<int64_t>(src) * G.in_ts + dt_off]);
        }
    };

    uint64_t h[2][kN / 2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int n = 0; n < kN / 2; ++n) h[ch][n] = 0ull;
    float winx[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, winb[3] = {0.f, 0.f, 0.f};

    auto token = [&](int buf, int jj) {
        const ulonglong2* bc = reinterpret_cast<const ulonglong2*>(&S.bc[jj][0]);
        ulonglong2 Bq[4], Cq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { Bq[q] = bc[q]; Cq[q] = bc[4 + q]; }
        const float dA = S.dA[jj], dt = S.dt[jj];
        const uint64_t dA2 = pack2(dA, dA);
        const int row = S.rows[buf][jj];
        float s2 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const float xr = to_f32<T>(S.xs[buf][jj][ch * 32 + lane]);
            float acc = bx[ch];
            acc = fmaf(wx[ch][0], winx[ch][0], acc);
            acc = fmaf(wx[ch][1], winx[ch][1], acc);
            acc = fmaf(wx[ch][2], winx[ch][2], acc);
            acc = fmaf(wx[ch][3], xr, acc);
            const float xv = silu_fast(acc);
            winx[ch][0] = winx[ch][1]; winx[ch][1] = winx[ch][2]; winx[ch][2] = xr;
            const float dtx = dt * xv;
            const uint64_t dtx2 = pack2(dtx, dtx);
            uint64_t y0 = 0ull, y1 = 0ull;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                h[ch][2 * q] = fma2(dA2, h[ch][2 * q], mul2(dtx2, Bq[q].x));
                y0 = fma2(h[ch][2 * q], Cq[q].x, y0);
                h[ch][2 * q + 1] = fma2(dA2, h[ch][2 * q + 1], mul2(dtx2, Bq[q].y));
                y1 = fma2(h[ch][2 * q + 1], Cq[q].y, y1);
            }
            float ya, yb, yc, yd;
            unpack2(y0, ya, yb);
            unpack2(y1, yc, yd);
            float v = fmaf(Dh, xv, (ya + yb) + (yc + yd));
            if (p.gate) v *= silu_fast(to_f32<T>(S.zs[buf][jj][ch * 32 + lane]));
            out_base[static_cast<int64_t>(row) * G.out_ts + ch * 32] = from_f32<T>(v);
            s2 = fmaf(v, v, s2);
        }
        if (ssq) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (lane == 0) atomicAdd(ssq + row, s2);
        }
    };

    const int n_chunks = (L + kCH - 1) / kCH;
    prefetch(0);
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int buf = ci & 1, j0 = ci * kCH;
        const float dt_raw = dt_next;
        if (ci + 1 < n_chunks) {
            prefetch(ci + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        // ---- per-chunk prologue: conv + SiLU of B|C (lane = B|C channel), dt / decay (lane = token) ----
#pragma unroll
        for (int jj = 0; jj < kCH; ++jj) {
            const float v = to_f32<T>(S.bcs[buf][jj][lane]);
            float acc = bb;
            acc = fmaf(wb[0], winb[0], acc);
            acc = fmaf(wb[1], winb[1], acc);
            acc = fmaf(wb[2], winb[2], acc);
            acc = fmaf(wb[3], v, acc);
            S.bc[jj][lane] = silu_fast(acc);
            winb[0] = winb[1]; winb[1] = winb[2]; winb[2] = v;
        }
        if (lane < kCH) {
            const float dt = softplus_f(dt_raw + dtb);
            S.dt[lane] = dt;
            S.dA[lane] = ex2_approx(dt * A2);
        }
        __syncwarp();

        if (j0 + kCH <= L) {
#pragma unroll
            for (int jj = 0; jj < kCH; ++jj) token(buf, jj);
        } else {
            const int nrows = L - j0;
#pragma unroll 1
            for (int jj = 0; jj < nrows; ++jj) token(buf, jj);
        }
        __syncwarp();
    }
}

// ---- tensor-core chunked SSD (bf16, headdim 64, d_state 16); see dm_mamba2_chunk.cuh -------------------------------
__global__ void __launch_bounds__(ssd::kThreads, 4) m2_ssd_chunk_kernel(const __grid_constant__ M2P p) {
    using namespace ssd;
    using T = __nv_bfloat16;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, r4 = lane >> 2;
    const int H = p.H, D = p.D, L = p.L;
    const int head = blockIdx.x % H, seq = blockIdx.x / H;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M2G& G = p.g[g];
    const int32_t* ord = dir_order(p, k);
    const T* in_base = static_cast<const T*>(G.in) + static_cast<int64_t>(b) * G.in_bs;
    T* out_base = static_cast<T*>(G.out) + static_cast<int64_t>(b) * G.out_bs + static_cast<int64_t>(k) * G.out_ds + head * P;
    float* ssq = G.sumsq ? G.sumsq + static_cast<int64_t>(b) * G.ss_bs + static_cast<int64_t>(k) * G.ss_ds : nullptr;
    const bool token_order = p.out_order == DM_OUT_TOKEN_ORDER;
    const float A2 = __ldg(G.A + head) * kLog2e;
    const float dtb = G.dt_bias ? __ldg(G.dt_bias + head) : 0.f;
    const float Dh = G.D ? __ldg(G.D + head) : 0.f;
    const int dt_off = 2 * D + 2 * NS + head;

    // conv taps of this thread: x channel pair (2*(tid%32)) for 16 tokens, B|C channel pair (2*(tid%16)) for 8 tokens
    const int xc = 2 * (tid & 31), xg = tid >> 5;
    const int bc = 2 * (tid & 15), bg = tid >> 4;
    float wx[2][kW], bx[2], wb[2][kW], bb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(head * P + xc + i) * kW));
        wx[i][0] = t.x; wx[i][1] = t.y; wx[i][2] = t.z; wx[i][3] = t.w;
        bx[i] = G.conv_b ? __ldg(G.conv_b + head * P + xc + i) : 0.f;
        t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(D + bc + i) * kW));
        wb[i][0] = t.x; wb[i][1] = t.y; wb[i][2] = t.z; wb[i][3] = t.w;
        bb[i] = G.conv_b ? __ldg(G.conv_b + D + bc + i) : 0.f;
    }

    float stacc[2][4];                               // state[n][p] of this warp's 16 channels, fp32
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) stacc[i][e] = 0.f;

    const int n_chunks = (L + Q - 1) / Q;
    for (int c = 0; c < n_chunks; ++c) {
        const int j0 = c * Q, nv = min(Q, L - j0);
        // ---- (a) gather the chunk's rows (scan order) ----
        for (int s = tid; s < (Q + 3) * 8; s += kThreads) {
            const int r = s >> 3, part = s & 7, j = j0 - 3 + r;
            T* d = &S.xraw[r][part * 8];
            if (j < 0 || j >= L) {
                *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                const int src = ord ? __ldg(ord + j) : j;
                cp_async16(smem_u32(d), in_base + static_cast<int64_t>(src) * G.in_ts + D + head * P + part * 8);
            }
        }
        for (int s = tid; s < (Q + 3) * 4; s += kThreads) {
            const int r = s >> 2, part = s & 3, j = j0 - 3 + r;
            T* d = &S.bcraw[r][part * 8];
            if (j < 0 || j >= L) {
                *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                const int src = ord ? __ldg(ord + j) : j;
                cp_async16(smem_u32(d), in_base + static_cast<int64_t>(src) * G.in_ts + 2 * D + part * 8);
            }
        }
        for (int s = tid; s < Q * 8; s += kThreads) {
            const int r = s >> 3, part = s & 7, j = j0 + r;
            T* d = &S.zs[r][part * 8];
            if (j >= L) {
                *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
                if (part == 0) S.rows[r] = 0;
            } else {
                const int src = ord ? __ldg(ord + j) : j;
                if (part == 0) S.rows[r] = token_order ? src : j;
                cp_async16(smem_u32(d), in_base + static_cast<int64_t>(src) * G.in_ts + head * P + part * 8);
            }
        }
        float dt_raw[2] = {0.f, 0.f};
        bool dt_ok[2] = {false, false};
        if (tid < 32) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int j = j0 + 2 * tid + i;
                if (j < L) {
                    const int src = ord ? __ldg(ord + j) : j;
                    dt_raw[i] = to_f32<T>(in_base[static_cast<int64_t>(src) * G.in_ts + dt_off]);
                    dt_ok[i] = true;
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();

        // ---- (b) conv + SiLU -> X, B, C ; dt, cumulative log-decay ----
        {
            float w0[2], w1[2], w2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                w0[i] = to_f32<T>(S.xraw[16 * xg + 0][xc + i]);
                w1[i] = to_f32<T>(S.xraw[16 * xg + 1][xc + i]);
                w2[i] = to_f32<T>(S.xraw[16 * xg + 2][xc + i]);
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int jj = 16 * xg + t;
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&S.xraw[jj + 3][xc]);
                const float xn[2] = {__low2float(v), __high2float(v)};
                float o[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float acc = bx[i];
                    acc = fmaf(wx[i][0], w0[i], acc); acc = fmaf(wx[i][1], w1[i], acc);
                    acc = fmaf(wx[i][2], w2[i], acc); acc = fmaf(wx[i][3], xn[i], acc);
                    o[i] = silu_t(acc);
                    w0[i] = w1[i]; w1[i] = w2[i]; w2[i] = xn[i];
                }
                *reinterpret_cast<uint32_t*>(&S.xs[jj][xc]) = pk(o[0], o[1]);
            }
        }
        {
            float w0[2], w1[2], w2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                w0[i] = to_f32<T>(S.bcraw[8 * bg + 0][bc + i]);
                w1[i] = to_f32<T>(S.bcraw[8 * bg + 1][bc + i]);
                w2[i] = to_f32<T>(S.bcraw[8 * bg + 2][bc + i]);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int jj = 8 * bg + t;
                const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&S.bcraw[jj + 3][bc]);
                const float xn[2] = {__low2float(v), __high2float(v)};
                float o[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float acc = bb[i];
                    acc = fmaf(wb[i][0], w0[i], acc); acc = fmaf(wb[i][1], w1[i], acc);
                    acc = fmaf(wb[i][2], w2[i], acc); acc = fmaf(wb[i][3], xn[i], acc);
                    o[i] = silu_fast(acc);
                    w0[i] = w1[i]; w1[i] = w2[i]; w2[i] = xn[i];
                }
                if (bc < NS) *reinterpret_cast<uint32_t*>(&S.bs[jj][bc]) = pk(o[0], o[1]);
                else *reinterpret_cast<uint32_t*>(&S.cs[jj][bc - NS]) = pk(o[0], o[1]);
            }
        }
        if (tid < 32) {          // tokens 2*tid, 2*tid+1: dt, inclusive cumsum of dt*A*log2(e)
            const float d0 = dt_ok[0] ? softplus_f(dt_raw[0] + dtb) : 0.f;
            const float d1 = dt_ok[1] ? softplus_f(dt_raw[1] + dtb) : 0.f;
            const float s0 = d0 * A2, s1 = s0 + d1 * A2;
            float incl = s1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const float excl = incl - s1;
            S.cum[2 * tid] = excl + s0; S.cum[2 * tid + 1] = excl + s1;
            S.dtv[2 * tid] = d0; S.dtv[2 * tid + 1] = d1;
        }
        __syncthreads();

        // ---- (c)-(g) this warp's 16 chunk rows ----
        {
            const int i0 = 16 * warp + r4, i1 = i0 + 8;
            const float ci0 = S.cum[i0], ci1 = S.cum[i1];
            uint32_t a_c[4];
            ldsm_x4(a_c, smem_u32(&S.cs[16 * warp + (lane & 15)][(lane >> 4) * 8]));
            float y[8][4];
#pragma unroll
            for (int pt = 0; pt < 8; ++pt)
#pragma unroll
                for (int e = 0; e < 4; ++e) y[pt][e] = 0.f;
            for (int kk = 0; kk <= warp; ++kk) {
                uint32_t bfr[4];
                ldsm_x4(bfr, smem_u32(&S.bs[16 * kk + (lane & 7) + ((lane >> 4) & 1) * 8][((lane >> 3) & 1) * 8]));
                float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
                mma16816(s0, a_c, bfr[0], bfr[1]);
                mma16816(s1, a_c, bfr[2], bfr[3]);
                const int ja = 16 * kk + 2 * q, jb = ja + 8;
                const float2 cja = *reinterpret_cast<const float2*>(&S.cum[ja]), cjb = *reinterpret_cast<const float2*>(&S.cum[jb]);
                const float2 dja = *reinterpret_cast<const float2*>(&S.dtv[ja]), djb = *reinterpret_cast<const float2*>(&S.dtv[jb]);
                auto mk = [&](float s, float ci, float cj, float dj, int i, int j) {
                    return (j <= i) ? s * ex2_approx(ci - cj) * dj : 0.f;
                };
                uint32_t am[4];
                am[0] = pk(mk(s0[0], ci0, cja.x, dja.x, i0, ja), mk(s0[1], ci0, cja.y, dja.y, i0, ja + 1));
                am[1] = pk(mk(s0[2], ci1, cja.x, dja.x, i1, ja), mk(s0[3], ci1, cja.y, dja.y, i1, ja + 1));
                am[2] = pk(mk(s1[0], ci0, cjb.x, djb.x, i0, jb), mk(s1[1], ci0, cjb.y, djb.y, i0, jb + 1));
                am[3] = pk(mk(s1[2], ci1, cjb.x, djb.x, i1, jb), mk(s1[3], ci1, cjb.y, djb.y, i1, jb + 1));
#pragma unroll
                for (int pt2 = 0; pt2 < 4; ++pt2) {
                    uint32_t xb[4];
                    ldsm_x4_t(xb, smem_u32(&S.xs[16 * kk + (lane & 15)][pt2 * 16 + (lane >> 4) * 8]));
                    mma16816(y[2 * pt2], am, xb[0], xb[1]);
                    mma16816(y[2 * pt2 + 1], am, xb[2], xb[3]);
                }
            }
            if (c > 0) {                              // contribution of the state carried in from earlier chunks
                const float e0 = ex2_approx(ci0), e1 = ex2_approx(ci1);
#pragma unroll
                for (int pt2 = 0; pt2 < 4; ++pt2) {
                    uint32_t sb[4];
                    ldsm_x4_t(sb, smem_u32(&S.st[lane & 15][pt2 * 16 + (lane >> 4) * 8]));
                    float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
                    mma16816(t0, a_c, sb[0], sb[1]);
                    mma16816(t1, a_c, sb[2], sb[3]);
                    y[2 * pt2][0] = fmaf(e0, t0[0], y[2 * pt2][0]); y[2 * pt2][1] = fmaf(e0, t0[1], y[2 * pt2][1]);
                    y[2 * pt2][2] = fmaf(e1, t0[2], y[2 * pt2][2]); y[2 * pt2][3] = fmaf(e1, t0[3], y[2 * pt2][3]);
                    y[2 * pt2 + 1][0] = fmaf(e0, t1[0], y[2 * pt2 + 1][0]); y[2 * pt2 + 1][1] = fmaf(e0, t1[1], y[2 * pt2 + 1][1]);
                    y[2 * pt2 + 1][2] = fmaf(e1, t1[2], y[2 * pt2 + 1][2]); y[2 * pt2 + 1][3] = fmaf(e1, t1[3], y[2 * pt2 + 1][3]);
                }
            }
            // epilogue: D skip, gate, store, sum of squares
            float sq0 = 0.f, sq1 = 0.f;
            const int64_t ro0 = static_cast<int64_t>(S.rows[i0]) * G.out_ts, ro1 = static_cast<int64_t>(S.rows[i1]) * G.out_ts;
#pragma unroll
            for (int pt = 0; pt < 8; ++pt) {
                const int pc = pt * 8 + 2 * q;
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int i = hrow ? i1 : i0;
                    const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&S.xs[i][pc]);
                    float v0 = fmaf(Dh, __low2float(xv), y[pt][2 * hrow]), v1 = fmaf(Dh, __high2float(xv), y[pt][2 * hrow + 1]);
                    if (p.gate) {
                        const __nv_bfloat162 zv = *reinterpret_cast<const __nv_bfloat162*>(&S.zs[i][pc]);
                        v0 *= silu_t(__low2float(zv));
                        v1 *= silu_t(__high2float(zv));
                    }
                    if (i < nv) {
                        *reinterpret_cast<uint32_t*>(out_base + (hrow ? ro1 : ro0) + pc) = pk(v0, v1);
                        if (hrow) sq1 = fmaf(v0, v0, fmaf(v1, v1, sq1)); else sq0 = fmaf(v0, v0, fmaf(v1, v1, sq0));
                    }
                }
            }
            if (ssq) {
                sq0 += __shfl_xor_sync(0xffffffffu, sq0, 1); sq0 += __shfl_xor_sync(0xffffffffu, sq0, 2);
                sq1 += __shfl_xor_sync(0xffffffffu, sq1, 1); sq1 += __shfl_xor_sync(0xffffffffu, sq1, 2);
                if (q == 0) {
                    if (i0 < nv) atomicAdd(ssq + S.rows[i0], sq0);
                    if (i1 < nv) atomicAdd(ssq + S.rows[i1], sq1);
                }
            }
        }
        if (c + 1 == n_chunks) break;
        __syncthreads();                                // every warp is done reading the old state copy

        // ---- (h) state update: State = exp(cum_last) State + (B o w)^T X, warp owns channels [16w, 16w+16) ----
        {
            const float wl = S.cum[Q - 1];
            float sn[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) sn[i][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < Q / 16; ++kk) {
                uint32_t ab[4];
                ldsm_x4_t(ab, smem_u32(&S.bs[16 * kk + (lane & 7) + ((lane >> 4) & 1) * 8][((lane >> 3) & 1) * 8]));
                const int ja = 16 * kk + 2 * q, jb = ja + 8;
                const float2 cja = *reinterpret_cast<const float2*>(&S.cum[ja]), cjb = *reinterpret_cast<const float2*>(&S.cum[jb]);
                const float2 dja = *reinterpret_cast<const float2*>(&S.dtv[ja]), djb = *reinterpret_cast<const float2*>(&S.dtv[jb]);
                const __nv_bfloat162 wa = __floats2bfloat162_rn(ex2_approx(wl - cja.x) * dja.x, ex2_approx(wl - cja.y) * dja.y);
                const __nv_bfloat162 wbv = __floats2bfloat162_rn(ex2_approx(wl - cjb.x) * djb.x, ex2_approx(wl - cjb.y) * djb.y);
                auto scale = [](uint32_t v, __nv_bfloat162 w) {
                    __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&v), w);
                    return *reinterpret_cast<uint32_t*>(&r);
                };
                ab[0] = scale(ab[0], wa); ab[1] = scale(ab[1], wa); ab[2] = scale(ab[2], wbv); ab[3] = scale(ab[3], wbv);
                uint32_t xb[4];
                ldsm_x4_t(xb, smem_u32(&S.xs[16 * kk + (lane & 15)][16 * warp + (lane >> 4) * 8]));
                mma16816(sn[0], ab, xb[0], xb[1]);
                mma16816(sn[1], ab, xb[2], xb[3]);
            }
            const float el = ex2_approx(wl);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int e = 0; e < 4; ++e) stacc[i][e] = fmaf(el, stacc[i][e], sn[i][e]);
                const int pc = 16 * warp + i * 8 + 2 * q;
                *reinterpret_cast<uint32_t*>(&S.st[r4][pc]) = pk(stacc[i][0], stacc[i][1]);
                *reinterpret_cast<uint32_t*>(&S.st[r4 + 8][pc]) = pk(stacc[i][2], stacc[i][3]);
            }
        }
        __syncthreads();                                // new state visible; chunk buffers free for the next gather
    }
}

template <typename T>
int launch_m2(const M2P& p, cudaStream_t stream) {
    int dev = 0, n_sm = 0;
    if (int e = current_device(&dev, &n_sm); e != DM_OK) return e;
    if constexpr (sizeof(T) == 2) {
        // bf16, headdim 64: the chunked tensor-core form (DM_M2_CHUNK=0 forces the sequential kernel)
        static const int use_chunk = env_int("DM_M2_CHUNK", 1);
        if (use_chunk && p.P == ssd::P) {
            static PerDeviceOnce cfg;
            if (!cfg.done(dev)) {
                DM_CUDA_TRY(cudaFuncSetAttribute(m2_ssd_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(sizeof(ssd::Smem))));
                DM_CUDA_TRY(cudaFuncSetAttribute(m2_ssd_chunk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                cfg.set(dev);
            }
            const int units = p.n_groups * p.B * p.K * p.H;
            m2_ssd_chunk_kernel<<<units, ssd::kThreads, sizeof(ssd::Smem), stream>>>(p);
            DM_CUDA_TRY(cudaGetLastError());
            return DM_OK;
        }
    }
    const int n_units = p.n_groups * p.B * p.K * (p.D / kSC);
    const size_t bytes = sizeof(SsdSmem<T>);
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        DM_CUDA_TRY(cudaFuncSetAttribute(m2_ssd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(bytes)));
        DM_CUDA_TRY(cudaFuncSetAttribute(m2_ssd_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured.set(dev);
    }
    m2_ssd_kernel<T><<<n_units, 32, bytes, stream>>>(p, n_units);
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}

}  // namespace
}  // namespace dm

extern "C" int dm_mamba2_ssd_fwd(const dm_mamba2_args* a, void* stream) {
    using namespace dm;
    if (a == nullptr) return DM_ERR_INVALID_ARG;
    if (a->batch <= 0 || a->n_dir <= 0 || a->seqlen <= 0 || a->n_groups <= 0 || a->n_groups > DM_MAX_GROUPS)
        return DM_ERR_INVALID_ARG;
    if (a->out_order != DM_OUT_SCAN_ORDER && a->out_order != DM_OUT_TOKEN_ORDER) return DM_ERR_INVALID_ARG;
    if (a->act_dtype != DM_F32 && a->act_dtype != DM_BF16) return DM_ERR_UNSUPPORTED;
    if (a->d_state != kN || a->d_conv != kW) return DM_ERR_UNSUPPORTED;
    if (a->nheads <= 0 || a->d_inner <= 0 || a->d_inner % a->nheads != 0) return DM_ERR_INVALID_ARG;
    const int P = a->d_inner / a->nheads;
    if (P % 64 != 0) return DM_ERR_UNSUPPORTED;      // a warp owns 64 channels of one head
    const size_t es = dtype_size(a->act_dtype);
    if ((static_cast<size_t>(2 * a->d_inner) * es) % 16 != 0) return DM_ERR_UNSUPPORTED;
    M2P p{};
    p.B = a->batch; p.K = a->n_dir; p.L = a->seqlen; p.D = a->d_inner; p.H = a->nheads; p.P = P;
    p.out_order = a->out_order; p.n_groups = a->n_groups; p.gate = a->gate ? 1 : 0;
    p.order = a->order;
    for (int g = 0; g < a->n_groups; ++g) {
        const dm_mamba2_group& s = a->group[g];
        if (!s.zxbcdt || !s.out || !s.conv_weight || !s.A) return DM_ERR_INVALID_ARG;
        if (!aligned16(s.zxbcdt) || !aligned16(s.conv_weight)) return DM_ERR_INVALID_ARG;
        if ((s.in_batch_stride * es) % 16 || (s.in_token_stride * es) % 16) return DM_ERR_INVALID_ARG;
        if (s.in_token_stride < 2 * a->d_inner + 2 * kN + a->nheads) return DM_ERR_INVALID_ARG;
        M2G& d = p.g[g];
        d.in = s.zxbcdt; d.in_bs = s.in_batch_stride; d.in_ts = s.in_token_stride;
        d.out = s.out; d.out_bs = s.out_batch_stride; d.out_ds = s.out_dir_stride; d.out_ts = s.out_token_stride;
        d.sumsq = s.sumsq; d.ss_bs = s.sumsq_batch_stride; d.ss_ds = s.sumsq_dir_stride;
        d.conv_w = s.conv_weight; d.conv_b = s.conv_bias; d.dt_bias = s.dt_bias; d.A = s.A; d.D = s.D;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return a->act_dtype == DM_F32 ? launch_m2<float>(p, st) : launch_m2<__nv_bfloat16>(p, st);
}
