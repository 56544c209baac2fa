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
// the FP32 pipe.  (The chunked "SSD" tensor-core form is the planned next step; see DESIGN.md.)
#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int kN = 16;
constexpr int kW = 4;
constexpr int kCH = 16;
constexpr int kWarps = 2;

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

template <typename T> struct SsdSmem {
    T xs[2][kCH][32];
    T zs[2][kCH][32];
    T bcs[2][kCH][32];      // raw [B | C] rows of the chunk
    float bc[kCH][32];      // conv + SiLU of them
    float dt[kCH], dA[kCH];
    int rows[2][kCH];
};

__device__ __forceinline__ const int32_t* dir_order(const M2P& p, int k) {
    if (p.order == nullptr) return nullptr;
    const int32_t* o = p.order + static_cast<int64_t>(k) * p.L;
    return (__ldg(o) < 0) ? nullptr : o;
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) m2_ssd_kernel(const __grid_constant__ M2P p, int n_units) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * kWarps + warp;
    if (unit >= n_units) return;
    SsdSmem<T>& S = reinterpret_cast<SsdSmem<T>*>(smem_raw)[warp];

    const int D = p.D, L = p.L;
    const int slices = D >> 5;
    const int cs = unit % slices, seq = unit / slices;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M2G& G = p.g[g];
    const int c0 = cs * 32, c = c0 + lane, head = c0 / p.P;
    const int32_t* ord = dir_order(p, k);
    const T* in_base = static_cast<const T*>(G.in) + static_cast<int64_t>(b) * G.in_bs;
    T* out_base = static_cast<T*>(G.out) + static_cast<int64_t>(b) * G.out_bs + static_cast<int64_t>(k) * G.out_ds + c;
    float* ssq = G.sumsq ? G.sumsq + static_cast<int64_t>(b) * G.ss_bs + static_cast<int64_t>(k) * G.ss_ds : nullptr;

    // conv taps: own x channel (conv channel c) and own B|C channel (conv channel D + lane)
    float wx[kW], wb[kW];
    {
        float4 t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c) * kW));
        wx[0] = t.x; wx[1] = t.y; wx[2] = t.z; wx[3] = t.w;
        t = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(D + lane) * kW));
        wb[0] = t.x; wb[1] = t.y; wb[2] = t.z; wb[3] = t.w;
    }
    const float bx = G.conv_b ? __ldg(G.conv_b + c) : 0.f;
    const float bb = G.conv_b ? __ldg(G.conv_b + D + lane) : 0.f;
    const float A2 = __ldg(G.A + head) * kLog2e;
    const float dtb = G.dt_bias ? __ldg(G.dt_bias + head) : 0.f;
    const float Dh = G.D ? __ldg(G.D + head) : 0.f;
    const int dt_off = 2 * D + 2 * kN + head;

    constexpr int kSeg = 32 * sizeof(T) / 16;
    float dt_next = 0.f;
    auto prefetch = [&](int ci) {
        const int buf = ci & 1, j0 = ci * kCH;
        const int nrows = min(kCH, L - j0);
        const uint32_t xdst = smem_u32(&S.xs[buf][0][0]), zdst = smem_u32(&S.zs[buf][0][0]),
                       bdst = smem_u32(&S.bcs[buf][0][0]);
        for (int s = lane; s < nrows * kSeg; s += 32) {
            const int r = s / kSeg, part = s - r * kSeg;
            const int j = j0 + r;
            const int src = ord ? __ldg(ord + j) : j;
            if (part == 0) S.rows[buf][r] = src;
            const char* row = reinterpret_cast<const char*>(in_base + static_cast<int64_t>(src) * G.in_ts);
            cp_async16(zdst + s * 16, row + static_cast<size_t>(c0) * sizeof(T) + part * 16);
            cp_async16(xdst + s * 16, row + static_cast<size_t>(D + c0) * sizeof(T) + part * 16);
            cp_async16(bdst + s * 16, row + static_cast<size_t>(2 * D) * sizeof(T) + part * 16);
        }
        cp_async_commit();
        if (lane < nrows) {                       // raw dt of (token j0+lane, own head): one scalar per lane
            const int src = ord ? __ldg(ord + j0 + lane) : (j0 + lane);
            dt_next = to_f32<T>(in_base[static_cast<int64_t>(src) * G.in_ts + dt_off]);
        }
    };

    float h[kN];
#pragma unroll
    for (int n = 0; n < kN; ++n) h[n] = 0.f;
    float winx[3] = {0.f, 0.f, 0.f}, winb[3] = {0.f, 0.f, 0.f};

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
        const int nrows = min(kCH, L - j0);

        // ---- per-chunk prologue: conv + SiLU of B|C (lane = B|C channel), dt / decay (lane = token) ----
        for (int jj = 0; jj < nrows; ++jj) {
            const float v = to_f32<T>(S.bcs[buf][jj][lane]);
            float acc = bb;
            acc = fmaf(wb[0], winb[0], acc);
            acc = fmaf(wb[1], winb[1], acc);
            acc = fmaf(wb[2], winb[2], acc);
            acc = fmaf(wb[3], v, acc);
            S.bc[jj][lane] = silu_fast(acc);
            winb[0] = winb[1]; winb[1] = winb[2]; winb[2] = v;
        }
        if (lane < nrows) {
            const float dt = softplus_f(dt_raw + dtb);
            S.dt[lane] = dt;
            S.dA[lane] = ex2_approx(dt * A2);
        }
        __syncwarp();

        // ---- recurrence; lane = channel ----
#pragma unroll 2
        for (int jj = 0; jj < nrows; ++jj) {
            const float xr = to_f32<T>(S.xs[buf][jj][lane]);
            float acc = bx;
            acc = fmaf(wx[0], winx[0], acc);
            acc = fmaf(wx[1], winx[1], acc);
            acc = fmaf(wx[2], winx[2], acc);
            acc = fmaf(wx[3], xr, acc);
            const float xv = silu_fast(acc);
            winx[0] = winx[1]; winx[1] = winx[2]; winx[2] = xr;
            const float dA = S.dA[jj], dtx = S.dt[jj] * xv;
            const float4* bc = reinterpret_cast<const float4*>(&S.bc[jj][0]);
            float y = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 Bq = bc[q], Cq = bc[4 + q];
                const float Bv[4] = {Bq.x, Bq.y, Bq.z, Bq.w}, Cv[4] = {Cq.x, Cq.y, Cq.z, Cq.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = q * 4 + i;
                    h[n] = fmaf(dA, h[n], dtx * Bv[i]);
                    y = fmaf(h[n], Cv[i], y);
                }
            }
            y = fmaf(Dh, xv, y);
            float v = y;
            if (p.gate) v *= silu_fast(to_f32<T>(S.zs[buf][jj][lane]));
            const int row = (p.out_order == DM_OUT_TOKEN_ORDER) ? S.rows[buf][jj] : (j0 + jj);
            out_base[static_cast<int64_t>(row) * G.out_ts] = from_f32<T>(v);
            if (ssq) {
                float s2 = v * v;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                if (lane == 0) atomicAdd(ssq + row, s2);
            }
        }
        __syncwarp();
    }
}

template <typename T>
int launch_m2(const M2P& p, cudaStream_t stream) {
    const int n_units = p.n_groups * p.B * p.K * (p.D / 32);
    const size_t bytes = sizeof(SsdSmem<T>) * kWarps;
    static thread_local bool configured = false;
    if (!configured) {
        DM_CUDA_TRY(cudaFuncSetAttribute(m2_ssd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(bytes)));
        configured = true;
    }
    m2_ssd_kernel<T><<<(n_units + kWarps - 1) / kWarps, kWarps * 32, bytes, stream>>>(p, n_units);
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
    if (P % 32 != 0) return DM_ERR_UNSUPPORTED;
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
