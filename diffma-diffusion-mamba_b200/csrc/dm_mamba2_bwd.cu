// Mamba-2 backward (training: reference block/mamba2.py:392 reached from train.py:258-259
// with --use-mamba2; upstream MambaSplitConv1dScanCombinedFn.backward minus the RMSNorm scale and the out-projection).
//
// The SSD recurrence is the S6 recurrence with A[d, n] = A_head(d) and delta[d] = dt_head(d), so the reverse scan and the
// conv backward of the x channels run on dm_mamba1_scan_bwd (csrc/dm_mamba1_bwd.cu) fed with SSD operands.  What this
// file adds is everything that used to be ~250 torch launches per block around that kernel (gathers, a depthwise
// conv1d forward and its autograd backward, concatenations, casts):
//
//   phase 0  operand preparation in ONE pass over the gathered rows: causal conv1d + SiLU of x, B, C in scan order
//            -> u (act dtype) and the x_dbl rows in dm_mamba1's format [dt hi | dt lo | B | C], dt of head h in the
//            dt_low slot h (the "dt_proj" of the S6 view is the one-hot head map);
//   phase 2  conv backward of the B and C channels (d B, d C come out of the reverse scan as d_x_dbl[..., 32:64]).
//
// The x channels' conv backward is dm_mamba1_scan_bwd's phase 2 called with xz = zxbcdt + d_inner (x at offset 0), the
// reverse scan its phase 1 with xz = zxbcdt - d_inner (z at offset d_inner): no copy of the activations is made.
#include "dm_common.cuh"

namespace dm {
namespace {

constexpr int kN2 = 16, kW2 = 4, kR2 = 32, kE2 = 64;

struct M2B {
    const void* in; int64_t in_bs, in_ts;
    void* u; float* x_dbl; const float* d_x_dbl; float* d_bc;
    const float* conv_w; const float* conv_b; float* d_conv_w; float* d_conv_b;
};
struct M2BP {
    int B, K, L, D, H, n_groups;
    const int32_t* order;
    M2B g[DM_MAX_GROUPS];
};

__device__ __forceinline__ const int32_t* dir_order2(const M2BP& p, int k) {
    if (p.order == nullptr) return nullptr;
    const int32_t* o = p.order + static_cast<int64_t>(k) * p.L;
    return (__ldg(o) < 0) ? nullptr : o;
}

// thread = (sequence, channel) over the D + 2N conv channels plus 32 dt slots; sliding window over the scanned tokens
template <typename T>
__global__ void __launch_bounds__(128) m2_bwd_prep_kernel(const __grid_constant__ M2BP p) {
    const int D = p.D, L = p.L, Cc = D + 2 * kN2;
    const int cblocks = (Cc + kR2 + 127) / 128;
    const int seq = blockIdx.x / cblocks, c = (blockIdx.x % cblocks) * 128 + threadIdx.x;
    if (c >= Cc + kR2) return;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M2B& G = p.g[g];
    const int32_t* ord = dir_order2(p, k);
    const int64_t sg = static_cast<int64_t>(b) * p.K + k;
    const T* base = static_cast<const T*>(G.in) + static_cast<int64_t>(b) * G.in_bs;
    float* xd = G.x_dbl + sg * L * kE2;
    if (c >= Cc) {                                   // dt slot h: hi / lo halves of the raw dt of head h (0 beyond nheads)
        const int h = c - Cc;
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(xd);
        for (int j = 0; j < L; ++j) {
            float v = 0.f;
            if (h < p.H) {
                const int src = ord ? __ldg(ord + j) : j;
                v = to_f32<T>(base[static_cast<int64_t>(src) * G.in_ts + D + Cc + h]);
            }
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            row[static_cast<int64_t>(j) * 2 * kE2 + h] = hi;
            row[static_cast<int64_t>(j) * 2 * kE2 + kR2 + h] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
        return;
    }
    const float4 wv = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c) * kW2));
    const float bias = G.conv_b ? __ldg(G.conv_b + c) : 0.f;
    const T* x_base = base + D + c;                  // [z | x | B | C | dt]: conv channel c at offset d_inner + c
    T* u_out = static_cast<T*>(G.u) + sg * L * D + c;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    for (int j = 0; j < L; ++j) {
        const int src = ord ? __ldg(ord + j) : j;
        const float xn = to_f32<T>(x_base[static_cast<int64_t>(src) * G.in_ts]);
        float pre = bias;
        pre = fmaf(wv.x, x0, pre); pre = fmaf(wv.y, x1, pre); pre = fmaf(wv.z, x2, pre); pre = fmaf(wv.w, xn, pre);
        const float a = pre * sigmoid_fast(pre);
        if (c < D) u_out[static_cast<int64_t>(j) * D] = from_f32<T>(a);
        else xd[static_cast<int64_t>(j) * kE2 + kR2 + (c - D)] = a;
        x0 = x1; x1 = x2; x2 = xn;
    }
}

// conv backward of the 2N channels B | C: thread = (sequence, channel)
template <typename T>
__global__ void __launch_bounds__(2 * kN2) m2_bwd_conv_bc_kernel(const __grid_constant__ M2BP p) {
    const int D = p.D, L = p.L;
    const int seq = blockIdx.x, ch = threadIdx.x, c = D + ch;
    const int k = seq % p.K, b = (seq / p.K) % p.B, g = seq / (p.K * p.B);
    const M2B& G = p.g[g];
    const int32_t* ord = dir_order2(p, k);
    const int64_t sg = static_cast<int64_t>(b) * p.K + k;
    const T* x_base = static_cast<const T*>(G.in) + static_cast<int64_t>(b) * G.in_bs + D + c;
    const float* dact = G.d_x_dbl + sg * L * kE2 + kR2 + ch;
    float* dx_out = G.d_bc + sg * L * 2 * kN2 + ch;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(G.conv_w + static_cast<int64_t>(c) * kW2));
    const float w[kW2] = {wv.x, wv.y, wv.z, wv.w};
    const float bias = G.conv_b ? __ldg(G.conv_b + c) : 0.f;
    float xw[3] = {0.f, 0.f, 0.f}, dcw[3] = {0.f, 0.f, 0.f};
    float dw[kW2] = {0.f, 0.f, 0.f, 0.f}, db = 0.f;
    for (int j = 0; j < L + 3; ++j) {
        float xn = 0.f, dc = 0.f;
        if (j < L) {
            const int src = ord ? __ldg(ord + j) : j;
            xn = to_f32<T>(x_base[static_cast<int64_t>(src) * G.in_ts]);
            float pre = bias;
            pre = fmaf(w[0], xw[0], pre); pre = fmaf(w[1], xw[1], pre); pre = fmaf(w[2], xw[2], pre); pre = fmaf(w[3], xn, pre);
            const float s = sigmoid_fast(pre);
            dc = dact[static_cast<int64_t>(j) * kE2] * s * fmaf(pre, 1.0f - s, 1.0f);
            dw[0] = fmaf(dc, xw[0], dw[0]); dw[1] = fmaf(dc, xw[1], dw[1]); dw[2] = fmaf(dc, xw[2], dw[2]);
            dw[3] = fmaf(dc, xn, dw[3]);
            db += dc;
        }
        if (j >= 3) dx_out[static_cast<int64_t>(j - 3) * 2 * kN2] = fmaf(dcw[0], w[3], fmaf(dcw[1], w[2], fmaf(dcw[2], w[1], dc * w[0])));
        xw[0] = xw[1]; xw[1] = xw[2]; xw[2] = xn;
        dcw[0] = dcw[1]; dcw[1] = dcw[2]; dcw[2] = dc;
    }
#pragma unroll
    for (int t = 0; t < kW2; ++t) atomicAdd(G.d_conv_w + static_cast<int64_t>(c) * kW2 + t, dw[t]);
    if (G.d_conv_b) atomicAdd(G.d_conv_b + c, db);
}

}  // namespace
}  // namespace dm

extern "C" int dm_mamba2_ssd_bwd(const dm_mamba2_args* a, const dm_mamba2_bwd_group* gr, int phase, void* stream) {
    using namespace dm;
    if (a == nullptr || gr == nullptr || (phase != 0 && phase != 2)) return DM_ERR_INVALID_ARG;
    if (a->batch <= 0 || a->n_dir <= 0 || a->seqlen <= 0 || a->n_groups <= 0 || a->n_groups > DM_MAX_GROUPS)
        return DM_ERR_INVALID_ARG;
    if (a->act_dtype != DM_F32 && a->act_dtype != DM_BF16) return DM_ERR_UNSUPPORTED;
    if (a->d_state != kN2 || a->d_conv != kW2 || a->nheads <= 0 || a->nheads > kR2 || a->d_inner <= 0) return DM_ERR_UNSUPPORTED;
    M2BP p{};
    p.B = a->batch; p.K = a->n_dir; p.L = a->seqlen; p.D = a->d_inner; p.H = a->nheads; p.n_groups = a->n_groups;
    p.order = a->order;
    for (int g = 0; g < a->n_groups; ++g) {
        const dm_mamba2_group& s = a->group[g];
        const dm_mamba2_bwd_group& r = gr[g];
        if (!s.zxbcdt || !s.conv_weight || !aligned16(s.conv_weight)) return DM_ERR_INVALID_ARG;
        if (phase == 0 && (!r.u || !r.x_dbl || !aligned16(r.x_dbl))) return DM_ERR_INVALID_ARG;
        if (phase == 2 && (!r.d_x_dbl || !r.d_bc || !r.d_conv_weight)) return DM_ERR_INVALID_ARG;
        M2B& d = p.g[g];
        d.in = s.zxbcdt; d.in_bs = s.in_batch_stride; d.in_ts = s.in_token_stride;
        d.u = r.u; d.x_dbl = r.x_dbl; d.d_x_dbl = r.d_x_dbl; d.d_bc = r.d_bc;
        d.conv_w = s.conv_weight; d.conv_b = s.conv_bias; d.d_conv_w = r.d_conv_weight; d.d_conv_b = r.d_conv_bias;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_seq = p.n_groups * p.B * p.K;
    if (phase == 0) {
        const int cblocks = (p.D + 2 * kN2 + kR2 + 127) / 128;
        if (a->act_dtype == DM_F32) m2_bwd_prep_kernel<float><<<n_seq * cblocks, 128, 0, st>>>(p);
        else m2_bwd_prep_kernel<__nv_bfloat16><<<n_seq * cblocks, 128, 0, st>>>(p);
    } else {
        if (a->act_dtype == DM_F32) m2_bwd_conv_bc_kernel<float><<<n_seq, 2 * kN2, 0, st>>>(p);
        else m2_bwd_conv_bc_kernel<__nv_bfloat16><<<n_seq, 2 * kN2, 0, st>>>(p);
    }
    DM_CUDA_TRY(cudaGetLastError());
    return DM_OK;
}
