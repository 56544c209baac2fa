// Mamba-2 "SSD" chunked form on tensor cores (bf16 I/O), the matmul formulation of the same recurrence that
// m2_ssd_kernel evaluates sequentially (upstream ssd_combined / ssd_minimal_discrete):
//
//   per (sequence, head), 64-token chunks, a_t = dt_t A_h, cum_t = inclusive cumsum of a within the chunk:
//     S      = C B^T                                   (64 x 64, K = 16)        tensor core
//     M[i,j] = S[i,j] exp(cum_i - cum_j) dt_j  (j <= i)                          registers (accumulator -> A fragment)
//     Y      = M X  +  diag(exp(cum)) C State_prev     (64 x 64, K = 64 / 16)   tensor core
//     State  = exp(cum_last) State_prev + (B o w)^T X,  w_j = exp(cum_last - cum_j) dt_j   (16 x 64, K = 64)
//     v      = (Y + D_h X) silu(z)   ->  out, sum_c v^2 per token
//   X, B, C = silu(causal_conv1d(.)) are produced in shared memory from the gathered rows (scan order folded into
//   the cp.async row gather).  One CTA (4 warps) per (sequence, head); warp w owns chunk rows [16w, 16w+16) for the
//   S / M / Y products and channels [16w, 16w+16) of the state.  Exps per (token, head): ~Q/2 instead of the sequential
//   form's P*... per channel; MUFU and FP32 load drop by ~4x, the contractions run on the (legacy mma.sync) tensor path.
#pragma once

namespace dm {
namespace ssd {

constexpr int Q = 64;          // chunk length (tokens)
constexpr int P = 64;          // head dim (channels per head)
constexpr int NS = 16;         // d_state
constexpr int LDX = P + 8;     // padded bf16 row strides (odd multiples of 16 B: ldmatrix conflict-free)
constexpr int LDB = NS + 8;
constexpr int kThreads = 128;

struct Smem {
    __nv_bfloat16 xraw[Q + 3][P];        // gathered x rows incl. 3-token halo
    __nv_bfloat16 bcraw[Q + 3][2 * NS];  // gathered [B | C] rows incl. halo
    __nv_bfloat16 zs[Q][P];              // gathered z rows
    __nv_bfloat16 xs[Q][LDX];            // X = silu(conv(x))
    __nv_bfloat16 bs[Q][LDB];            // B
    __nv_bfloat16 cs[Q][LDB];            // C
    __nv_bfloat16 st[NS][LDX];           // running state, bf16 copy for the inter-chunk product
    float cum[Q];                        // inclusive cumsum of dt*A*log2(e) inside the chunk
    float dtv[Q];                        // dt (0 beyond the end of the sequence)
    float part[2];                       // cumsum hand-off between the two scanning warps
    int rows[Q];                         // output row of each token
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float tanh_ap(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_t(float x) {       // x*sigmoid(x) with one MUFU; result is rounded to bf16 by callers
    const float h = 0.5f * x;
    return fmaf(h, tanh_ap(h), h);
}

}  // namespace ssd
}  // namespace dm
