// Kernel P3 (bf16, d_inner 1024): gather + causal conv1d + SiLU -> u, and x_dbl = u . W_x^T on the 5th-generation
// tensor cores (tcgen05.mma, accumulator in TMEM).  Included by dm_mamba1.cu (uses its M1P / M1G and helpers).
//
// Why: the mma.sync version (m1_conv_xproj_persistent) re-reads W_x (128 KB) from shared memory through ldmatrix for
// every 16-token tile and serialises load / conv / MMA / reduce with five block barriers per tile: 43 us at the headline
// shape = 0.29 of HBM, `mio_throttle` on LDSM its top stall (profiles/r01_ncu_m1_conv_xproj.txt).  Here
//   * the tile is 128 scanned tokens (M = 128): W_x is read by the tensor core straight from shared memory through a
//     descriptor -- once per 128 tokens instead of once per 16, and not through the LSU at all;
//   * the contraction is asynchronous: the conv + SiLU of channel slice s+1 (all 16 warps) overlaps the MMA of slice s;
//   * ONE block barrier per 64-channel slice (the cp.async staging of slice s+1 is waited for before that same barrier).
//
// Work decomposition.  Rows = scanned tokens of one mixer ("group") flattened over (batch, direction, position); a tile
// is 128 consecutive rows and never crosses a group (different weights), but may cross sequences: every row carries its
// position j in its own sequence, and the conv window is reset where j = 0.  At the headline shape (2 mixers x 9 408
// rows) that is 148 tiles = one per SM.
//
// Per tile, for each of the 16 slices of 64 channels:
//   cp.async   x rows (gather by scan order) of slice s+2 -> XS[(s+2) % 3]     (131 rows x 128 B: 3 halo rows)
//   conv       thread = (channel pair p, 8-token segment): sliding window over XS[s % 3] -> u (global, bf16) and the A
//              operand tile AS[s % 2] (128 rows x 64 bf16) in the 128-byte-swizzled K-major layout the MMA descriptor names
//   barrier    (+ fence.proxy.async: generic-proxy writes -> async-proxy reads)
//   MMA        one thread: 4 x tcgen05.mma M128 N64 K16, D (TMEM, 128 lanes x 64 fp32 columns) += AS[s % 2] . WX[s]^T;
//              tcgen05.commit -> a_free[s % 2] (the conv of slice s+2 may overwrite the buffer)
// (A variant without any block barrier inside a tile -- staging completion, buffer release and the A hand-over all on
// mbarriers, a 17th warp issuing the MMAs -- was built and measured: 37.6 us against 30.9 us for this one; 512 threads
// polling mbarriers cost more issue slots than the hardware barrier they replaced.  profiles/r02_notes.md.)
// then tcgen05.commit -> acc_full; warps 0..3 read the accumulator (tcgen05.ld, lane = row) and write the x_dbl rows
// [dt_low hi | dt_low lo | B | C].  W_x lives in shared memory as 16 slices of [64 rows (N) x 64 bf16] (8 KB each, same
// swizzle), loaded once per CTA and group with 16-byte cp.async, slice s riding along with the x rows of slice s of the
// first tile (the MMA of slice s needs only WX[s]).
// Small batches (fewer tiles than half the SMs; config C5's batch 8 has 74 tiles for 148 SMs): the <kSplit = true> variant
// gives a tile to the two CTAs of a cluster -- each convolves 8 of the 16 channel slices and contracts them against its half
// of W_x (split K), CTA 1 ships its partial accumulator to CTA 0 through distributed shared memory (mapa + st.shared::cluster,
// a start-of-kernel cluster arrival + one exchange barrier), CTA 0 adds and writes the rows.  42.0 -> 31.7 us for
// conv + x_proj + delta at batch 8.
// Conv inner loop (r02 ncu, first version: 407 instructions per warp and slice, a third of the stall samples on the
// per-token "does a new sequence start here" branch and the dependent shared-memory load behind it; second version: a
// separate slow path for the ~4 % of segments that contain a sequence start kept the other 15 warps of the CTA waiting at
// the slice barrier): EVERY segment takes the branch-free path (11 window values loaded up front, the two channels of a
// thread on packed fp32x2 FMAs), and the warp that owns the <= 3 tokens with fewer than 3 predecessors recomputes just
// those afterwards; the staging copies' addresses are computed once per tile.
#pragma once

namespace p3 {

constexpr int kTile = 128;                   // rows (scanned tokens) per tile = UMMA M
constexpr int kSl = 64;                      // channels per slice = UMMA K per stage
constexpr int kNS = 1024 / kSl;              // 16 slices
constexpr int kThreads = 512;
constexpr int kXRows = kTile + 3;            // with the conv halo
constexpr int kXSBytes = kXRows * 128;       // 16 768
constexpr int kASBytes = kTile * 128;        // 16 384
constexpr int kWXBytes = kE * 1024 * 2;      // 131 072
// shared-memory carve (1024-byte aligned base): WX | AS[2] | XS[3] | srcoff[131] | jpos[131] | barriers | tmem slot
constexpr int kOffAS = kWXBytes;
constexpr int kOffXS = kOffAS + 2 * kASBytes;
constexpr int kOffTab = kOffXS + 3 * kXSBytes;
constexpr int kOffBar = kOffTab + ((kXRows * 8 + 15) & ~15);
constexpr int kSmemBytes = kOffBar + 128 + 1024;         // 12 mbarriers + TMEM slot, + alignment slack

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {       // K-major, 128 B swizzle, 8-row groups 1024 B apart
    return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A / B bf16, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kE >> 3) << 17) | (static_cast<uint32_t>(kTile >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bounded(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;               // common case: already complete
#pragma unroll 1                                          // (unrolled 64x by default: with ~10 call sites the kernel outgrew the instruction cache)
    for (uint32_t i = 0; i < (1u << 26); ++i)
        if (mbar_try_wait(bar, parity)) return;
    __trap();                                             // a wrong descriptor must not hang the GPU
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace p3

// kSplit (small batches: fewer tiles than half the SMs): a tile is shared by the two CTAs of a cluster, CTA `half` convolves
// channel slices half * 8 .. half * 8 + 7 and contracts them against its half of W_x (split K); CTA 1 then hands its
// partial accumulator (128 x 64 fp32) to CTA 0 through distributed shared memory, CTA 0 adds it to its own and writes the
// x_dbl rows.  One tile per cluster (grid = 2 x tiles), so nothing persists across tiles.
template <bool kSplit>
__global__ void __launch_bounds__(p3::kThreads, 1)
m1_conv_xproj_tc(const __grid_constant__ M1P p, int rows_per_group, int tiles_per_group) {
    using namespace p3;
    constexpr int kNSl = kSplit ? kNS / 2 : kNS;                             // slices this CTA works through
    const int half = kSplit ? static_cast<int>(blockIdx.x & 1) : 0;
    const int s0 = half * kNSl;                                              // first (global) slice of this CTA
    using T = __nv_bfloat16;
    constexpr int kD = 1024;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    int32_t* srcoff = reinterpret_cast<int32_t*>(sm + kOffTab);              // element offset of each staged row's x row, -1 = none
    int32_t* jpos = srcoff + kXRows;                                         // position of the row in its own sequence
    const uint32_t bar0 = base + kOffBar;                                    // a_free[0], a_free[1], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 96);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, K = p.K;

    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_init(bar0 + 16, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;
    pdl_launch_dependents();
    if constexpr (kSplit) {
        // distributed shared memory may only be touched once its CTA is known to be running: both CTAs arrive here, the
        // matching wait sits in front of the exchange (by then long complete)
        asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    }

    const int n_tiles = p.n_groups * tiles_per_group;
    int cur_group = -1;
    uint32_t tiles_done = 0u;               // parity bookkeeping: A buffer ab has been handed to the MMA 8 * tiles_done + (s >> 1) times
    const int pr = tid & 31, seg = tid >> 5;          // conv role: channel pair of the slice, 8-token segment

    for (int tile = kSplit ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x); tile < n_tiles;
         tile += kSplit ? n_tiles : static_cast<int>(gridDim.x), ++tiles_done) {
        const int g = tile / tiles_per_group, t_in = tile - g * tiles_per_group;
        const M1G& G = p.g[g];
        const int row0 = t_in * kTile;                                       // first row of the tile inside the group
        __syncthreads();                                                     // previous tile fully retired (tables, XS, AS)
        // ---- per-tile row tables ----
        if (tid < kXRows) {
            const int row = row0 - 3 + tid;
            int off = -1, j = 0;
            if (row >= 0 && row < rows_per_group) {
                const int seq = row / L;
                j = row - seq * L;
                const int b = seq / K, k = seq - b * K;
                int src = j;
                if (p.order != nullptr) {
                    const int32_t* o = p.order + static_cast<int64_t>(k) * L;
                    if (__ldg(o) >= 0) src = __ldg(o + j);
                }
                off = static_cast<int32_t>(static_cast<int64_t>(b) * G.xz_bs + static_cast<int64_t>(src) * G.xz_ts);
            }
            srcoff[tid] = off;
            jpos[tid] = j;
        }
        const float* conv_w = G.conv_w;                                      // (kept in registers: re-reading the kernel parameter
        const float* conv_b = G.conv_b;                                      //  block with a dynamic group index costs an LDC per slice)
        const bool load_w = g != cur_group;                                  // W_x slices ride along with the x slices of this tile
        cur_group = g;
        __syncthreads();                                                     // tables visible
        // ---- this thread's staging copies (the same (row, 16-byte chunk) for every slice): up to 3 of the 131 x 8 ----
        const T* xz = static_cast<const T*>(G.xz);
        const T* cp_src[3];
        uint32_t cp_dst[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int i = tid + q * kThreads;
            const int r = i >> 3, c = i & 7;
            const int off = (i < kXRows * 8) ? srcoff[r] : -1;
            cp_src[q] = off >= 0 ? xz + off + c * 8 : nullptr;
            cp_dst[q] = r * 128 + c * 16;
        }
        // W_x chunk of this thread: row n = tid / 8, chunk cc = tid % 8 of every 64-channel slice (128-byte swizzle)
        const T* w_src = static_cast<const T*>(G.wx) + static_cast<int64_t>(tid >> 3) * kD + (tid & 7) * 8;
        const uint32_t w_dst = base + (tid >> 3) * 128 + ((((tid & 7) ^ ((tid >> 3) & 7))) << 4);
        auto stage = [&](int ls) {                                           // x rows (+ W_x) of local slice ls -> XS[ls % 3] (WX[s])
            const int s = s0 + ls;
            const uint32_t dst0 = base + kOffXS + (ls % 3) * kXSBytes;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (cp_src[q] != nullptr) cp_async16(dst0 + cp_dst[q], cp_src[q] + s * kSl);
            if (load_w) cp_async16(w_dst + s * (kE * 128), w_src + s * kSl);
            cp_async_commit();
        };
        if (tiles_done == 0u) pdl_wait();                                    // tables / barriers / TMEM are set up; xz is the predecessor's output
        stage(0);
        stage(1);
        // this thread's 8 rows: validity, A-operand offsets
        const int r_first = seg * 8;                                         // tile row of the first token of the segment
        T* u_out = static_cast<T*>(G.u) + static_cast<int64_t>(row0 + r_first) * kD + 2 * pr;
        const int nvalid = min(8, rows_per_group - (row0 + r_first));        // rows of the segment inside the group (<= 0: none)
        // Sequence starts inside the window of these 8 rows (rare: one row in L).  ts = local index of the first token of
        // a sequence (negative: it started 1 or 2 rows before the segment; 99: none).  Tokens ts .. ts+2 have fewer than 3
        // predecessors: every warp runs the branch-free path, the warp that owns such tokens then recomputes just those.
        int ts = 99;
        {
            const int j0 = jpos[r_first + 3];                                // position of the segment's first token
#pragma unroll
            for (int t = 7; t >= 1; --t)
                if (jpos[r_first + 3 + t] == 0) ts = t;                      // (L >= 11: at most one start per window)
            if (j0 < 3) ts = -j0;
        }
        uint32_t a_off[8];                                                   // row r, channels 2 pr, 2 pr + 1: chunk (pr / 4) ^ (r % 8)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int r = r_first + t;
            a_off[t] = r * 128 + ((((pr >> 2) ^ (r & 7)) << 4) | ((pr & 3) << 2));
        }
        // conv weights of this thread's two channels of a slice; fetched one slice ahead (an L2 round trip per slice on the
        // critical path otherwise)
        float4 w0n = __ldg(reinterpret_cast<const float4*>(conv_w + static_cast<int64_t>(s0 * kSl + 2 * pr) * kW));
        float4 w1n = __ldg(reinterpret_cast<const float4*>(conv_w + static_cast<int64_t>(s0 * kSl + 2 * pr + 1) * kW));
        float2 bbn = conv_b ? __ldg(reinterpret_cast<const float2*>(conv_b + s0 * kSl + 2 * pr)) : make_float2(0.f, 0.f);

        for (int ls = 0; ls < kNSl; ++ls) {
            const int s = s0 + ls;                                           // global slice: channels s * 64 .. s * 64 + 63
            if (ls + 2 < kNSl) stage(ls + 2); else cp_async_commit();        // (empty group keeps the wait counts uniform)
            const float4 w0 = w0n, w1 = w1n;
            const float2 bb = bbn;
            if (ls + 1 < kNSl) {
                const int c = (s + 1) * kSl + 2 * pr;
                w0n = __ldg(reinterpret_cast<const float4*>(conv_w + static_cast<int64_t>(c) * kW));
                w1n = __ldg(reinterpret_cast<const float4*>(conv_w + static_cast<int64_t>(c + 1) * kW));
                if (conv_b) bbn = __ldg(reinterpret_cast<const float2*>(conv_b + c));
            }
            cp_async_wait<2>();                                              // slice s has landed (this thread's copies)
            if (ls == 0) __syncthreads();                                    // ... and everybody else's (later slices: the loop barrier)
            const int ab = ls & 1;
            const uint8_t* xs = sm + kOffXS + (ls % 3) * kXSBytes + pr * 4 + r_first * 128;
            uint8_t* as = sm + kOffAS + ab * kASBytes;
            T* u_s = u_out + s * kSl;
            // branch-free path: 11 independent window loads, the thread's two channels on packed fp32x2 math
            uint32_t xv[11];
#pragma unroll
            for (int i = 0; i < 11; ++i) xv[i] = *reinterpret_cast<const uint32_t*>(xs + i * 128);
            uint64_t x2[11];
#pragma unroll
            for (int i = 0; i < 11; ++i) x2[i] = pack2(__uint_as_float(xv[i] << 16), __uint_as_float(xv[i] & 0xffff0000u));
            const uint64_t k0 = pack2(w0.x, w1.x), k1 = pack2(w0.y, w1.y), k2 = pack2(w0.z, w1.z), k3 = pack2(w0.w, w1.w);
            const uint64_t kb = pack2(bb.x, bb.y), half2 = pack2(0.5f, 0.5f);
            // the MMA that read AS[s % 2] two slices ago must have completed
            const uint32_t used = tiles_done * (kNSl / 2) + (ls >> 1);
            if (used > 0) wait_bounded(bar0 + 8 * ab, (used - 1) & 1);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                uint64_t y = fma2(k0, x2[t], kb);
                y = fma2(k1, x2[t + 1], y);
                y = fma2(k2, x2[t + 2], y);
                y = fma2(k3, x2[t + 3], y);
                const uint64_t h = mul2(y, half2);                           // silu(y) = h + h tanh(h), h = y / 2
                float h0, h1;
                unpack2(h, h0, h1);
                float o0, o1;
                unpack2(fma2(h, pack2(tanh_approx(h0), tanh_approx(h1)), h), o0, o1);
                const uint32_t packed = pack_bf16(o0, o1);
                *reinterpret_cast<uint32_t*>(as + a_off[t]) = packed;
                if (t < nvalid) *reinterpret_cast<uint32_t*>(u_s + static_cast<int64_t>(t) * kD) = packed;
            }
            if (ts != 99) {                                                  // warp-uniform, ~1 warp in 25
                // tokens ts .. ts+2 (those inside the segment): n = t - ts predecessors belong to the same sequence
#pragma unroll 1
                for (int t = max(ts, 0); t <= min(ts + 2, 7); ++t) {
                    const int n = t - ts;
                    auto ldx = [&](int i) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(xs + i * 128);
                        return pack2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
                    };
                    uint64_t y = fma2(k3, ldx(t + 3), kb);
                    if (n >= 1) y = fma2(k2, ldx(t + 2), y);
                    if (n >= 2) y = fma2(k1, ldx(t + 1), y);
                    const uint64_t h = mul2(y, half2);
                    float h0, h1;
                    unpack2(h, h0, h1);
                    float o0, o1;
                    unpack2(fma2(h, pack2(tanh_approx(h0), tanh_approx(h1)), h), o0, o1);
                    const uint32_t packed = pack_bf16(o0, o1);
                    const int r = r_first + t;
                    *reinterpret_cast<uint32_t*>(as + r * 128 + ((((pr >> 2) ^ (r & 7)) << 4) | ((pr & 3) << 2))) = packed;
                    if (t < nvalid) *reinterpret_cast<uint32_t*>(u_s + static_cast<int64_t>(t) * kD) = packed;
                }
            }
            cp_async_wait<1>();                                              // slice s+1 staged (this thread's part) before the barrier
            fence_proxy_async();                                             // A tile: generic-proxy writes -> tensor-core reads
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t ad = desc_sw128(base + kOffAS + ab * kASBytes);
                const uint64_t bd = desc_sw128(base + s * (kE * 128));
#pragma unroll
                for (int k = 0; k < kSl / 16; ++k) umma(tmem_acc, ad + 2 * k, bd + 2 * k, (ls | k) != 0);
                commit(bar0 + 8 * ab);
                if (ls == kNSl - 1) commit(bar0 + 16);
            }
        }
        // ---- split K: CTA 1's partial accumulator -> CTA 0's shared memory (the upper half of its W_x region, which only
        //      holds slices 8..15 and is never touched by CTA 0), 272-byte rows ----
        constexpr int kPartStride = 68;                                      // floats per partial row (16-byte accesses conflict free)
        float* part = reinterpret_cast<float*>(sm + kWXBytes / 2);
        if constexpr (kSplit) {
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");     // (start-of-kernel arrival of the peer CTA)
            if (half == 1 && warp < 4) {
                wait_bounded(bar0 + 16, tiles_done & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int r = warp * 32 + lane;
                uint32_t remote;                                             // the same offset in CTA 0's shared window
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(part + r * kPartStride)), "r"(0));
                uint32_t v[32];
#pragma unroll
                for (int hsel = 0; hsel < 2; ++hsel) {
                    tmem_ld32(tmem_acc + (static_cast<uint32_t>(warp * 32) << 16) + 32 * hsel, v);
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (32 * hsel + i) * 4),
                                     "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3]) : "memory");
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            // every thread of both CTAs: CTA 1's stores are complete and visible before CTA 0 reads them; CTA 0's shared
            // memory stays allocated until CTA 1 has arrived
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        // ---- epilogue (warps 0..3 [of CTA 0]): accumulator (128 rows x 64) -> x_dbl rows ----
        if (warp < 4 && half == 0) {
            wait_bounded(bar0 + 16, tiles_done & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int r = warp * 32 + lane;
            const bool ok = row0 + r < rows_per_group;
            float* rowp = G.x_dbl + static_cast<int64_t>(row0 + (ok ? r : 0)) * kE;
            uint32_t v[32];
            tmem_ld32(tmem_acc + (static_cast<uint32_t>(warp * 32) << 16), v);             // columns 0..31: dt_low
            if constexpr (kSplit) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(part + r * kPartStride + i);
                    v[i] = __float_as_uint(__uint_as_float(v[i]) + q.x);
                    v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + q.y);
                    v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + q.z);
                    v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + q.w);
                }
            }
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        split_bf16(__uint_as_float(v[i + 2 * q]), __uint_as_float(v[i + 2 * q + 1]), hi[q], lo[q]);
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(rowp) + i / 2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(rowp) + 16 + i / 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            tmem_ld32(tmem_acc + (static_cast<uint32_t>(warp * 32) << 16) + 32, v);        // columns 32..63: B, C
            if constexpr (kSplit) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(part + r * kPartStride + 32 + i);
                    v[i] = __float_as_uint(__uint_as_float(v[i]) + q.x);
                    v[i + 1] = __float_as_uint(__uint_as_float(v[i + 1]) + q.y);
                    v[i + 2] = __float_as_uint(__uint_as_float(v[i + 2]) + q.z);
                    v[i + 3] = __float_as_uint(__uint_as_float(v[i + 3]) + q.w);
                }
            }
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<uint4*>(rowp + kR + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(64) : "memory");
}
