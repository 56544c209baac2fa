import sys, torch
sys.path.insert(0, '.')
from diffma_b200 import ops
torch.manual_seed(0)
dev = torch.device('cuda')
for (G, M, N, K) in [(1, 128, 128, 64), (1, 128, 128, 256), (2, 256, 256, 512), (2, 3136, 2048, 512), (2, 3136, 512, 3072), (1, 200, 136, 72)]:
    a = torch.randn(G, M, K, device=dev).bfloat16()
    b = torch.randn(G, N, K, device=dev).bfloat16()
    rs = torch.rand(G, M, device=dev) + 0.5
    c = ops.gemm_bf16_tn(a, b)
    torch.cuda.synchronize()
    ref = torch.bmm(a.float(), b.float().transpose(1, 2))
    err = (c.float() - ref).abs().max().item() / ref.abs().max().item()
    c2 = ops.gemm_bf16_tn(a, b, rs)
    err2 = (c2.float() - ref * rs[..., None]).abs().max().item() / ref.abs().max().item()
    print((G, M, N, K), 'relerr', err, err2, flush=True)
# timing vs cuBLAS
a = torch.randn(2, 3136, 512, device=dev).bfloat16(); b = torch.randn(2, 2048, 512, device=dev).bfloat16()
a2 = torch.randn(2, 3136, 3072, device=dev).bfloat16(); b2 = torch.randn(2, 512, 3072, device=dev).bfloat16()
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3
bt = b.transpose(1, 2).contiguous(); b2t = b2.transpose(1, 2).contiguous()
import os
print('config', os.environ.get('DM_GEMM_CONFIG'))
print('in_proj  ours %.1f us  cublas %.1f us' % (t(lambda: ops.gemm_bf16_tn(a, b)), t(lambda: torch.bmm(a, bt))))
# what a merged out-projection (sum of the 3 directions in the A producer => K = 1024) costs as a plain GEMM, and the
# attention_network Linear (1024 -> 512)
a3 = torch.randn(2, 3136, 1024, device=dev).bfloat16(); b3 = torch.randn(2, 512, 1024, device=dev).bfloat16()
b3t = b3.transpose(1, 2).contiguous()
a2m = torch.randn(2, 3136, 3, 1024, device=dev).bfloat16()
print("out_proj ours K=3072 %.1f us  merged n_sum=3 %.1f us  cublas K=3072 %.1f us" % (t(lambda: ops.gemm_bf16_tn(a2, b2)), t(lambda: ops.gemm_bf16_tn(a2m, b3)), t(lambda: torch.bmm(a2, b2t))))
print('K=1024   ours %.1f us  cublas %.1f us' % (t(lambda: ops.gemm_bf16_tn(a3, b3)), t(lambda: torch.bmm(a3, b3t))))
a4 = torch.randn(1, 3136, 1024, device=dev).bfloat16(); b4 = torch.randn(1, 512, 1024, device=dev).bfloat16()
print('att lin  ours %.1f us  cublas %.1f us' % (t(lambda: ops.gemm_bf16_tn(a4, b4)), t(lambda: torch.nn.functional.linear(a4[0], b4[0]))))
