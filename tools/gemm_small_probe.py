"""Back-to-back launch time of the two bf16 contraction kernels (TMA / tcgen05 vs small-shape mma.sync) on the per-step shapes of
the XS dream rollout: a dependent chain of 200 launches on one stream (each accumulates into the same C), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_v3_b200 import ops

shapes = [(1024, 256, 256), (1024, 768, 256), (1024, 1024, 256), (1024, 256, 768), (1024, 256, 1024), (1024, 256, 1280),
          (512, 512, 512), (2048, 256, 256)]
for M, N, K in shapes:
    A16 = ops.to_bf16(torch.randn(M, K).cuda())
    W = torch.randn(K, N).cuda()
    out = []
    for small in (False, True):
        for beta in (0.0, 1.0):
            run = ops.gemm_bf16(A16, W, beta=beta, small=small)
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            st = torch.cuda.Stream()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                with torch.cuda.graph(gr, stream=st):   # a dependent chain of 100 launches as graph kernel nodes
                    for _ in range(100):
                        run()
                gr.replay()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                for _ in range(5):
                    gr.replay()
                b.record(st)
                torch.cuda.synchronize()
            out.append(a.elapsed_time(b) / 500 * 1e3)
    print(f"M{M} N{N} K{K}: tma beta0 {out[0]:.2f} beta1 {out[1]:.2f} us | small beta0 {out[2]:.2f} beta1 {out[3]:.2f} us", flush=True)
