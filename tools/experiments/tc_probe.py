"""Timing probe for the tensor-core retrieval path (CUDA events, each repetition printed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gpt2_image_captioning_b200.database import _GpuFlatIndex
dev = 'cuda:0'
g = torch.Generator().manual_seed(2)
D, B = 512, 1024
for N in (118287, 131072, 591753):
    db = torch.randn(N, D, generator=g); db /= db.norm(dim=-1, keepdim=True)
    q = torch.randn(B, D, generator=g); q /= q.norm(dim=-1, keepdim=True)
    dbd = db.to(dev); qd = q.to(dev)
    for tc in (False, True):
        ix = _GpuFlatIndex(dbd, tensor_cores=tc)
        for k in (5, 15):
            ts = []
            for _ in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); s, i = ix.search_device(qd, k); b.record(); torch.cuda.synchronize()
                ms = round(a.elapsed_time(b), 2)
                if tc:
                    al = lambda v: (v + 255) // 256 * 256
                    chunk = min(N, 32768)
                    base = (ix._ws.data_ptr() + 255) // 256 * 256 - ix._ws.data_ptr()
                    off = base + al(B * chunk * 4) + 2 * al(B * D * 2) + al(B * 32 * 4) + al(B * 32 * 8)
                    ms = (ms, int(ix._ws[off + B * 4: off + B * 4 + 4].view(torch.int32).item()))
                ts.append(ms)
            print('N', N, 'tc' if tc else 'exact', 'k', k, 'ms', ts, flush=True)
    del dbd
