"""Which A-operand addressings does tcgen05.mma accept for a SWIZZLE_128B K-major tile?  (start row offset, group stride, base_offset)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pvo_b200 import _lib
from pvo_b200._lib import check, ptr, stream
g = torch.Generator(device='cuda').manual_seed(0)
rows = 512
G = torch.randn(rows, 64, device='cuda', generator=g).half()
B = torch.randn(64, 64, device='cuda', generator=g).half()
D = torch.empty(128, 64, device='cuda')
for sbo in (1024, 2048, 1280, 1152):
    for off in (0, 1, 2, 3, 5, 8, 9, 17):
        for bo in sorted({0, off & 7}):
            pitch = sbo // 128
            idx = torch.tensor([off + (i // 8) * pitch + (i % 8) for i in range(128)], device='cuda')
            if int(idx.max()) >= rows:
                continue
            want = G[idx].float() @ B.float().t()
            D.zero_()
            check(_lib.lib().pvo_umma_probe(ptr(G), ptr(B), rows, off, sbo, bo, ptr(D), stream()), 'umma_probe')
            torch.cuda.synchronize()
            err = (D - want).abs().max().item() / want.abs().max().item()
            print('sbo=%4d row_off=%2d base_offset=%d  rel err %.3e  %s' % (sbo, off, bo, err, 'OK' if err < 2e-3 else 'WRONG'), flush=True)
