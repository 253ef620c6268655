import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pvo_b200.modules.corr import CorrBlock
g = torch.Generator(device='cuda').manual_seed(0)
E, h, w = 36, 60, 80
f1 = torch.randn(1, E, 128, h, w, device='cuda', generator=g).half()
f2 = torch.randn(1, E, 128, h, w, device='cuda', generator=g).half()
blk = CorrBlock(f1, f2)
yy, xx = torch.meshgrid(torch.arange(h, device='cuda').float(), torch.arange(w, device='cuda').float(), indexing='ij')
grid = torch.stack([xx, yy], -1)[None, None].repeat(1, E, 1, 1, 1).contiguous()
blk2 = CorrBlock(f1, f2)
blk2.corr_pyramid = [p * 2 for p in blk.corr_pyramid]
c = grid + 0.37
a, b = blk2(c).float(), blk(c).float() * 2
bad = ((a - b).abs() > 4e-3 * b.abs() + 1e-6)
print('shape', a.shape, 'bad', bad.sum().item(), 'of', bad.numel(), 'neq', (a != b).sum().item())
idx = bad.nonzero()[:10]
for i in idx.tolist():
    print(i, a[tuple(i)].item(), b[tuple(i)].item())
print('max vol', [p.abs().max().item() for p in blk.corr_pyramid])
