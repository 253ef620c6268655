"""ConvGRU with the reference's parameters (VO_Module/droid_slam/modules/gru.py:6-32) on the fused CUDA path."""
import torch
import torch.nn as nn

from .. import nn as pnn


class ConvGRU(nn.Module):
    """z, r, q gates with a global-context term; net' = (1-z) net + z q.

    Parameter names/shapes are the reference's (convz/convr/convq: (128,448,3,3); w, conv*_glo: (128,128,1,1)), so
    checkpoints load unchanged.  forward() is inference-only (no autograd through the CUDA kernels)."""

    def __init__(self, h_planes=128, i_planes=128):
        super().__init__()
        self.do_checkpoint = False
        self.h_planes, self.i_planes = h_planes, i_planes
        self.convz = nn.Conv2d(h_planes + i_planes, h_planes, 3, padding=1)
        self.convr = nn.Conv2d(h_planes + i_planes, h_planes, 3, padding=1)
        self.convq = nn.Conv2d(h_planes + i_planes, h_planes, 3, padding=1)
        self.w = nn.Conv2d(h_planes, h_planes, 1, padding=0)
        self.convz_glo = nn.Conv2d(h_planes, h_planes, 1, padding=0)
        self.convr_glo = nn.Conv2d(h_planes, h_planes, 1, padding=0)
        self.convq_glo = nn.Conv2d(h_planes, h_planes, 1, padding=0)

    def forward(self, net, *inputs):
        if self.h_planes != 128 or self.i_planes != 320:
            raise RuntimeError('ConvGRU: the CUDA path is specialised for h_planes=128, i_planes=320 (droid_net.py:224)')
        if net.requires_grad or any(i.requires_grad for i in inputs):
            raise NotImplementedError('ConvGRU: inference only')
        sd = {'gru.' + k: v for k, v in self.state_dict().items()}
        return pnn.gru_forward(sd, net, torch.cat(inputs, dim=1))
