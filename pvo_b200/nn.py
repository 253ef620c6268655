"""Weight packing and the fused update-operator engine.

Host-side mirror of DynamicUpdateModule / ConvGRU / GraphAgg (VO_Module/droid_slam/droid_net.py:64-95,167-314,
modules/gru.py:6-32): same parameters (checkpoint names), same arithmetic, but every
activation lives in a persistent pixel-major (NHWC) device buffer and every nn.Conv2d + torch.cat + elementwise
chain is one CUDA launch through the C ABI (include/pvo_b200.h).  Python only sequences launches; the whole
sequence is CUDA-graph capturable (no allocation, no host sync inside run()).

  mode 'fp32' : float32 activations, CUDA-core implicit GEMM -- the PARITY path (gated against the CPU oracle)
  mode 'fp16' : float16 activations (what the reference's autocast inference does, factor_graph.py:227,242);
                the dense layers run on tcgen05 tensor cores (csrc/conv_tc.cu), small/odd layers on CUDA cores.
"""
import torch

from . import _lib
from ._lib import c_void_p, check, ptr, stream

CORR_CH = 196          # 4 levels x 7 x 7
CORR_STRIDE = 208      # padded so a pixel record is a multiple of 16 bytes (TMA) and of 8 channels (vector loads)
HEAD_NAMES = ('delta', 'delta_dy', 'weight', 'delta_mask')


def pack_conv_weight(w, cin_pad=None):
    """(Co,Ci,KH,KW) -> [(kh*KW+kw)*Ci' + ci][Co'] float32, Ci' = cin_pad or Ci, Co' = Co rounded up to 4."""
    co, ci, kh, kw = w.shape
    cip = cin_pad or ci
    cop = (co + 3) // 4 * 4
    out = torch.zeros(kh * kw, cip, cop, dtype=torch.float32, device=w.device)
    out[:, :ci, :co] = w.float().permute(2, 3, 1, 0).reshape(kh * kw, ci, co)
    return out.reshape(kh * kw * cip, cop).contiguous()


def pack_conv_weight_tc(w, segs=None):
    """(Co,Ci,KH,KW) -> [Co][tap][segment][64-padded channel] float16 (K-major rows: the UMMA B operand).
    segs: real channel count of each input segment (sum = Ci); each is zero-padded to a multiple of 64, matching
    the 64-channel K blocks pvo_conv2d_tc walks per segment."""
    co, ci, kh, kw = w.shape
    segs = segs or [ci]
    assert sum(segs) == ci
    wt = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).half()
    parts, c0 = [], 0
    for c in segs:
        cp = (c + 63) // 64 * 64
        blk = torch.zeros(co, kh * kw, cp, dtype=torch.float16, device=w.device)
        blk[:, :, :c] = wt[:, :, c0:c0 + c]
        parts.append(blk)
        c0 += c
    return torch.cat(parts, dim=2).reshape(co, -1).contiguous()


def pack_small_weight(ws):
    """list of G tensors (CG,Cg,3,3) -> [G][9][Cg][CG] float32 for pvo_conv3x3_small."""
    return torch.stack([w.float().permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]) for w in ws]).contiguous()


def pack_heads_mma(ws):
    """list of G tensors (2,Cg,3,3) -> mma.sync m16n8k16 B fragments [G][9][Cg/16][32 lanes][4 halves]:
    lane l holds B[k][n] for n = l//4 and k = (l%4)*2 + {0, 1, 8, 9}; only the group's 2 real outputs (n < 2) are non-zero."""
    G = len(ws)
    cg = ws[0].shape[1]
    w = torch.stack([x.float().permute(2, 3, 1, 0).reshape(9, cg, 2) for x in ws])           # [G][9][Cg][2]
    wp = torch.zeros(G, 9, cg // 16, 16, 8, device=w.device)
    wp[..., :2] = w.reshape(G, 9, cg // 16, 16, 2)
    lanes = torch.arange(32, device=w.device)
    n = lanes // 4
    k0 = (lanes % 4) * 2
    kidx = torch.stack([k0, k0 + 1, k0 + 8, k0 + 9], dim=1)                                    # [32][4]
    frag = wp[:, :, :, kidx, n[:, None].expand(32, 4)]                                         # [G][9][Cg/16][32][4]
    return frag.half().contiguous()


class _Conv:
    """One packed convolution (weights for both arithmetic paths)."""

    def __init__(self, w, b, cin_pad=None, tc=False, segs=None, simt=True):
        self.cout, self.cin = w.shape[0], cin_pad or w.shape[1]
        self.kh, self.kw = w.shape[2], w.shape[3]
        self.pad = self.kh // 2
        self.W = pack_conv_weight(w, cin_pad) if simt else None
        self.wstride = self.W.shape[1] if simt else 0
        self.bias = b.float().contiguous()
        self.Wh = pack_conv_weight_tc(w, segs) if tc else None


def conv2d(cv, segs, nimg, H, W, dtype, out, ostride, ooff=0, epilogue=0, out2=None, aux0=None, aux_stride=0,
           aux1=None, use_tc=False, stride=1):
    """segs: list of (tensor_or_ptr, channels, stride).  All pointers are device pointers."""
    L = _lib.lib()
    n = len(segs)
    sp = (c_void_p * n)(*[s[0] if isinstance(s[0], int) else s[0].data_ptr() for s in segs])
    sc = _lib.int_array([s[1] for s in segs])
    ss = _lib.int_array([s[2] for s in segs])
    if use_tc:
        check(L.pvo_conv2d_tc(sp, sc, ss, n, ptr(cv.Wh), ptr(cv.bias), cv.cout, cv.kh, cv.kw, nimg, H, W, epilogue,
                              ptr(out), ostride, ooff, ptr(out2), ptr(aux0), aux_stride, ptr(aux1), stream()),
              'conv2d_tc')
    else:
        check(L.pvo_conv2d(sp, sc, ss, n, ptr(cv.W), cv.wstride, ptr(cv.bias), cv.cout, cv.kh, cv.kw, cv.pad, nimg, H,
                           W, dtype, epilogue, ptr(out), ostride, ooff, ptr(out2), ptr(aux0), aux_stride, ptr(aux1), stride,
                           stream()), 'conv2d')


class UpdateEngine:
    """Fused DynamicUpdateModule.forward on persistent pixel-major buffers.

    sd: state dict with the reference's parameter names relative to ``DroidNet.update`` (``gru.convz.weight`` ...).
    """

    def __init__(self, sd, mode='fp32', device='cuda', use_tc=None):
        assert mode in ('fp32', 'fp16')
        self.mode = mode
        self.device = torch.device(device)
        self.tdtype = torch.float32 if mode == 'fp32' else torch.float16
        self.dcode = _lib.F32 if mode == 'fp32' else _lib.F16
        self.use_tc = (mode == 'fp16') if use_tc is None else (use_tc and mode == 'fp16')
        g = lambda k: sd[k].detach().to(self.device)
        tc = self.use_tc
        self.c_corr0 = _Conv(g('corr_encoder.0.weight'), g('corr_encoder.0.bias'), cin_pad=CORR_STRIDE, tc=tc)
        self.c_corr2 = _Conv(g('corr_encoder.2.weight'), g('corr_encoder.2.bias'), tc=tc)
        self.c_flow0 = _Conv(g('flow_encoder.0.weight'), g('flow_encoder.0.bias'))
        if tc:
            # 7x7 over 8 channels == 7x1 over the x-expanded 64-channel input (pvo_expand_x7): W'[co][dxi*8+c][kh]
            w7 = g('flow_encoder.0.weight')                                    # (128, 8, 7, 7) = (co, c, kh, kw)
            wx = torch.zeros(128, 64, 7, 1, device=w7.device, dtype=w7.dtype)
            wx[:, :56, :, 0] = w7.permute(0, 3, 1, 2).reshape(128, 56, 7)       # (co, kw, c, kh) -> (co, kw*8+c, kh)
            self.c_flow0x = _Conv(wx, g('flow_encoder.0.bias'), tc=True)
        self.c_flow2 = _Conv(g('flow_encoder.2.weight'), g('flow_encoder.2.bias'), tc=tc)
        self.c_w = _Conv(g('gru.w.weight'), g('gru.w.bias'), tc=tc)
        self.c_zr = _Conv(torch.cat([g('gru.convz.weight'), g('gru.convr.weight')], 0),
                          torch.cat([g('gru.convz.bias'), g('gru.convr.bias')], 0), tc=tc)
        self.c_q = _Conv(g('gru.convq.weight'), g('gru.convq.bias'), tc=tc, segs=[128, 320])
        self.glo_W = torch.cat([g('gru.conv%s_glo.weight' % n).reshape(128, 128) for n in 'zrq'], 0).float().contiguous()
        self.glo_b = torch.cat([g('gru.conv%s_glo.bias' % n) for n in 'zrq'], 0).float().contiguous()
        # the four heads' first convs as one 128->512 conv; agg.conv1 (128->128) separately: 640 columns would leave the third
        # 256-wide tensor-core tile half empty, 512 + 128 tile exactly (and the 128-wide launch shares weight tiles across M tiles)
        self.c_heads1 = _Conv(torch.cat([g(h + '.0.weight') for h in HEAD_NAMES], 0),
                              torch.cat([g(h + '.0.bias') for h in HEAD_NAMES], 0), tc=tc)
        self.c_agg1 = _Conv(g('agg.conv1.weight'), g('agg.conv1.bias'), tc=tc)
        self.heads2_W = pack_small_weight([g(h + '.2.weight') for h in HEAD_NAMES])
        self.heads2_b = torch.cat([g(h + '.2.bias') for h in HEAD_NAMES], 0).float().contiguous()
        self.heads2_frag = pack_heads_mma([g(h + '.2.weight') for h in HEAD_NAMES]) if mode == 'fp16' else None
        self.c_heads2t = None
        if mode == 'fp16' and tc:
            # tensor-core form of the 3x3 128->2 head convs: block-diagonal 1x1 conv onto (tap, output) pairs + a 9-tap
            # gather-sum (csrc/heads_stencil.cu).  Row g*32 + tap*2 + o, column g*128 + c  <-  w_g[o, c, tap]
            wbd = torch.zeros(128, 512, 1, 1, device=self.device)
            for gi, h in enumerate(HEAD_NAMES):
                w2 = g(h + '.2.weight').float()                                 # (2, 128, 3, 3)
                wbd[gi * 32:gi * 32 + 18, gi * 128:(gi + 1) * 128, 0, 0] = w2.permute(2, 3, 0, 1).reshape(18, 128)
            self.c_heads2t = _Conv(wbd, torch.zeros(128, device=self.device), tc=True, simt=False, segs=[128, 128, 128, 128])
        self.c_agg2 = _Conv(g('agg.conv2.weight'), g('agg.conv2.bias'), tc=tc)
        self.eta_W = pack_small_weight([g('agg.eta.0.weight')])
        self.eta_b = g('agg.eta.0.bias').float().contiguous()
        self.c_upmask = _Conv(g('agg.upmask_disp.0.weight'), g('agg.upmask_disp.0.bias'), tc=tc)
        self.E = self.K = self.H = self.W = 0
        self.timers = None            # dict -> per-launch CUDA events (bench.py)
        self.generation = 0           # bumped by every run(): lets lazily evaluated outputs notice that their inputs are gone

    def fork(self):
        """An engine that shares the packed weights of this one but owns its activation buffers and hidden state.  Every
        FactorGraph keeps its edges' hidden state resident in its engine between iterations, and the module's own forward()
        (MotionFilter's one-step flow estimate) runs between those iterations: they must not share buffers."""
        import copy
        e = copy.copy(self)
        e.E = e.K = e.H = e.W = 0          # the first reserve() allocates fresh buffers
        e.timers = None
        e.base = self
        return e

    # ------------------------------------------------------------------ buffers
    def reserve(self, E, K, H, W, nframes=None):
        """(Re)allocate the activation buffers for E edges, K source frames, HxW grid."""
        if (E, K, H, W) == (self.E, self.K, self.H, self.W):
            return
        dev, dt = self.device, self.tdtype
        M, HW = E * H * W, H * W
        z = lambda *s, dtype=dt: torch.zeros(*s, dtype=dtype, device=dev)
        self.X = z(M, 448)             # [net | inp | corr_enc | flow_enc]   (gru.py:20-21 channel order)
        self.corr = z(M, CORR_STRIDE)
        self.motn = z(M, 8)
        self.motx = z(M, 64) if self.use_tc else None      # x-expanded motion features (tensor-core flow encoder)
        self.col_scratch = z(E * 32 * 128, dtype=torch.float32)
        self.glo_part = z(E * ((H + 7) // 8) * ((W + 15) // 16) * 4 * 128, dtype=torch.float32) if self.use_tc else None
        self.c1 = z(M, 128)
        self.f1 = z(M, 128)
        self.glo_tmp = z(M, 128)
        self.glo = z(E, 128, dtype=torch.float32)
        self.glo3 = z(E, 384, dtype=torch.float32)
        self.z = z(M, 128)
        self.rnet = z(M, 128)
        self.Hb = z(M, 640)            # relu(first conv) of the 4 heads | relu(agg.conv1)
        self.head_out = z(M, 8, dtype=torch.float32)       # [delta | delta_dy | weight | delta_mask]
        self.T2 = z(M, 128) if self.c_heads2t is not None else None   # per-tap partial outputs of the heads' second conv
        self.A = z(K * HW, 128)
        self.a2 = z(K * HW, 128)
        self.eta = z(K, H, W, dtype=torch.float32)
        self.upmask = None
        nf = nframes or 1024
        self.g_frame2k = z(nf, dtype=torch.int32)
        self.g_kx = z(max(K, 1), dtype=torch.int32)
        self.g_ptr = z(max(K, 1) + 1, dtype=torch.int32)
        self.g_idx = z(max(E, 1), dtype=torch.int32)
        self.g_meta = z(8, dtype=torch.int32)
        self.nframes = nf
        self.E, self.K, self.H, self.W = E, K, H, W

    def group(self, ii):
        """torch.unique(ii) + edges grouped by source frame, on the device (droid_net.py:83)."""
        check(_lib.lib().pvo_group_edges(ptr(ii), self.E, self.nframes, 0, 0, max(self.K, 1), ptr(self.g_frame2k),
                                         ptr(self.g_kx), ptr(self.g_ptr), ptr(self.g_idx), ptr(self.g_meta), stream()),
              'group_edges')

    # ------------------------------------------------------------------ layout conversion at the API boundary
    def load_nchw(self, src, dst, dstride, doff):
        """(N,C,H,W) any float dtype -> dst[(n*HW+p)*dstride + doff + c]."""
        N, C, H, W = src.shape
        check(_lib.lib().pvo_nchw_to_nhwc(ptr(src), _lib.dtype_code(src), ptr(dst), _lib.dtype_code(dst), N, C, H * W,
                                          dstride, doff, stream()), 'nchw_to_nhwc')

    def store_nchw(self, src, sstride, soff, C, N, out_dtype=None):
        out = torch.empty(N, C, self.H, self.W, dtype=out_dtype or src.dtype, device=self.device)
        check(_lib.lib().pvo_nhwc_to_nchw(ptr(src), _lib.dtype_code(src), sstride, soff, ptr(out), _lib.dtype_code(out),
                                          N, C, self.H * self.W, stream()), 'nhwc_to_nchw')
        return out

    # ------------------------------------------------------------------ the operator
    def run(self, with_agg=True):
        """X[:, :256] (net|inp), corr, motn must be loaded.  Updates X[:, :128] (net) in place and fills head_out,
        eta (when with_agg).  ~14 launches, no host sync."""
        L = _lib.lib()
        T = lambda name: _lib.timed(self.timers, name)
        self.generation += 1
        E, K, H, W, dc, tc = self.E, self.K, self.H, self.W, self.dcode, self.use_tc
        HW = H * W
        X = self.X
        esz = X.element_size()
        # corr_encoder: 1x1 196(208)->128 + ReLU, 3x3 128->128 + ReLU  -> X[:,256:384]
        with T('conv_corr0'):
            conv2d(self.c_corr0, [(self.corr, CORR_STRIDE, CORR_STRIDE)], E, H, W, dc, self.c1, 128, 0, 1, use_tc=tc)
        with T('conv_corr2'):
            conv2d(self.c_corr2, [(self.c1, 128, 128)], E, H, W, dc, X, 448, 256, 1, use_tc=tc)
        # flow_encoder: 7x7 8->128 + ReLU, 3x3 128->64 + ReLU  -> X[:,384:448]
        if tc:
            with T('expand_x7'):
                check(L.pvo_expand_x7(ptr(self.motn), ptr(self.motx), E, H, W, stream()), 'expand_x7')
            with T('conv_flow0'):
                conv2d(self.c_flow0x, [(self.motx, 64, 64)], E, H, W, dc, self.f1, 128, 0, 1, use_tc=True)
        else:
            with T('conv_flow0'):
                conv2d(self.c_flow0, [(self.motn, 8, 8)], E, H, W, dc, self.f1, 128, 0, 1)
        with T('conv_flow2'):
            conv2d(self.c_flow2, [(self.f1, 128, 128)], E, H, W, dc, X, 448, 384, 1, use_tc=tc)
        # global context: mean_HW(sigmoid(w(net)) * net) -> three 128x128 mat-vecs
        if tc:       # gate conv on tensor cores; its epilogue reduces over each 8x16 tile, the tail finishes mean + mat-vecs
            tiles = ((H + 7) // 8) * ((W + 15) // 16)
            with T('conv_glo'):
                conv2d(self.c_w, [(X, 128, 448)], E, H, W, dc, self.glo_part, 128, 0, 2, aux0=X, aux_stride=448, use_tc=True)
            with T('glo_matvec'):
                check(L.pvo_glo_reduce_matvec(ptr(self.glo_part), tiles * 4, HW, ptr(self.glo_W), ptr(self.glo_b), ptr(self.glo3),
                                              E, stream()), 'glo_reduce_matvec')
        else:
            with T('conv_glo'):
                conv2d(self.c_w, [(X, 128, 448)], E, H, W, dc, self.glo_tmp, 128, 0, 2, aux0=X, aux_stride=448)
            with T('colmean'):
                check(L.pvo_colmean(ptr(self.glo_tmp), ptr(self.glo), ptr(self.col_scratch), E, HW, 128, dc, stream()), 'colmean',
                      launches=2)
            with T('glo_matvec'):
                check(L.pvo_glo_matvec(ptr(self.glo), ptr(self.glo_W), ptr(self.glo_b), ptr(self.glo3), E, stream()), 'glo_matvec')
        # z|r in one 448->256 conv; epilogue emits z and r*net
        with T('conv_zr'):
            conv2d(self.c_zr, [(X, 448, 448)], E, H, W, dc, self.z, 128, 0, 3, out2=self.rnet, aux0=X, aux_stride=448,
                   aux1=self.glo3, use_tc=tc)
        # q on [r*net | inp | corr | flow]; epilogue applies net = (1-z) net + z q in place
        with T('conv_q'):
            conv2d(self.c_q, [(self.rnet, 128, 128), (X.data_ptr() + 128 * esz, 320, 448)], E, H, W, dc, X, 448, 0, 4,
                   aux0=self.z, aux1=self.glo3, use_tc=tc)
        # 4 heads' first convs (128->512) and agg.conv1 (128->128), + ReLU, into Hb[:, :512] / Hb[:, 512:]
        with T('conv_heads1'):
            conv2d(self.c_heads1, [(X, 128, 448)], E, H, W, dc, self.Hb, 640, 0, 1, use_tc=tc)
            if with_agg:
                conv2d(self.c_agg1, [(X, 128, 448)], E, H, W, dc, self.Hb, 640, 512, 1, use_tc=tc)
        with T('heads2'):
            if self.c_heads2t is not None:
                hb = self.Hb.data_ptr()
                conv2d(self.c_heads2t, [(hb + gi * 128 * esz, 128, 640) for gi in range(4)], E, H, W, dc, self.T2, 128, 0, 0,
                       use_tc=True)
                check(L.pvo_heads_stencil(ptr(self.T2), ptr(self.heads2_b), ptr(self.head_out), 8, E, H, W, stream()), 'heads_stencil')
            else:
                check(L.pvo_conv3x3_small(ptr(self.Hb), 640, 128, ptr(self.heads2_W), ptr(self.heads2_b), ptr(self.head_out), 8,
                                          4, 2, E, H, W, dc, 0, ptr(self.heads2_frag), stream()), 'heads2')
        if with_agg:
            with T('segment_mean'):
                check(L.pvo_segment_mean(ptr(self.Hb), 640, 512, ptr(self.A), ptr(self.g_ptr), ptr(self.g_idx),
                                         ptr(self.g_meta), K, HW, 128, dc, stream()), 'segment_mean')
            with T('conv_agg2'):
                conv2d(self.c_agg2, [(self.A, 128, 128)], K, H, W, dc, self.a2, 128, 0, 1, use_tc=tc)
            with T('eta'):
                check(L.pvo_conv3x3_small(ptr(self.a2), 128, 128, ptr(self.eta_W), ptr(self.eta_b), ptr(self.eta), 1, 1, 1,
                                          K, H, W, dc, 1, None, stream()), 'eta')

    def run_upmask(self):
        """GraphAgg.upmask_disp (droid_net.py:76-77,93): only DroidNet.forward consumes it; FactorGraph.update
        discards it (factor_graph.py:242), so it is computed on demand."""
        K, H, W = self.K, self.H, self.W
        if self.upmask is None or self.upmask.shape[0] != K * H * W:
            self.upmask = torch.empty(K * H * W, 576, dtype=self.tdtype, device=self.device)
        conv2d(self.c_upmask, [(self.a2, 128, 128)], K, H, W, self.dcode, self.upmask, 576, 0, 0, use_tc=self.use_tc)
        return self.upmask


def gru_forward(sd, net, inp):
    """Stand-alone ConvGRU.forward (modules/gru.py:19-32) for (N,128,h,w) net and (N,320,h,w) inp."""
    _lib.require_cuda(net.contiguous(), inp.contiguous())
    N, _, H, W = net.shape
    mode = 'fp16' if net.dtype == torch.float16 else 'fp32'
    dt = torch.float16 if mode == 'fp16' else torch.float32
    dc = _lib.F16 if mode == 'fp16' else _lib.F32
    dev = net.device
    g = lambda k: sd[k].detach().to(dev)
    c_w = _Conv(g('gru.w.weight'), g('gru.w.bias'))
    c_zr = _Conv(torch.cat([g('gru.convz.weight'), g('gru.convr.weight')], 0),
                 torch.cat([g('gru.convz.bias'), g('gru.convr.bias')], 0))
    c_q = _Conv(g('gru.convq.weight'), g('gru.convq.bias'))
    glo_W = torch.cat([g('gru.conv%s_glo.weight' % n).reshape(128, 128) for n in 'zrq'], 0).float().contiguous()
    glo_b = torch.cat([g('gru.conv%s_glo.bias' % n) for n in 'zrq'], 0).float().contiguous()
    M, HW = N * H * W, H * W
    L = _lib.lib()
    X = torch.empty(M, 448, dtype=dt, device=dev)
    check(L.pvo_nchw_to_nhwc(ptr(net.contiguous()), _lib.dtype_code(net), ptr(X), dc, N, 128, HW, 448, 0, stream()), 'nchw_to_nhwc')
    check(L.pvo_nchw_to_nhwc(ptr(inp.contiguous()), _lib.dtype_code(inp), ptr(X), dc, N, 320, HW, 448, 128, stream()), 'nchw_to_nhwc')
    tmp = torch.empty(M, 128, dtype=dt, device=dev)
    glo = torch.empty(N, 128, dtype=torch.float32, device=dev)
    glo3 = torch.empty(N, 384, dtype=torch.float32, device=dev)
    z = torch.empty(M, 128, dtype=dt, device=dev)
    rnet = torch.empty(M, 128, dtype=dt, device=dev)
    conv2d(c_w, [(X, 128, 448)], N, H, W, dc, tmp, 128, 0, 2, aux0=X, aux_stride=448)
    scratch = torch.empty(N * 32 * 128, dtype=torch.float32, device=dev)
    check(L.pvo_colmean(ptr(tmp), ptr(glo), ptr(scratch), N, HW, 128, dc, stream()), 'colmean', launches=2)
    check(L.pvo_glo_matvec(ptr(glo), ptr(glo_W), ptr(glo_b), ptr(glo3), N, stream()), 'glo_matvec')
    conv2d(c_zr, [(X, 448, 448)], N, H, W, dc, z, 128, 0, 3, out2=rnet, aux0=X, aux_stride=448, aux1=glo3)
    conv2d(c_q, [(rnet, 128, 128), (X.data_ptr() + 128 * X.element_size(), 320, 448)], N, H, W, dc, X, 448, 0, 4,
           aux0=z, aux1=glo3)
    out = torch.empty(N, 128, H, W, dtype=net.dtype, device=dev)
    check(L.pvo_nhwc_to_nchw(ptr(X), dc, 448, 0, ptr(out), _lib.dtype_code(out), N, 128, HW, stream()), 'nhwc_to_nchw')
    return out
