"""The subset of lietorch's SE3 API that PVO's VO path touches (thirdparty/lietorch/lietorch/groups.py:51-231,
261-275), as a light tensor wrapper.  Data layout (..., 7) = [tx ty tz qx qy qz qw].

The per-pixel group actions of the hot path (Gij * X, adjT for the Jacobians, retraction) live inside the fused CUDA
kernels (csrc/reproject.cu, csrc/ba.cu); this class only has to carry pose tensors through the reference's call
signatures (``lietorch.SE3(video.poses[None])``, ``Gs[:, ii]``, ``.inv()``, ``*``, ``.retr``, ``.log``, ``.adjT``,
``.data``, ``lietorch.cat``) and do the handful of O(#frames) operations the callers perform on them (pose interpolation in
the trajectory filler, the initial pose of a new keyframe).  Checked against lietorch's own property tests
(run_tests.py:16-52) in tests/test_host_logic_cpu.py.
"""
import torch


def _qnorm(q):
    return q / q.norm(dim=-1, keepdim=True)


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _qrot(q, p):
    v, w = q[..., :3], q[..., 3:]
    uv = 2 * torch.linalg.cross(v, p, dim=-1)
    return p + w * uv + torch.linalg.cross(v, uv, dim=-1)


class SE3:
    group_name, group_id, manifold_dim, embedded_dim = 'SE3', 3, 6, 7

    def __init__(self, data):
        self.data = data.data if isinstance(data, SE3) else data

    shape = property(lambda self: self.data.shape[:-1])
    device = property(lambda self: self.data.device)
    dtype = property(lambda self: self.data.dtype)

    @classmethod
    def Identity(cls, *batch_shape, **kw):
        if isinstance(batch_shape[0], (tuple, list)):
            batch_shape = tuple(batch_shape[0])
        d = torch.zeros(tuple(batch_shape) + (7,), **kw)
        d[..., 6] = 1
        return cls(d)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def exp(cls, xi):
        tau, phi = xi[..., :3], xi[..., 3:]
        th2 = (phi * phi).sum(-1, keepdim=True)
        th = th2.sqrt()
        small = th < 1e-6
        ths = torch.where(small, torch.ones_like(th), th)
        th2s = torch.where(small, torch.ones_like(th2), th2)
        imag = torch.where(small, 0.5 - th2 / 48 + th2 * th2 / 3840, torch.sin(0.5 * ths) / ths)
        real = torch.where(small, 1 - th2 / 8 + th2 * th2 / 384, torch.cos(0.5 * ths))
        c1 = torch.where(small, 0.5 - th2 / 24, (1 - torch.cos(ths)) / th2s)
        c2 = torch.where(small, 1.0 / 6 - th2 / 120, (ths - torch.sin(ths)) / (th2s * ths))
        a = torch.linalg.cross(phi, tau, dim=-1)
        b = torch.linalg.cross(phi, a, dim=-1)
        return cls(torch.cat([tau + c1 * a + c2 * b, _qnorm(torch.cat([imag * phi, real], -1))], -1))

    def inv(self):
        t, q = self.data[..., :3], _qnorm(self.data[..., 3:])
        qi = _qnorm(torch.cat([-q[..., :3], q[..., 3:]], -1))
        return SE3(torch.cat([-_qrot(qi, t), qi], -1))

    def mul(self, other):
        a, b = torch.broadcast_tensors(self.data, other.data)
        qa, qb = _qnorm(a[..., 3:]), _qnorm(b[..., 3:])
        return SE3(torch.cat([a[..., :3] + _qrot(qa, b[..., :3]), _qnorm(_qmul(qa, qb))], -1))

    def act(self, p):
        d = self.data
        shp = torch.broadcast_shapes(d.shape[:-1], p.shape[:-1])
        d, p = d.expand(shp + (7,)), p.expand(shp + p.shape[-1:])
        q = _qnorm(d[..., 3:])
        if p.shape[-1] == 3:
            return _qrot(q, p) + d[..., :3]
        return torch.cat([_qrot(q, p[..., :3]) + d[..., :3] * p[..., 3:], p[..., 3:]], -1)

    def retr(self, a):
        return SE3.exp(a).mul(self)

    def log(self):
        """Lie-algebra coordinates [tau, phi] (include/se3.h:124-131, so3.h:123-150): phi from the atan form of the
        quaternion log, tau = V^-1 t with V^-1 = I - 1/2 [phi]x + c [phi]x^2, c = (1 - th/2 cot(th/2)) / th^2"""
        t, q = self.data[..., :3], _qnorm(self.data[..., 3:])
        v, w = q[..., :3], q[..., 3:]
        n2 = (v * v).sum(-1, keepdim=True)
        n = n2.sqrt()
        small = n < 1e-6
        ns = torch.where(small, torch.ones_like(n), n)
        # 2 atan(n / w) / n, continued through w < 0 with atan2; series 2/w - 2 n^2 / (3 w^3) near n = 0
        k = torch.where(small, 2.0 / w - 2.0 * n2 / (3.0 * w * w * w), 2.0 * torch.atan2(ns * torch.sign(w), w.abs()) / ns)
        phi = k * v
        th2 = (phi * phi).sum(-1, keepdim=True)
        th = th2.sqrt()
        tiny = th < 1e-6
        ths = torch.where(tiny, torch.ones_like(th), th)
        c = torch.where(tiny, torch.full_like(th, 1.0 / 12) + th2 / 720, (1 - 0.5 * ths * torch.cos(0.5 * ths) / torch.sin(0.5 * ths)) / (ths * ths))
        a = torch.linalg.cross(phi, t, dim=-1)
        tau = t - 0.5 * a + c * torch.linalg.cross(phi, a, dim=-1)
        return torch.cat([tau, phi], -1)

    def adj(self, a):
        """Ad_X a for a = [tau, phi] (include/se3.h:58-70): [R tau + t x (R phi), R phi]"""
        d, a = self.data, a
        shp = torch.broadcast_shapes(d.shape[:-1], a.shape[:-1])
        d, a = d.expand(shp + (7,)), a.expand(shp + (6,))
        q, t = _qnorm(d[..., 3:]), d[..., :3]
        rphi = _qrot(q, a[..., 3:])
        return torch.cat([_qrot(q, a[..., :3]) + torch.linalg.cross(t, rphi, dim=-1), rphi], -1)

    def adjT(self, a):
        """Ad_X^T a (include/se3.h:72-86, groups.py:159-164): [R^T tau, R^T (phi - t x tau)]"""
        d, a = self.data, a
        shp = torch.broadcast_shapes(d.shape[:-1], a.shape[:-1])
        d, a = d.expand(shp + (7,)), a.expand(shp + (6,))
        q, t = _qnorm(d[..., 3:]), d[..., :3]
        qi = torch.cat([-q[..., :3], q[..., 3:]], -1)
        return torch.cat([_qrot(qi, a[..., :3]), _qrot(qi, a[..., 3:] - torch.linalg.cross(t, a[..., :3], dim=-1))], -1)

    def matrix(self):
        """(..., 4, 4) homogeneous transform (groups.py:122-127)"""
        d = self.data
        eye = torch.eye(3, dtype=d.dtype, device=d.device).expand(d.shape[:-1] + (3, 3))
        R = _qrot(_qnorm(d[..., 3:])[..., None, :], eye)                       # rows = R e_k  ->  R^T
        top = torch.cat([R.transpose(-1, -2), d[..., :3, None]], -1)
        bottom = torch.zeros(d.shape[:-1] + (1, 4), dtype=d.dtype, device=d.device)
        bottom[..., 0, 3] = 1
        return torch.cat([top, bottom], -2)

    def __mul__(self, other):
        return self.mul(other) if isinstance(other, SE3) else self.act(other)

    def __getitem__(self, index):
        return SE3(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def view(self, dims):
        return SE3(self.data.view(tuple(dims) + (7,)))

    def detach(self):
        return SE3(self.data.detach())

    def to(self, *a, **k):
        return SE3(self.data.to(*a, **k))

    def cuda(self):
        return SE3(self.data.cuda())

    def cpu(self):
        return SE3(self.data.cpu())

    def vec(self):
        return self.data

    def translation(self):
        return self.data[..., :3]


def cat(group_list, dim):
    return SE3(torch.cat([g.data for g in group_list], dim=dim))


def stack(group_list, dim):
    return SE3(torch.stack([g.data for g in group_list], dim=dim))
