"""SE3 group operations on CPU (torch, any float dtype) -- oracle, test infrastructure only.

Restates the subset of lietorch that PVO's VO path uses.  Element layout is lietorch's:
``data[..., 7] = [tx, ty, tz, qx, qy, qz, qw]``.

Follows (reference file:line, all under VO_Module/thirdparty/lietorch/lietorch/):
  * include/so3.h:31-37    quaternion is normalised on every construction
  * include/so3.h:39-60    inv, mul, rotate-a-point
  * include/so3.h:123-175  Log (atan form), Exp
  * include/so3.h:177-214  left_jacobian, left_jacobian_inverse
  * include/se3.h:36-56    inv, mul, act4
  * include/se3.h:58-86    Adj / AdjT
  * include/se3.h:124-142  Log / Exp
  * groups.py:153-164      retr(a) = Exp(a) * X ; adjT
"""
import torch

EPS = 1e-6  # include/common.h:7


def _normalize_q(q):
    return q / torch.linalg.vector_norm(q, dim=-1, keepdim=True)


def quat_mul(a, b):
    """Hamilton product, (x, y, z, w) storage (Eigen convention)."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], dim=-1)


def quat_conj(q):
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def so3_act(q, p):
    """so3.h:54-60: p + w*uv + v x uv with uv = 2 (v x p)."""
    v = q[..., :3]
    w = q[..., 3:4]
    uv = torch.linalg.cross(v, p, dim=-1)
    uv = uv + uv
    return p + w * uv + torch.linalg.cross(v, uv, dim=-1)


def so3_matrix(q):
    """Rotation matrix of a unit quaternion (Eigen toRotationMatrix)."""
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = torch.stack([
        1 - (tyy + tzz), txy - twz, txz + twy,
        txy + twz, 1 - (txx + tzz), tyz - twx,
        txz - twy, tyz + twx, 1 - (txx + tyy)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def hat(phi):
    x, y, z = phi.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=-1).reshape(phi.shape[:-1] + (3, 3))


def so3_exp(phi):
    theta2 = (phi * phi).sum(-1, keepdim=True)
    theta = torch.sqrt(theta2)
    theta4 = theta2 * theta2
    small = theta < EPS
    safe = torch.where(small, torch.ones_like(theta), theta)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * theta2 + (1.0 / 3840.0) * theta4,
                       torch.sin(0.5 * safe) / safe)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * theta2 + (1.0 / 384.0) * theta4,
                       torch.cos(0.5 * safe))
    return _normalize_q(torch.cat([imag * phi, real], dim=-1))


def so3_log(q):
    v = q[..., :3]
    w = q[..., 3:4]
    n2 = (v * v).sum(-1, keepdim=True)
    n = torch.sqrt(n2)
    small = n2 < EPS * EPS
    n_safe = torch.where(small, torch.ones_like(n), n)
    w_small = w.abs() < EPS
    w_safe = torch.where(w_small, torch.ones_like(w), w)
    pi = torch.as_tensor(3.14159265358979323846, dtype=q.dtype)
    f_small = 2.0 / w_safe - (2.0 / 3.0) * n2 / (w_safe * w_safe * w_safe)
    f_wzero = torch.where(w > 0, pi / n_safe, -pi / n_safe)
    f_gen = 2.0 * torch.atan(n_safe / w_safe) / n_safe
    f = torch.where(small, f_small, torch.where(w_small, f_wzero, f_gen))
    return f * v


def so3_left_jacobian(phi):
    I = torch.eye(3, dtype=phi.dtype).expand(phi.shape[:-1] + (3, 3))
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(-1)[..., None, None]
    theta = torch.sqrt(theta2)
    small = theta < EPS
    t_safe = torch.where(small, torch.ones_like(theta), theta)
    t2_safe = torch.where(small, torch.ones_like(theta2), theta2)
    c1 = torch.where(small, 0.5 - (1.0 / 24.0) * theta2, (1.0 - torch.cos(t_safe)) / t2_safe)
    c2 = torch.where(small, 1.0 / 6.0 - (1.0 / 120.0) * theta2,
                     (t_safe - torch.sin(t_safe)) / (t2_safe * t_safe))
    return I + c1 * Phi + c2 * Phi2


def so3_left_jacobian_inverse(phi):
    I = torch.eye(3, dtype=phi.dtype).expand(phi.shape[:-1] + (3, 3))
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(-1)[..., None, None]
    theta = torch.sqrt(theta2)
    small = theta < EPS
    t_safe = torch.where(small, torch.ones_like(theta), theta)
    half = 0.5 * t_safe
    c2 = torch.where(small, torch.full_like(theta, 1.0 / 12.0),
                     (1.0 - t_safe * torch.cos(half) / (2.0 * torch.sin(half))) / (t_safe * t_safe))
    return I - 0.5 * Phi + c2 * Phi2


# ---------------------------------------------------------------- SE3 on raw (...,7) data

def _split(X):
    return X[..., :3], _normalize_q(X[..., 3:7])      # so3.h:35-37 normalises on construction


def se3_inv(X):
    t, q = _split(X)
    qi = _normalize_q(quat_conj(q))
    return torch.cat([-so3_act(qi, t), qi], dim=-1)


def se3_mul(X, Y):
    tx, qx = _split(X)
    ty, qy = _split(Y)
    return torch.cat([tx + so3_act(qx, ty), _normalize_q(quat_mul(qx, qy))], dim=-1)


def se3_act4(X, p):
    t, q = _split(X)
    return torch.cat([so3_act(q, p[..., :3]) + t * p[..., 3:4], p[..., 3:4]], dim=-1)


def se3_act3(X, p):
    t, q = _split(X)
    return so3_act(q, p) + t


def se3_adjT(X, a):
    """b = Adj(X)^T a, se3.h:58-86."""
    t, q = _split(X)
    Rt = so3_matrix(q).transpose(-1, -2)
    at, ar = a[..., :3], a[..., 3:]
    top = (Rt @ at[..., None])[..., 0]
    bot = (Rt @ (ar - torch.linalg.cross(t, at, dim=-1))[..., None])[..., 0]
    return torch.cat([top, bot], dim=-1)


def se3_adj(X, a):
    t, q = _split(X)
    R = so3_matrix(q)
    at, ar = a[..., :3], a[..., 3:]
    Rar = (R @ ar[..., None])[..., 0]
    top = (R @ at[..., None])[..., 0] + torch.linalg.cross(t, Rar, dim=-1)
    return torch.cat([top, Rar], dim=-1)


def se3_exp(xi):
    tau, phi = xi[..., :3], xi[..., 3:]
    q = so3_exp(phi)
    t = (so3_left_jacobian(phi) @ tau[..., None])[..., 0]
    return torch.cat([t, q], dim=-1)


def se3_log(X):
    t, q = _split(X)
    phi = so3_log(q)
    tau = (so3_left_jacobian_inverse(phi) @ t[..., None])[..., 0]
    return torch.cat([tau, phi], dim=-1)


def se3_retr(X, a):
    """groups.py:153-156: Exp(a) * X."""
    return se3_mul(se3_exp(a), X)


def se3_matrix(X):
    t, q = _split(X)
    T = torch.zeros(X.shape[:-1] + (4, 4), dtype=X.dtype)
    T[..., :3, :3] = so3_matrix(q)
    T[..., :3, 3] = t
    T[..., 3, 3] = 1
    return T


class SE3:
    """Minimal stand-in for ``lietorch.SE3`` (groups.py:51-231, 261-275) used to run the
    reference's own Python geometry (geom/*.py) on CPU when generating golden vectors."""
    group_name = 'SE3'
    group_id = 3
    manifold_dim = 6
    embedded_dim = 7
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data
        self.data = data

    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        if isinstance(batch_shape[0], (tuple, list)):
            batch_shape = tuple(batch_shape[0])
        data = cls.id_elem.to(kwargs.get('dtype', torch.float32)).repeat(tuple(batch_shape) + (1,))
        return cls(data)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, dtype=G.data.dtype)

    @classmethod
    def exp(cls, x):
        return cls(se3_exp(x))

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        if isinstance(batch_shape[0], (tuple, list)):
            batch_shape = tuple(batch_shape[0])
        return cls.exp(sigma * torch.randn(tuple(batch_shape) + (6,), **kwargs))

    @staticmethod
    def _bc(x, y):
        """lietorch broadcasting (broadcasting.py:10-31): same rank, 1-dims expand."""
        assert x.dim() == y.dim(), (x.shape, y.shape)
        shp = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
        return x.expand(shp + x.shape[-1:]), y.expand(shp + y.shape[-1:])

    def log(self):
        return se3_log(self.data)

    def inv(self):
        return SE3(se3_inv(self.data))

    def mul(self, other):
        a, b = self._bc(self.data, other.data)
        return SE3(se3_mul(a, b))

    def retr(self, a):
        a, d = self._bc(a, self.data)
        return SE3(se3_retr(d, a))

    def adj(self, a):
        d, a = self._bc(self.data, a)
        return se3_adj(d, a)

    def adjT(self, a):
        d, a = self._bc(self.data, a)
        return se3_adjT(d, a)

    def act(self, p):
        d, p = self._bc(self.data, p)
        return se3_act3(d, p) if p.shape[-1] == 3 else se3_act4(d, p)

    def matrix(self):
        return se3_matrix(self.data)

    def translation(self):
        return self.data[..., :3]

    def vec(self):
        return self.data

    def detach(self):
        return SE3(self.data.detach())

    def view(self, dims):
        return SE3(self.data.view(tuple(dims) + (7,)))

    def __mul__(self, other):
        if isinstance(other, SE3):
            return self.mul(other)
        return self.act(other)

    def __getitem__(self, index):
        return SE3(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def to(self, *a, **k):
        return SE3(self.data.to(*a, **k))

    def cpu(self):
        return self

    def clone(self):
        return SE3(self.data.clone())


def cat(group_list, dim):
    """lietorch.cat (groups.py:313-316)."""
    return SE3(torch.cat([g.data for g in group_list], dim=dim))


def stack(group_list, dim):
    return SE3(torch.stack([g.data for g in group_list], dim=dim))
