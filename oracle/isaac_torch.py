"""TEST INFRASTRUCTURE — restatement of third-party arithmetic the reference imports but does not vendor.

Package: Isaac Sim 2022.2.1 ``omni.isaac.core.utils.torch.rotations`` / ``.maths`` (version string at
/root/reference/RobotLearning/setup.py:25; no lock file).  Call sites in the reference:
quadruped_pose_control.py:7,319-350,430; quadruped_manipulate_plate.py:7,338-359,439;
utils/math.py:4,191; robot/base/robot.py:9,450-457.  All quaternions are **wxyz**.

PARITY UNPINNED for this file: the upstream source is not in the container and the reference holds
no test vectors for it.  The op order below is the published Isaac/IsaacGym ``torch_jit_utils`` form
(8-multiplication Hamilton product, ``a ± b + c`` rotate).  It is cross-checked only by algebraic
properties in tests/test_oracle_isaac.py (unit norm, q ⊗ conj(q) = 1, rotate ∘ rotate_inverse = id,
agreement with SciPy rotations).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (locomanipulationrl_b200/) never does.
"""
import torch


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    w1, x1, y1, z1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    w2, x2, y2, z2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([w, x, y, z], dim=-1).view(shape)


def quat_conjugate(a: torch.Tensor) -> torch.Tensor:
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((a[:, 0:1], -a[:, 1:]), dim=-1).view(shape)


def _dot3(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    n = a.shape[0]
    return torch.bmm(a.reshape(n, 1, 3), b.reshape(n, 3, 1)).squeeze(-1)


def quat_rotate(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q_w = q[:, 0]
    q_vec = q[:, 1:]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * _dot3(q_vec, v) * 2.0
    return a + b + c


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q_w = q[:, 0]
    q_vec = q[:, 1:]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * _dot3(q_vec, v) * 2.0
    return a - b + c


def quat_axis(q: torch.Tensor, axis: int = 0) -> torch.Tensor:
    basis_vec = torch.zeros(q.shape[0], 3, device=q.device)
    basis_vec[:, axis] = 1
    return quat_rotate(q, basis_vec)


def quat_apply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, 1:]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 0:1] * t + xyz.cross(t, dim=-1)).view(shape)


def quat_from_euler_xyz(roll: torch.Tensor, pitch: torch.Tensor, yaw: torch.Tensor) -> torch.Tensor:
    cy = torch.cos(yaw * 0.5)
    sy = torch.sin(yaw * 0.5)
    cr = torch.cos(roll * 0.5)
    sr = torch.sin(roll * 0.5)
    cp = torch.cos(pitch * 0.5)
    sp = torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qw, qx, qy, qz], dim=-1)


def unscale_transform(x: torch.Tensor, lower: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    offset = (lower + upper) * 0.5
    return x * (upper - lower) * 0.5 + offset


def set_seed(seed, torch_deterministic=False):
    torch.manual_seed(seed)
    return seed
