"""Host-side camera math: turns camera parameters into the per-view inputs of the visibility path.

In an engine integration these numbers come straight from the reference's Camera
(Camera::GetView Camera.cpp:585-595, Camera::GetProjection :648-691, Camera::GetFrustum :284-308);
this module exists so that the standalone bench/tests can make the same inputs without the engine.
Everything is evaluated in float32.  The results are *inputs* to the hot path ("libm is used
only on the host side ... those results are inputs passed to the GPU verbatim"), so they are fed byte-for-byte
to the CUDA path, the C oracle and the compiled reference alike.
"""
import numpy as np

f32 = np.float32
M_DEGTORAD = f32(np.pi / 180.0)
M_DEGTORAD_2 = f32(np.pi / 360.0)


def quat_from_euler(x_deg, y_deg, z_deg):
    """(w,x,y,z); rotation order Z, X, Y as Quaternion::FromEulerAngles (Quaternion.cpp:49-66)."""
    x = f32(x_deg) * M_DEGTORAD_2
    y = f32(y_deg) * M_DEGTORAD_2
    z = f32(z_deg) * M_DEGTORAD_2
    sx, cx = f32(np.sin(x)), f32(np.cos(x))
    sy, cy = f32(np.sin(y)), f32(np.cos(y))
    sz, cz = f32(np.sin(z)), f32(np.cos(z))
    w_ = cy * cx * cz + sy * sx * sz
    x_ = cy * sx * cz + sy * cx * sz
    y_ = sy * cx * cz - cy * sx * sz
    z_ = cy * cx * sz - sy * sx * cz
    return np.array([w_, x_, y_, z_], dtype=f32)


def quat_to_matrix3(q):
    """Quaternion::RotationMatrix (Quaternion.cpp:234-247)."""
    w, x, y, z = [f32(v) for v in q]
    two = f32(2.0)
    one = f32(1.0)
    return np.array([
        [one - two * y * y - two * z * z, two * x * y - two * w * z, two * x * z + two * w * y],
        [two * x * y + two * w * z, one - two * x * x - two * z * z, two * y * z - two * w * x],
        [two * x * z - two * w * y, two * y * z + two * w * x, one - two * x * x - two * y * y],
    ], dtype=f32)


def world_transform(pos, quat):
    """3x4 world transform of an unscaled node."""
    m = np.zeros((3, 4), dtype=f32)
    m[:, :3] = quat_to_matrix3(quat)
    m[:, 3] = np.asarray(pos, dtype=f32)
    return m


def inverse_rigid(m):
    """Inverse of an unscaled 3x4 transform (what Camera::GetView needs): R^T, -R^T t."""
    r = m[:, :3].T.astype(f32)
    t = m[:, 3].astype(f32)
    out = np.zeros((3, 4), dtype=f32)
    out[:, :3] = r
    for i in range(3):
        out[i, 3] = -(r[i, 0] * t[0] + r[i, 1] * t[1] + r[i, 2] * t[2])
    return out


def projection(fov_deg, aspect, near, far, ortho=False, ortho_size=20.0, zoom=1.0):
    """4x4 D3D-style projection, z in [0,1] (Camera::UpdateProjection Camera.cpp:648-691)."""
    p = np.zeros((4, 4), dtype=f32)
    near, far, zoom, aspect = f32(near), f32(far), f32(zoom), f32(aspect)
    if not ortho:
        h = (f32(1.0) / f32(np.tan(f32(fov_deg) * M_DEGTORAD * f32(0.5)))) * zoom
        w = h / aspect
        q = far / (far - near)
        r = -q * near
        p[0, 0], p[1, 1], p[2, 2], p[2, 3], p[3, 2] = w, h, q, r, f32(1.0)
    else:
        h = (f32(1.0) / (f32(ortho_size) * f32(0.5))) * zoom
        w = h / aspect
        p[0, 0], p[1, 1], p[2, 2], p[2, 3], p[3, 3] = w, h, f32(1.0) / far, f32(0.0), f32(1.0)
    return p


def _plane_from_points(v0, v1, v2):
    """Plane::Define(v0,v1,v2) (Plane.h:65-81): normal = normalize((v1-v0) x (v2-v0)), d = -n.v0."""
    a = (v1 - v0).astype(f32)
    b = (v2 - v0).astype(f32)
    n = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=f32)
    len_sq = n[0] * n[0] + n[1] * n[1] + n[2] * n[2]
    if len_sq > 0 and abs(float(len_sq) - 1.0) > 1e-6:
        n = (n * (f32(1.0) / f32(np.sqrt(len_sq)))).astype(f32)
    d = -(n[0] * v0[0] + n[1] * v0[1] + n[2] * v0[2])
    return np.array([n[0], n[1], n[2], d], dtype=f32)


def frustum_planes(fov_deg, aspect, near, far, xform, ortho=False, ortho_size=20.0, zoom=1.0):
    """6 planes (nx,ny,nz,d), order NEAR,LEFT,RIGHT,UP,DOWN,FAR (Frustum.h:37-45),
    built like Frustum::Define/DefineOrtho + UpdatePlanes (Frustum.cpp:75-104, 136-149, 225-244)."""
    near = f32(max(float(near), 0.0))
    far = f32(max(float(far), float(near)))
    if not ortho:
        half = f32(np.tan(f32(fov_deg) * M_DEGTORAD_2)) / f32(zoom)
        ny = near * half
        nx = ny * f32(aspect)
        fy = far * half
        fx = fy * f32(aspect)
    else:
        half = f32(ortho_size) * f32(0.5) / f32(zoom)
        ny = fy = half
        nx = fx = ny * f32(aspect)
    local = np.array([
        [nx, ny, near], [nx, -ny, near], [-nx, -ny, near], [-nx, ny, near],
        [fx, fy, far], [fx, -fy, far], [-fx, -fy, far], [-fx, fy, far]], dtype=f32)
    v = np.zeros((8, 3), dtype=f32)
    for i in range(8):
        for r in range(3):
            v[i, r] = xform[r, 0] * local[i, 0] + xform[r, 1] * local[i, 1] + xform[r, 2] * local[i, 2] + xform[r, 3]
    planes = np.stack([
        _plane_from_points(v[2], v[1], v[0]),
        _plane_from_points(v[3], v[7], v[6]),
        _plane_from_points(v[1], v[5], v[4]),
        _plane_from_points(v[0], v[4], v[7]),
        _plane_from_points(v[6], v[5], v[1]),
        _plane_from_points(v[5], v[6], v[7]),
    ]).astype(f32)
    n = planes[0]
    if n[0] * v[5, 0] + n[1] * v[5, 1] + n[2] * v[5, 2] + n[3] < 0:
        planes = -planes
    return planes


class CameraView:
    """Per-view inputs: view (3x4), projection (4x4), frustum planes (6x4), near/far, reverseCulling."""

    def __init__(self, view, proj, planes, near, far, reverse_culling=False, position=(0.0, 0.0, 0.0), ortho=False):
        self.view = np.ascontiguousarray(view, dtype=f32)
        self.proj = np.ascontiguousarray(proj, dtype=f32)
        self.planes = np.ascontiguousarray(planes, dtype=f32)
        self.near = float(near)
        self.far = float(far)
        self.reverse_culling = bool(reverse_culling)
        self.position = np.ascontiguousarray(position, dtype=f32)  # camera node world position (Camera::GetDistance)
        self.ortho = bool(ortho)
        self.half_view_size = 1.0  # Camera::GetHalfViewSize(); make_view fills it in
        self.ortho_size = 20.0     # Camera::GetOrthoSize() (DEFAULT_ORTHOSIZE, Camera.cpp:39)


def make_view(pos, yaw_deg, pitch_deg, fov_deg, aspect, near, far, ortho=False, ortho_size=20.0, zoom=1.0):
    q = quat_from_euler(pitch_deg, yaw_deg, 0.0)
    xf = world_transform(pos, q)
    if ortho:
        near = 0.0  # Camera::GetNearClip() reports projNearClip_, which stays 0 for orthographic cameras (Camera.cpp:676-678)
    cv = CameraView(inverse_rigid(xf), projection(fov_deg, aspect, near, far, ortho, ortho_size, zoom),
                    frustum_planes(fov_deg, aspect, near, far, xf, ortho, ortho_size, zoom), near, far, position=pos, ortho=ortho)
    # Camera::GetHalfViewSize (Camera.cpp:479-485) and GetOrthoSize: inputs of View::UpdateOccluders' size heuristic
    if not ortho:
        cv.half_view_size = float(f32(np.tan(f32(f32(f32(fov_deg) * f32(np.pi / 180.0)) * f32(0.5)))) / f32(zoom))
    else:
        cv.half_view_size = float(f32(f32(ortho_size) * f32(0.5)) / f32(zoom))
    cv.ortho_size = float(f32(ortho_size))
    return cv


def view_proj(proj, view):
    """Matrix4 * Matrix3x4 with the reference's evaluation order (Matrix4.cpp:43-63), float32, no contraction."""
    p = np.asarray(proj, dtype=f32)
    v = np.asarray(view, dtype=f32)
    out = np.zeros((4, 4), dtype=f32)
    for i in range(4):
        for j in range(3):
            out[i, j] = f32(f32(p[i, 0] * v[0, j]) + f32(p[i, 1] * v[1, j])) + f32(p[i, 2] * v[2, j])
        out[i, 3] = f32(f32(f32(p[i, 0] * v[0, 3]) + f32(p[i, 1] * v[1, 3])) + f32(p[i, 2] * v[2, 3])) + p[i, 3]
    return out
