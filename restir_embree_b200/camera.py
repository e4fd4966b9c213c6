"""Host-side mirror of the reference's pinhole camera (P/camera.{h,cpp}).

Only the part the ReSTIR path consumes: position, view matrix (glm::lookAt with the
re-orthogonalised up vector of Camera::recalculate_m_c_w, P/camera.cpp:44-58), its
inverse, and the focal length in pixels (Camera::setFOV, :81-84). z is up
(P/camera.h:68). All arithmetic in float32, glm operation order.
"""
import numpy as np

from .abi import RbCamera

f32 = np.float32


def _normalize(v):
    v = v.astype(f32)
    d = f32(v[0] * v[0]) + f32(v[1] * v[1]) + f32(v[2] * v[2])
    return (v * (f32(1.0) / np.sqrt(d, dtype=f32))).astype(f32)


def _cross(x, y):
    return np.array([x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]], dtype=f32)


def _dot(a, b):
    return f32(f32(a[0] * b[0]) + f32(a[1] * b[1]) + f32(a[2] * b[2]))


class Camera:
    """Camera(width, height, fov_y_degrees, view_from, view_at) — ctor of P/camera.cpp:12-18."""

    up = np.array([0.0, 0.0, 1.0], dtype=f32)

    def __init__(self, width, height, fov_y, view_from, view_at):
        self.width, self.height = int(width), int(height)
        self.view_from = np.asarray(view_from, dtype=f32)
        self.view_at = np.asarray(view_at, dtype=f32)
        self.setFOV(fov_y)
        self.recalculate_m_c_w()

    def setFOV(self, fov_deg):
        self.fov_y = f32(np.radians(f32(fov_deg)))
        self.f_y = f32(f32(self.height) / (f32(2.0) * np.tan(self.fov_y / f32(2.0), dtype=f32)))

    def setPosition(self, pos):
        self.view_from = np.asarray(pos, dtype=f32)
        self.recalculate_m_c_w()

    def recalculate_m_c_w(self):
        z_c = _normalize(self.view_from - self.view_at)
        x_c = _normalize(_cross(self.up, z_c))
        y_c = _normalize(_cross(z_c, x_c))
        # glm::lookAtRH(eye, center, up = y_c), P/glm/ext/matrix_transform.inl:99-119
        f = _normalize(self.view_at - self.view_from)
        s = _normalize(_cross(f, y_c))
        u = _cross(s, f)
        m = np.identity(4, dtype=f32)  # m[col][row]
        m[0][0], m[1][0], m[2][0] = s
        m[0][1], m[1][1], m[2][1] = u
        m[0][2], m[1][2], m[2][2] = -f
        m[3][0] = -_dot(s, self.view_from)
        m[3][1] = -_dot(u, self.view_from)
        m[3][2] = _dot(f, self.view_from)
        self.viewMat = m  # column-major: m[c] is column c
        # rigid transform: inverse = [R^T | -R^T t]; evaluated in float64 then rounded once
        M = m.T.astype(np.float64)  # row-major math view
        self.invViewMat = np.linalg.inv(M).T.astype(f32)

    def getPosition(self):
        return self.view_from

    def getViewMat(self):
        return self.viewMat

    def getInvViewMat(self):
        return self.invViewMat

    def getFocalLength(self):
        return self.f_y

    def to_abi(self):
        c = RbCamera()
        for i in range(3):
            c.pos[i] = float(self.view_from[i])
        c.focal_px = float(self.f_y)
        vm = self.viewMat.reshape(-1)
        iv = self.invViewMat.reshape(-1)
        for i in range(16):
            c.viewMat[i] = float(vm[i])
            c.invViewMat[i] = float(iv[i])
        return c
