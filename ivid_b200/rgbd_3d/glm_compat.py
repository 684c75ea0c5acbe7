"""The three PyGLM functions inference/sample.py needs (glm.lookAt / glm.perspective / glm.inverse,
sample.py:305-336), as float32 numpy matrices in mathematical orientation (m[row, col], p_cam = M @ p_world).
PyGLM matrices are accepted wherever a modelview is expected (see `as_matrix`)."""
import numpy as np


def vec3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


def radians(deg):
    return np.float32(np.deg2rad(deg))


def lookAt(eye, center, up):
    eye, center, up = (np.asarray(v, dtype=np.float32) for v in (eye, center, up))
    f = center - eye
    f = f / np.float32(np.sqrt(np.dot(f, f)))
    s = np.cross(f, up)
    s = s / np.float32(np.sqrt(np.dot(s, s)))
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def perspective(fovy, aspect, near, far):
    t = np.float32(np.tan(np.float32(fovy) / np.float32(2)))
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = np.float32(1) / (np.float32(aspect) * t)
    m[1, 1] = np.float32(1) / t
    m[2, 2] = -(np.float32(far) + np.float32(near)) / (np.float32(far) - np.float32(near))
    m[3, 2] = -np.float32(1)
    m[2, 3] = -(np.float32(2) * np.float32(far) * np.float32(near)) / (np.float32(far) - np.float32(near))
    return m


def inverse(m):
    return np.linalg.inv(as_matrix(m).astype(np.float64)).astype(np.float32)


def as_matrix(m) -> np.ndarray:
    """4x4 float32, mathematical orientation.  PyGLM mat4 objects index as m[col][row]; numpy arrays / nested lists
    are taken as m[row][col]."""
    if type(m).__module__.split(".")[0] in ("glm", "pyglm"):
        return np.array([[m[c][r] for c in range(4)] for r in range(4)], dtype=np.float32)
    a = np.asarray(m, dtype=np.float32)
    assert a.shape == (4, 4), "modelview must be a 4x4 matrix"
    return np.ascontiguousarray(a)
