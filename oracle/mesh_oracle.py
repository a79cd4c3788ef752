"""CPU restatement (numpy, float32, one rounding per operation) of the mesh rasteriser on the path of SURVEY 8 (f4):
tools/mvrenderer.py:33-178 as compute_surfelinfo.py:114-116 calls it.  TEST INFRASTRUCTURE ONLY (imported by tests/).

PARITY UNPINNED.  The algorithm lives in pytorch3d, a third-party dependency the reference neither vendors nor pins
(README.md:16: "Download PyTorch3D following the instruction here") and that is absent from this image, so nothing here could be
checked against the package itself.  What is restated is its published rasterisation path for the settings the reference uses
(RasterizationSettings(faces_per_pixel=1, blur_radius=0), PerspectiveCameras(in_ndc=False), mvrenderer.py:50-68):
  * cameras: screen-space focal length / principal point -> NDC (scale = min(H, W) / 2, +X left, +Y up; the reference's T_calib
    rotates the OpenCV pose by pi about z, mvrenderer.py:47-48,143-150); x_ndc = (f x + p z) / z; the z kept is the view depth;
  * rasterize_meshes (fine pass): pixel centres from pix_to_non_square_ndc with the image flipped in both axes; a face covers a
    pixel when all three barycentrics (edge functions / (area + 1e-8)) are > 0 and its area is not within 1e-8 of zero;
    perspective-correct barycentrics w_i = b_i z_j z_k / max(sum, 1e-8); depth = sum w_i z_i, discarded when < 0; the nearest
    face wins (equal depth: lower face index);
  * shading: interpolate_face_attributes with those barycentrics (vertex colours under AmbientLights = the colours themselves;
    NOCS coordinates of mvrenderer.py:695-722), softmax_rgb_blend(sigma = gamma = 1e-4, black background, znear 0, zfar 1e4):
    with one face per pixel the colour is prob c / (prob + 1e-10), prob = sigmoid(d^2 / sigma), d = distance to the nearest edge;
  * fragments.zbuf: -1 where no face covers the pixel.
"""
import numpy as np

f32 = np.float32
EPS = f32(1e-8)


def nocs_coordinates(verts):
    """mvrenderer.py:695-722."""
    v = verts.astype(np.float32)
    d = v - v.mean(axis=0, keepdims=True, dtype=np.float32)
    return ((d / np.abs(d).max(axis=0, keepdims=True) + f32(1)) / f32(2)).astype(np.float32)


def project(verts, pose_row, K, H, W):
    """[V,3] object space -> [V,3] (x_ndc, y_ndc, view depth), float32, one rounding per operation."""
    P = pose_row.astype(np.float32).reshape(12)
    K = K.astype(np.float32).reshape(9)
    x, y, z = (verts[:, i].astype(np.float32) for i in range(3))
    dot = lambda a, b, c, d: ((a * x + b * y) + c * z) + d
    xc, yc, zc = dot(P[0], P[1], P[2], P[3]), dot(P[4], P[5], P[6], P[7]), dot(P[8], P[9], P[10], P[11])
    s = f32(0.5) * f32(min(H, W))
    fx, fy = K[0] / s, K[4] / s
    px, py = -(K[2] - f32(0.5) * f32(W)) / s, -(K[5] - f32(0.5) * f32(H)) / s
    xn = (fx * (-xc) + px * zc) / zc
    yn = (fy * (-yc) + py * zc) / zc
    return np.stack([xn, yn, zc], axis=1).astype(np.float32)


def pix_to_ndc(i, S1, S2):
    rng = (f32(2.0) * f32(S1)) / f32(S2) if S1 > S2 else f32(2.0)
    off = rng / f32(S1)
    return (-(rng / f32(2.0)) + off / f32(2.0)) + off * f32(i)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _cover(px, py, v0, v1, v2):
    area = _edge(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1])
    if -EPS <= area <= EPS:
        return None
    a = area + EPS
    b0 = _edge(px, py, v1[0], v1[1], v2[0], v2[1]) / a
    b1 = _edge(px, py, v2[0], v2[1], v0[0], v0[1]) / a
    b2 = _edge(px, py, v0[0], v0[1], v1[0], v1[1]) / a
    if not (b0 > 0 and b1 > 0 and b2 > 0):
        return None
    t0, t1, t2 = (b0 * v1[2]) * v2[2], (v0[2] * b1) * v2[2], (v0[2] * v1[2]) * b2
    den = max((t0 + t1) + t2, EPS)
    w = (t0 / den, t1 / den, t2 / den)
    pz = (w[0] * v0[2] + w[1] * v1[2]) + w[2] * v2[2]
    if not pz >= 0:
        return None
    return w, pz


def _seg_dist2(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    if l2 <= EPS:
        return (px - bx) ** 2 + (py - by) ** 2
    t = min(max((bax * (px - ax) + bay * (py - ay)) / l2, f32(0)), f32(1))
    qx, qy = ax + t * bax, ay + t * bay
    return (px - qx) ** 2 + (py - qy) ** 2


def render(verts, faces, attr, pose_row, K, H, W, sigma=1e-4):
    """-> (out [C,H,W], depth [H,W], pix_to_face [H,W]).  Pure-Python loops: small meshes / images only."""
    with np.errstate(all="ignore"):
        ndc = project(verts, pose_row, K, H, W)
        C = attr.shape[1]
        depth = np.full((H, W), -1.0, np.float32)
        p2f = np.full((H, W), -1, np.int32)
        bary = np.zeros((H, W, 3), np.float32)
        s = 0.5 * min(H, W)
        xs = [pix_to_ndc(W - 1 - c, W, H) for c in range(W)]
        ys = [pix_to_ndc(H - 1 - r, H, W) for r in range(H)]
        for f, (i0, i1, i2) in enumerate(faces):
            v0, v1, v2 = ndc[i0], ndc[i1], ndc[i2]
            if not (v0[2] > 0 and v1[2] > 0 and v2[2] > 0):
                continue
            xmax, xmin = max(v0[0], v1[0], v2[0]), min(v0[0], v1[0], v2[0])
            ymax, ymin = max(v0[1], v1[1], v2[1]), min(v0[1], v1[1], v2[1])
            c0, c1 = max(0, int(np.floor(0.5 * W - xmax * s - 2.5))), min(W - 1, int(np.ceil(0.5 * W - xmin * s + 1.5)))
            r0, r1 = max(0, int(np.floor(0.5 * H - ymax * s - 2.5))), min(H - 1, int(np.ceil(0.5 * H - ymin * s + 1.5)))
            for r in range(r0, r1 + 1):
                for c in range(c0, c1 + 1):
                    hit = _cover(xs[c], ys[r], v0, v1, v2)
                    if hit is None:
                        continue
                    w, pz = hit
                    if p2f[r, c] < 0 or pz < depth[r, c]:      # (ascending f: an equal depth keeps the lower face index)
                        depth[r, c], p2f[r, c], bary[r, c] = pz, f, w
        out = np.zeros((C, H, W), np.float32)
        for r, c in zip(*np.nonzero(p2f >= 0)):
            i0, i1, i2 = faces[p2f[r, c]]
            v0, v1, v2 = ndc[i0], ndc[i1], ndc[i2]
            px, py = xs[c], ys[r]
            d2 = min(_seg_dist2(px, py, v0[0], v0[1], v1[0], v1[1]), _seg_dist2(px, py, v1[0], v1[1], v2[0], v2[1]),
                     _seg_dist2(px, py, v2[0], v2[1], v0[0], v0[1]))
            prob = f32(1) / (f32(1) + np.exp(-d2 / f32(sigma)))
            w = bary[r, c]
            a = (w[0] * attr[i0] + w[1] * attr[i1]) + w[2] * attr[i2]
            out[:, r, c] = (prob * a) / (prob + f32(1e-10))
        return out, depth, p2f


def icosphere(subdiv=2, radius=1.0):
    """Unit icosphere (vertices [V,3] float32, faces [F,3] int32) -- the synthetic CAD model of the tests."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1),
         (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius).astype(np.float32), np.array(f, np.int32)
