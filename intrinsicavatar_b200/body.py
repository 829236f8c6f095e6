"""Synthetic SMPL-like body: host-side stand-in for the licensed SMPL ``.pkl``.

The reference drives its deformer from ``smplx.SMPL`` (reference
models/deformers/smplx/body_models.py:292-370, lbs.py:152-248,345-401), whose
data file is licensed and absent here.  This module keeps
the same *outputs* the render path consumes -- ``vertices[1,V,3]``,
``joints[1,24,3]``, ``A[1,24,4,4]`` (per-joint rigid transforms relative to the
rest joints, translation added to the last column) and ``lbs_weights[V,24]`` --
on the SMPL kinematic tree, from a procedurally generated capsule body.

Everything here is small host math (24 4x4 matrices per frame), numpy only.
"""
from __future__ import annotations

import numpy as np

# SMPL kinematic tree (reference models/pose/pose_encoder.py:29-57 lists the same parents).
PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.int64,
)

# Approximate rest (T-pose) joint locations of a neutral adult, metres.
REST_JOINTS = np.array(
    [
        [0.000, -0.240, 0.030],   # 0 pelvis
        [0.060, -0.330, 0.020],   # 1 l_hip
        [-0.060, -0.330, 0.020],  # 2 r_hip
        [0.000, -0.130, 0.000],   # 3 spine1
        [0.100, -0.710, 0.020],   # 4 l_knee
        [-0.100, -0.710, 0.020],  # 5 r_knee
        [0.000, 0.010, 0.020],    # 6 spine2
        [0.090, -1.110, -0.020],  # 7 l_ankle
        [-0.090, -1.110, -0.020], # 8 r_ankle
        [0.000, 0.060, 0.040],    # 9 spine3
        [0.120, -1.170, 0.100],   # 10 l_foot
        [-0.120, -1.170, 0.100],  # 11 r_foot
        [0.000, 0.270, 0.000],    # 12 neck
        [0.080, 0.180, 0.010],    # 13 l_collar
        [-0.080, 0.180, 0.010],   # 14 r_collar
        [0.000, 0.350, 0.050],    # 15 head
        [0.170, 0.210, 0.000],    # 16 l_shoulder
        [-0.170, 0.210, 0.000],   # 17 r_shoulder
        [0.430, 0.200, -0.020],   # 18 l_elbow
        [-0.430, 0.200, -0.020],  # 19 r_elbow
        [0.680, 0.210, -0.010],   # 20 l_wrist
        [-0.680, 0.210, -0.010],  # 21 r_wrist
        [0.770, 0.200, -0.020],   # 22 l_hand
        [-0.770, 0.200, -0.020],  # 23 r_hand
    ],
    dtype=np.float64,
)

# Capsule radius of the segment that follows each joint.
_RADII = np.array(
    [0.13, 0.085, 0.085, 0.13, 0.06, 0.06, 0.135, 0.045, 0.045, 0.14, 0.04, 0.04,
     0.06, 0.07, 0.07, 0.10, 0.055, 0.055, 0.042, 0.042, 0.035, 0.035, 0.03, 0.03]
)

N_VERTS = 6890  # same count as SMPL so the K=30 voxelisation recipe sees similar density


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Axis-angle [..., 3] -> rotation matrices [..., 3, 3] (lbs.py batch_rodrigues semantics)."""
    rvec = np.asarray(rvec, dtype=np.float64)
    angle = np.linalg.norm(rvec + 1e-8, axis=-1, keepdims=True)
    axis = rvec / angle
    c = np.cos(angle)[..., None]
    s = np.sin(angle)[..., None]
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    zeros = np.zeros_like(x)
    K = np.stack([zeros, -z, y, z, zeros, -x, -y, x, zeros], axis=-1).reshape(rvec.shape[:-1] + (3, 3))
    eye = np.broadcast_to(np.eye(3), K.shape)
    return eye + s * K + (1.0 - c) * (K @ K)


def rigid_chain(rot_mats: np.ndarray, joints: np.ndarray):
    """Kinematic chain, restating lbs.py:345-401 ``batch_rigid_transform`` for batch 1.

    Returns posed joints [24,3] and ``A`` [24,4,4]: world transform of each joint
    with the rest joint location removed (so ``A @ [x_rest,1]`` poses a rest point).
    """
    n = joints.shape[0]
    rel = joints.copy()
    rel[1:] -= joints[PARENTS[1:]]
    local = np.tile(np.eye(4), (n, 1, 1))
    local[:, :3, :3] = rot_mats
    local[:, :3, 3] = rel
    chain = [local[0]]
    for i in range(1, n):
        chain.append(chain[PARENTS[i]] @ local[i])
    G = np.stack(chain)
    posed = G[:, :3, 3].copy()
    A = G.copy()
    A[:, :3, 3] -= np.einsum("nij,nj->ni", G[:, :3, :3], joints)
    return posed, A


def _segments():
    """One capsule per joint: from the joint towards the mean of its children (stub for leaves)."""
    segs = []
    for j in range(24):
        kids = np.nonzero(PARENTS == j)[0]
        a = REST_JOINTS[j]
        if len(kids):
            b = REST_JOINTS[kids].mean(0)
        else:
            par = REST_JOINTS[PARENTS[j]]
            d = a - par
            b = a + 0.6 * d
        segs.append((a, b))
    return segs


def _point_segment_dist(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=-1)


class SyntheticBody:
    """Drop-in for the three things the deformer takes from ``smplx.SMPL``.

    ``__call__(body_pose, global_orient, transl)`` mirrors the subset of
    ``SMPL.forward`` the reference uses (snarf_deformer.py:55,95-104): returns a
    dict with ``vertices``, ``joints``, ``A`` for batch size 1 (float32).
    """

    def __init__(self, seed: int = 0, n_verts: int = N_VERTS):
        rng = np.random.RandomState(seed)
        segs = _segments()
        lengths = np.array([np.linalg.norm(b - a) + 2 * r for (a, b), r in zip(segs, _RADII)])
        areas = lengths * _RADII
        counts = np.maximum((areas / areas.sum() * n_verts).astype(int), 8)
        counts[0] += n_verts - counts.sum()
        verts = []
        for (a, b), r, c in zip(segs, _RADII, counts):
            axis = b - a
            L = np.linalg.norm(axis)
            axis = axis / max(L, 1e-9)
            # orthonormal frame around the capsule axis
            tmp = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
            u = np.cross(axis, tmp)
            u /= np.linalg.norm(u)
            v = np.cross(axis, u)
            s = rng.uniform(-r, L + r, size=c)          # position along the axis incl. caps
            phi = rng.uniform(0, 2 * np.pi, size=c)
            sc = np.clip(s, 0, L)
            cap = s - sc                                  # signed overshoot into a cap
            rad = np.sqrt(np.maximum(r * r - cap * cap, 0.0))
            pts = (a[None] + (sc + cap)[:, None] * axis[None]
                   + rad[:, None] * (np.cos(phi)[:, None] * u[None] + np.sin(phi)[:, None] * v[None]))
            verts.append(pts)
        verts = np.concatenate(verts, 0)
        # skinning weights: soft assignment by distance to every capsule axis
        d = np.stack([_point_segment_dist(verts, a, b) for (a, b) in segs], axis=1)  # [V,24]
        d = np.maximum(d - _RADII[None] * 0.5, 1e-3)
        w = np.exp(-0.5 * (d / 0.05) ** 2) + 1e-12
        # keep the 4 largest influences per vertex, like a rigged mesh
        idx = np.argsort(-w, axis=1)[:, 4:]
        np.put_along_axis(w, idx, 0.0, axis=1)
        w /= w.sum(1, keepdims=True)
        self.v_template = verts.astype(np.float32)
        self.lbs_weights = w.astype(np.float32)
        self.joints_rest = REST_JOINTS.copy()

    def __call__(self, body_pose=None, global_orient=None, transl=None):
        body_pose = np.zeros(69) if body_pose is None else np.asarray(body_pose, np.float64).reshape(69)
        global_orient = np.zeros(3) if global_orient is None else np.asarray(global_orient, np.float64).reshape(3)
        transl = np.zeros(3) if transl is None else np.asarray(transl, np.float64).reshape(3)
        pose = np.concatenate([global_orient, body_pose]).reshape(24, 3)
        R = rodrigues(pose)
        posed_joints, A = rigid_chain(R, self.joints_rest)
        T = np.einsum("vj,jab->vab", self.lbs_weights.astype(np.float64), A)
        vh = np.concatenate([self.v_template.astype(np.float64), np.ones((len(self.v_template), 1))], 1)
        verts = np.einsum("vab,vb->va", T, vh)[:, :3]
        A = A.copy()
        A[:, :3, 3] += transl
        return {
            "vertices": (verts + transl)[None].astype(np.float32),
            "joints": (posed_joints + transl)[None].astype(np.float32),
            "A": A[None].astype(np.float32),
        }


def a_pose() -> np.ndarray:
    """Canonical A-pose body pose (reference snarf_deformer.py:9-21, ``a_pose`` branch)."""
    p = np.zeros(69, dtype=np.float32)
    p[2] = 0.2
    p[5] = -0.2
    p[47] = -0.8
    p[50] = 0.8
    return p


class SMPLBody:
    """The real thing: SMPL linear blend skinning from a model file, same call contract as ``SyntheticBody``
    (SURVEY.md 8f.3).  Restates ``lbs`` (reference models/deformers/smplx/lbs.py:152-248: shape blend shapes ->
    joint regression -> pose-corrective blend shapes -> ``batch_rigid_transform`` -> skinning) and the translation
    handling of ``SMPL.forward`` (body_models.py:342-358: ``transl`` is added to vertices, joints and the last column
    of ``A``).  The licensed SMPL data is not shipped: ``from_file`` reads the official ``SMPL_*.pkl`` (or an ``.npz``
    with the same arrays); the constructor takes the arrays directly, which is what the parity test does with a
    random model of SMPL's shapes pushed through the reference's own ``lbs``.

      v_template [V,3], shapedirs [V,3,NB], posedirs [V,3,207] (or the reference's [207, V*3]), J_regressor [24,V],
      lbs_weights [V,24], parents [24]
    """

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents=PARENTS, betas=None):
        self.v_template = np.asarray(v_template, np.float64)
        V = self.v_template.shape[0]
        self.shapedirs = np.asarray(shapedirs, np.float64).reshape(V, 3, -1)
        pd = np.asarray(posedirs, np.float64)
        # body_models.py:156-159 stores posedirs as [P, V*3]; the .pkl holds [V,3,P]
        self.posedirs = pd if pd.ndim == 2 else pd.reshape(V * 3, -1).T
        self.J_regressor = np.asarray(J_regressor, np.float64)
        self.lbs_weights = np.asarray(lbs_weights, np.float32)
        self.parents = np.asarray(parents, np.int64)
        assert self.J_regressor.shape == (24, V) and self.lbs_weights.shape == (V, 24)
        assert self.posedirs.shape == (23 * 9, V * 3) and np.array_equal(self.parents[1:], PARENTS[1:])
        self.betas = np.zeros(self.shapedirs.shape[-1]) if betas is None else np.asarray(betas, np.float64).reshape(-1)

    @classmethod
    def from_file(cls, path: str, betas=None):
        """``SMPL_NEUTRAL.pkl`` (chumpy-free pickles or the official ones read with ``encoding='latin1'``) or ``.npz``."""
        if path.endswith(".npz"):
            d = dict(np.load(path, allow_pickle=True))
        else:
            import pickle
            with open(path, "rb") as f:
                d = pickle.load(f, encoding="latin1")
        def arr(x):
            x = getattr(x, "r", x)                       # chumpy arrays expose their value as .r
            return np.asarray(x.todense() if hasattr(x, "todense") else x)
        parents = arr(d["kintree_table"])[0].astype(np.int64) if "kintree_table" in d else PARENTS
        parents = parents.copy()
        parents[0] = -1
        return cls(arr(d["v_template"]), arr(d["shapedirs"])[..., :10], arr(d["posedirs"]), arr(d["J_regressor"]),
                   arr(d["weights"] if "weights" in d else d["lbs_weights"]), parents, betas)

    def __call__(self, body_pose=None, global_orient=None, transl=None, betas=None):
        body_pose = np.zeros(69) if body_pose is None else np.asarray(body_pose, np.float64).reshape(69)
        global_orient = np.zeros(3) if global_orient is None else np.asarray(global_orient, np.float64).reshape(3)
        transl = np.zeros(3) if transl is None else np.asarray(transl, np.float64).reshape(3)
        betas = self.betas if betas is None else np.asarray(betas, np.float64).reshape(-1)
        v_shaped = self.v_template + self.shapedirs[..., :len(betas)] @ betas          # blend_shapes
        J = self.J_regressor @ v_shaped                                                 # vertices2joints
        R = rodrigues(np.concatenate([global_orient, body_pose]).reshape(24, 3))
        pose_feature = (R[1:] - np.eye(3)).reshape(-1)
        v_posed = v_shaped + (pose_feature @ self.posedirs).reshape(-1, 3)
        posed_joints, A = rigid_chain(R, J)
        T = np.einsum("vj,jab->vab", self.lbs_weights.astype(np.float64), A)
        verts = np.einsum("vab,vb->va", T, np.concatenate([v_posed, np.ones((len(v_posed), 1))], 1))[:, :3]
        A = A.copy()
        A[:, :3, 3] += transl
        return {
            "vertices": (verts + transl)[None].astype(np.float32),
            "joints": (posed_joints + transl)[None].astype(np.float32),
            "A": A[None].astype(np.float32),
        }
