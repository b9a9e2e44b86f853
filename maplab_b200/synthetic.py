"""Seeded synthetic maps, queries and vocabularies for tests and bench.py.

Everything here is host-side numpy: it produces the INPUTS of the loop-closure path (binary
descriptors, keyframe headers, 3-D landmarks, keypoints, a quantizer file in maplab's
``common::Serialize`` format). None of it is on the measured path.
"""
import struct

import numpy as np

DESC_BYTES = 64  # 512-bit FREAK


# ----------------------------------------------------------------------------- vocabulary file
def serialize_matrix(m):
    """common::Serialize(Eigen::Matrix): int rows, int cols, column-major scalars
    (maplab-common/binary-serialization.h:128-161). m: [rows][cols] float32."""
    m = np.asarray(m, np.float32)
    return struct.pack("<ii", m.shape[0], m.shape[1]) + np.asfortranarray(m).tobytes(order="F")


def serialize_vocabulary(P, W1, W2, target_dim=None, version=100, pq=None):
    """InvertedMultiIndexVocabulary::Save layout (inverted-multi-index-interface.h:26-36);
    pq = (num_components, num_centers, dim_per_comp, Q1, Q2) appends the product block (:59-70)."""
    target_dim = target_dim if target_dim is not None else 2 * W1.shape[0]
    blob = struct.pack("<ii", version, target_dim)
    blob += serialize_matrix(P) + serialize_matrix(W1) + serialize_matrix(W2)
    if pq is not None:
        ncomp, ncent, dpc, Q1, Q2 = pq
        blob += struct.pack("<iiii", 200, ncomp, ncent, dpc) + serialize_matrix(Q1) + serialize_matrix(Q2)
    return blob


def parse_vocabulary(blob):
    off = 0

    def rd_int():
        nonlocal off
        v = struct.unpack_from("<i", blob, off)[0]
        off += 4
        return v

    def rd_mat():
        nonlocal off
        r, c = rd_int(), rd_int()
        m = np.frombuffer(blob, np.float32, r * c, off).reshape(c, r).T.copy()
        off += 4 * r * c
        return m

    version, dim = rd_int(), rd_int()
    return dict(version=version, target_dim=dim, P=rd_mat(), W1=rd_mat(), W2=rd_mat())


def unpack_bits(desc):
    """[n][bytes] uint8 -> [n][8*bytes] {0,1}, LSB first (descriptor-projection.h:92-115)."""
    return np.unpackbits(np.ascontiguousarray(desc, np.uint8), axis=1, bitorder="little")


def project_float(P, desc):
    """Plain fp32 projection used only to TRAIN synthetic vocabularies."""
    bits = unpack_bits(desc)[:, :P.shape[1]].astype(np.float32)
    return bits @ P.T.astype(np.float32)


def kmeans(x, k, iters, rng):
    x = np.asarray(x, np.float32)
    centers = x[rng.choice(len(x), size=k, replace=False)].copy()
    for _ in range(iters):
        assign = np.empty(len(x), np.int64)
        for s in range(0, len(x), 16384):
            xs = x[s:s + 16384]
            d = (xs * xs).sum(1)[:, None] - 2.0 * xs @ centers.T + (centers * centers).sum(1)[None]
            assign[s:s + 16384] = d.argmin(1)
        sums = np.zeros_like(centers, dtype=np.float64)
        np.add.at(sums, assign, x)
        cnt = np.bincount(assign, minlength=k)
        nz = cnt > 0
        centers[nz] = (sums[nz] / cnt[nz, None]).astype(np.float32)
        if (~nz).any():
            centers[~nz] = x[rng.choice(len(x), size=int((~nz).sum()), replace=False)]
    return centers


def make_vocabulary(train_desc, num_words=1000, target_dim=10, desc_bits=512, row_norm=2.0,
                    kmeans_iters=6, seed=7):
    """Random projection (rows of norm `row_norm`) + k-means words per half, trained on
    `train_desc` — the synthetic stand-in for train_projection_matrix / TrainProjectedVocabulary
    (offline trainers, out of scope). Returns (blob, dict)."""
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((target_dim, desc_bits)).astype(np.float32)
    P *= (row_norm / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32)
    y = project_float(P, train_desc)
    h = target_dim // 2
    W1 = kmeans(y[:, :h], min(num_words, len(y)), kmeans_iters, rng).T.copy()  # [h][W]
    W2 = kmeans(y[:, h:], min(num_words, len(y)), kmeans_iters, rng).T.copy()
    blob = serialize_vocabulary(P, W1, W2, target_dim)
    return blob, dict(P=P, W1=W1, W2=W2)


def add_product_quantizer(voc, num_components=10, num_centers=16, spread=0.6, seed=13):
    """Append a residual product quantiser (InvertedMultiIndexProductVocabulary,
    inverted-multi-index-interface.h:59-88) to a vocabulary dict of make_vocabulary: per coarse word
    and component `num_centers` seeded random centres (stand-in for TrainProjectedVocabulary's
    residual k-means, an offline trainer out of scope). Returns the version-200 blob."""
    rng = np.random.default_rng(seed)
    P, W1, W2 = voc["P"], voc["W1"], voc["W2"]
    half = num_components // 2
    dpc = W1.shape[0] // half
    assert half * dpc == W1.shape[0]
    Q1 = (spread * rng.standard_normal((dpc, half * num_centers * W1.shape[1]))).astype(np.float32)
    Q2 = (spread * rng.standard_normal((dpc, half * num_centers * W2.shape[1]))).astype(np.float32)
    return serialize_vocabulary(P, W1, W2, 2 * W1.shape[0], pq=(num_components, num_centers, dpc, Q1, Q2))


# ----------------------------------------------------------------------------- synthetic map
def _flip_mask(rng, n, nbytes, log2_inv_p):
    """Random bit mask with P(bit) = 2^-log2_inv_p (AND of independent uniform words)."""
    m = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    for _ in range(log2_inv_p - 1):
        m &= rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    return m


CAMERA = dict(fu=400.0, fv=400.0, cu=376.0, cv=240.0, width=752, height=480)
KF_STEP = 0.05     # metres between consecutive keyframes (camera moves along +x, looks along +z)
KF_WINDOW = 2      # a landmark is seen from keyframes within +-KF_WINDOW of its centre keyframe
OBS_PROB = 0.8


def make_map(num_landmarks, seed=1, desc_per_keyframe=500, flip_log2=6, num_missions=1,
             desc_bytes=DESC_BYTES):
    """Database side. Returns dict with
       frames (capi.FRAME_DTYPE fields as arrays), bits [N][desc_bytes], landmarks [N] int64,
       landmark_xyz [L][3], base [L][desc_bytes], kf_pos [KF][3], centre [L]."""
    rng = np.random.default_rng(seed)
    L = int(num_landmarks)
    span = 2 * KF_WINDOW + 1
    lm_per_kf = max(int(round(desc_per_keyframe / (span * OBS_PROB))), 1)
    num_kf = max((L + lm_per_kf - 1) // lm_per_kf, span)
    centre = (np.arange(L, dtype=np.int64) // lm_per_kf).clip(0, num_kf - 1)
    base = rng.integers(0, 256, size=(L, desc_bytes), dtype=np.uint8)
    xyz = np.empty((L, 3), np.float64)
    xyz[:, 0] = centre * KF_STEP + rng.uniform(-1.0, 1.0, L)
    xyz[:, 1] = rng.uniform(-1.5, 1.5, L)
    xyz[:, 2] = rng.uniform(4.0, 12.0, L)
    # observations: (keyframe, landmark) pairs
    obs_kf, obs_lm = [], []
    for off in range(-KF_WINDOW, KF_WINDOW + 1):
        seen = rng.random(L) < OBS_PROB
        kf = centre + off
        ok = seen & (kf >= 0) & (kf < num_kf)
        obs_kf.append(kf[ok])
        obs_lm.append(np.nonzero(ok)[0])
    obs_kf = np.concatenate(obs_kf)
    obs_lm = np.concatenate(obs_lm)
    order = np.lexsort((obs_lm, obs_kf))
    obs_kf, obs_lm = obs_kf[order], obs_lm[order]
    N = len(obs_lm)
    bits = base[obs_lm]
    for s in range(0, N, 1 << 20):
        bits[s:s + (1 << 20)] ^= _flip_mask(rng, min(1 << 20, N - s), desc_bytes, flip_log2)
    counts = np.bincount(obs_kf, minlength=num_kf).astype(np.int32)
    kf_ids = np.arange(num_kf, dtype=np.int64)
    kf_per_mission = (num_kf + num_missions - 1) // num_missions
    frames = dict(timestamp_ns=kf_ids * 1_000_000_000, vertex_id=kf_ids,
                  mission_id=kf_ids // kf_per_mission, frame_index=np.zeros(num_kf, np.int32),
                  num_descriptors=counts)
    kf_pos = np.stack([kf_ids * KF_STEP, np.zeros(num_kf), np.zeros(num_kf)], 1)
    return dict(frames=frames, bits=bits, landmarks=obs_lm.astype(np.int64), landmark_xyz=xyz,
                base=base, kf_pos=kf_pos, centre=centre, num_kf=num_kf, lm_per_kf=lm_per_kf,
                obs_kf=obs_kf)


def _rot_yaw_pitch(yaw, pitch):
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return Ry @ Rx


def make_queries(m, num_queries, seed=11, desc_per_keyframe=500, outlier_frac=0.2, flip_log2=6,
                 pixel_noise=0.8, query_mission=1_000_000):
    """Query keyframes of a held-out mission revisiting random places of map `m`.
    Returns dict: frames, bits [Nq][bytes], keypoints [Nq][2], true_landmark [Nq] (-1 = outlier),
    T_G_I [Q][3][4] ground-truth poses, revisit [Q]."""
    rng = np.random.default_rng(seed)
    num_kf, L = m["num_kf"], len(m["base"])
    desc_bytes = m["base"].shape[1]
    lm_per_kf = m["lm_per_kf"]
    n_true = int(round(desc_per_keyframe * (1 - outlier_frac)))
    n_out = desc_per_keyframe - n_true
    revisit = rng.integers(KF_WINDOW, max(num_kf - KF_WINDOW, KF_WINDOW + 1), size=num_queries)
    bits_all, kp_all, lm_all, poses, counts = [], [], [], [], []
    cam = CAMERA
    for q in range(num_queries):
        c = int(revisit[q])
        lo = max((c - KF_WINDOW) * lm_per_kf, 0)
        hi = min((c + KF_WINDOW + 1) * lm_per_kf, L)
        R = _rot_yaw_pitch(rng.uniform(-0.05, 0.05), rng.uniform(-0.03, 0.03))  # R_G_I
        t = np.array([c * KF_STEP, 0.0, 0.0]) + rng.uniform(-0.15, 0.15, 3)
        cand = rng.permutation(np.arange(lo, hi))
        pc = (m["landmark_xyz"][cand] - t) @ R            # points in the camera/body frame
        u = cam["fu"] * pc[:, 0] / pc[:, 2] + cam["cu"]
        v = cam["fv"] * pc[:, 1] / pc[:, 2] + cam["cv"]
        vis = (pc[:, 2] > 0.5) & (u >= 0) & (u < cam["width"]) & (v >= 0) & (v < cam["height"])
        cand, u, v = cand[vis][:n_true], u[vis][:n_true], v[vis][:n_true]
        nt = len(cand)
        d_true = m["base"][cand] ^ _flip_mask(rng, nt, desc_bytes, flip_log2)
        kp_true = np.stack([u, v], 1) + rng.normal(0, pixel_noise, (nt, 2))
        d_out = rng.integers(0, 256, size=(n_out, desc_bytes), dtype=np.uint8)
        kp_out = np.stack([rng.uniform(0, cam["width"], n_out), rng.uniform(0, cam["height"], n_out)], 1)
        perm = rng.permutation(nt + n_out)
        bits_all.append(np.concatenate([d_true, d_out])[perm])
        kp_all.append(np.concatenate([kp_true, kp_out])[perm])
        lm_all.append(np.concatenate([cand, -np.ones(n_out, np.int64)])[perm])
        poses.append(np.concatenate([R, t[:, None]], 1))
        counts.append(nt + n_out)
    qid = np.arange(num_queries, dtype=np.int64)
    frames = dict(timestamp_ns=(10_000_000 + qid) * 1_000_000_000, vertex_id=10_000_000 + qid,
                  mission_id=np.full(num_queries, query_mission, np.int64),
                  frame_index=np.zeros(num_queries, np.int32),
                  num_descriptors=np.asarray(counts, np.int32))
    return dict(frames=frames, bits=np.concatenate(bits_all), keypoints=np.concatenate(kp_all),
                true_landmark=np.concatenate(lm_all), T_G_I=np.stack(poses), revisit=revisit)


def camera_dict():
    return dict(fu=CAMERA["fu"], fv=CAMERA["fv"], cu=CAMERA["cu"], cv=CAMERA["cv"],
                R_B_C=np.eye(3), t_B_C=np.zeros(3))


# ----------------------------------------------------------------------------- BASELINE config 1
def _mt19937_canonical(seed, n):
    """n draws of std::uniform_real_distribution<double>(0, 1) over std::mt19937(seed) as libstdc++
    computes them (generate_canonical<double, 53>: two 32-bit outputs, low word first)."""
    raw = np.frombuffer(np.random.RandomState(seed).bytes(8 * n), "<u4").astype(np.float64).reshape(n, 2)
    return np.minimum((raw[:, 0] + raw[:, 1] * 4294967296.0) / 18446744073709551616.0, np.nextafter(1.0, 0.0))


def make_6dof_map(num_vertices=20, num_landmarks=500, seed=10, descriptor_seed=1):
    """The recipe of maplab's `vi-map-generator-6dof` (BASELINE.json config 1) restated:
    test/vi-map-generator-6dof/src/6dof-vi-map-gen.cc:25,30 (500 landmarks, 20 vertices),
    6dof-pose-graph-gen.cc:22-62 (pinhole fu = fv = 100, 640 x 480, zero distortion, camera turned 90 degrees
    about z against the IMU), :198-201 (one random 48-byte descriptor per landmark), :300-317 (every
    observation of a landmark carries the SAME descriptor with 30 deterministic bits OR-ed in: bit i % 8
    of byte i for i < 30), :323 (all frame timestamps 0, one mission — self loop-closure needs
    --lc_min_image_time_seconds=0), keypoints = exact projections, landmark positions without noise.
    Landmarks: algorithms/simulation/src/generic-path-generator.cc:326-353 drawn exactly like the
    reference draws them (std::mt19937(landmark_seed = 10), uniform_real_distribution, circle radius 10 m,
    5 m to the keypoints, 3 m variance, vertical factor 2). Not restated: the trajectory itself comes from a
    polynomial path file + RK4 IMU integration (mav_planning_utils, imu-integrator); here the vertices sit
    on the same 10 m circle, evenly spaced, with the roll / pitch / height excitation of
    6dof-test-trajectory-gen.cc:66-82 in closed form. Returns a dict like make_map plus keypoints, T_G_I
    and camera."""
    u = _mt19937_canonical(seed, 3 * num_landmarks).reshape(num_landmarks, 3)
    angle = u[:, 0] * 2 * np.pi
    radius = 10.0 + 5.0 + 3.0 * (u[:, 1] - 0.5)
    xyz = np.stack([radius * np.cos(angle), radius * np.sin(angle), 2.0 * 3.0 * (u[:, 2] - 0.5)], 1)
    rng = np.random.default_rng(descriptor_seed)
    base = rng.integers(0, 256, size=(num_landmarks, 48), dtype=np.uint8)
    for i in range(30):
        base[:, i % 48] |= np.uint8(1 << (i % 8))
    cam = dict(fu=100.0, fv=100.0, cu=320.0, cv=240.0, width=640, height=480)
    c, s = np.sqrt(0.5), np.sqrt(0.5)
    R_C_I = np.array([[c * c - s * s, -2 * c * s, 0], [2 * c * s, c * c - s * s, 0], [0, 0, 1.0]])  # C_q_I = (w, z) = (c, s)
    R_I_C = R_C_I.T
    duration = 33.0
    frames_bits, frames_kp, frames_lm, poses, counts = [], [], [], [], []
    for v in range(num_vertices):
        t = duration * v / num_vertices
        theta = 2 * np.pi * v / num_vertices
        yaw = theta + np.pi / 2
        roll = 0.75 / (0.5 * np.pi) * (1 - np.cos(0.5 * np.pi * t))     # integral of the x gyro excitation
        pitch = -0.75 / (0.5 * np.pi) * (1 - np.cos(0.5 * np.pi * t))   # y gyro excitation (phase pi)
        cz, sz, cr, sr, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
        Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        R_G_I = Rz @ Ry @ Rx
        p_G_I = np.array([10 * np.cos(theta), 10 * np.sin(theta), 5.0 / np.pi ** 2 * (1 - np.cos(np.pi * t))])
        R_G_C = R_G_I @ R_I_C
        pc = (xyz - p_G_I) @ R_G_C
        with np.errstate(divide="ignore", invalid="ignore"):
            uu = cam["fu"] * pc[:, 0] / pc[:, 2] + cam["cu"]
            vv = cam["fv"] * pc[:, 1] / pc[:, 2] + cam["cv"]
        vis = (pc[:, 2] > 0) & (uu >= 0) & (uu < cam["width"]) & (vv >= 0) & (vv < cam["height"])
        ids = np.nonzero(vis)[0]
        frames_bits.append(base[ids])
        frames_kp.append(np.stack([uu[ids], vv[ids]], 1))
        frames_lm.append(ids.astype(np.int64))
        poses.append(np.concatenate([R_G_I, p_G_I[:, None]], 1))
        counts.append(len(ids))
    vid = np.arange(num_vertices, dtype=np.int64)
    frames = dict(timestamp_ns=np.zeros(num_vertices, np.int64), vertex_id=vid,
                  mission_id=np.zeros(num_vertices, np.int64), frame_index=np.zeros(num_vertices, np.int32),
                  num_descriptors=np.asarray(counts, np.int32))
    camera = dict(fu=cam["fu"], fv=cam["fv"], cu=cam["cu"], cv=cam["cv"], R_B_C=R_I_C, t_B_C=np.zeros(3))
    return dict(frames=frames, bits=np.concatenate(frames_bits), keypoints=np.concatenate(frames_kp),
                landmarks=np.concatenate(frames_lm), landmark_xyz=xyz, base=base, T_G_I=np.stack(poses),
                camera=camera)
