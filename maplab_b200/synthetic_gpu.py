"""Counter-based synthetic maps for the LARGE configurations (BASELINE.json configs 3-5: 10 M / 50 M
landmarks, 100 M descriptors), generated with torch on the device they are consumed on.

Same world model as maplab_b200/synthetic.py: landmarks along a corridor, 125 landmark
centres per keyframe, each landmark seen from the keyframes within +-2 of its centre keyframe with
probability 0.8, 512-bit descriptors = the landmark's base descriptor with every bit flipped with
probability 2^-6, queries of a held-out mission revisiting random places with 20 % outlier descriptors.
Every random quantity is a hash of (what it belongs to), so that any rank can generate any part of the
world — its shard of the database, its slice of the queries — without materialising the rest, and a
small world can be materialised on the CPU for the oracle. Input generation only: nothing here is on
the measured path.
"""
import numpy as np
import torch

from . import synthetic

KF_WINDOW = synthetic.KF_WINDOW
KF_STEP = synthetic.KF_STEP
OBS_PROB = synthetic.OBS_PROB
DESC_WORDS = 8  # 512 bits

_M1 = -0x61C8864680B583EB          # 0x9E3779B97F4A7C15 as int64
_M2 = -0x40A7B892E31B1A47          # 0xBF58476D1CE4E5B9
_M3 = -0x6B2FB644ECCEEE15          # 0x94D049BB133111EB


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64(x):
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around = the uint64 arithmetic)."""
    x = x + _M1
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def _hash(ids, stream):
    return splitmix64(ids * 0x100 + stream)


def _uniform(ids, stream):
    """U[0, 1) doubles from the hash (53 bits)."""
    return _lsr(_hash(ids, stream), 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def layout(num_landmarks, desc_per_keyframe=500):
    span = 2 * KF_WINDOW + 1
    lm_per_kf = max(int(round(desc_per_keyframe / (span * OBS_PROB))), 1)
    num_kf = max((num_landmarks + lm_per_kf - 1) // lm_per_kf, span)
    return lm_per_kf, num_kf


def landmark_words(lm):
    """Base descriptor of landmarks `lm` (int64 tensor [n]) as [n][8] int64 words."""
    w = torch.arange(DESC_WORDS, device=lm.device, dtype=torch.int64)
    return _hash(lm[:, None] * DESC_WORDS + w[None, :], 1)


def flip_words(ids, stream, log2_inv_p=6):
    """Random bit masks with P(bit) = 2^-log2_inv_p for descriptor ids `ids`: [n][8] int64."""
    w = torch.arange(DESC_WORDS, device=ids.device, dtype=torch.int64)
    base = (ids[:, None] * DESC_WORDS + w[None, :]) * 16
    m = _hash(base, stream)
    for r in range(1, log2_inv_p):
        m = m & _hash(base + r, stream)
    return m


def words_to_bytes(words):
    """[n][8] int64 -> [n][64] uint8 (little-endian words, as the descriptor bytes)."""
    return words.contiguous().view(torch.uint8).reshape(words.shape[0], 8 * DESC_WORDS)


def landmark_xyz(lm, lm_per_kf, num_kf):
    centre = torch.clamp(lm // lm_per_kf, max=num_kf - 1).to(torch.float64)
    x = centre * KF_STEP + (_uniform(lm, 2) * 2.0 - 1.0)
    y = _uniform(lm, 3) * 3.0 - 1.5
    z = _uniform(lm, 4) * 8.0 + 4.0
    return torch.stack([x, y, z], 1)


def observations(num_landmarks, kf0, kf1, device, desc_per_keyframe=500):
    """Observations of keyframes [kf0, kf1): (counts per keyframe [kf1-kf0] int64, landmark of every
    observation in (keyframe, landmark) order — int64 [sum(counts)])."""
    lm_per_kf, num_kf = layout(num_landmarks, desc_per_keyframe)
    span = 2 * KF_WINDOW + 1
    kf = torch.arange(kf0, kf1, device=device, dtype=torch.int64)
    j = torch.arange(span * lm_per_kf, device=device, dtype=torch.int64)
    lm = (kf[:, None] - KF_WINDOW) * lm_per_kf + j[None, :]
    valid = (lm >= 0) & (lm < num_landmarks)
    centre = torch.clamp(lm // lm_per_kf, min=0, max=num_kf - 1)
    # the last keyframe also owns the landmarks clipped to it; every landmark must sit within the window
    off = kf[:, None] - centre
    valid &= (off >= -KF_WINDOW) & (off <= KF_WINDOW)
    seen = _lsr(_hash(lm * 8 + (off + KF_WINDOW), 5), 11).to(torch.float64) < OBS_PROB * 9007199254740992.0
    mask = valid & seen
    return mask.sum(1), lm[mask]


def frames_for(kf0, counts, num_missions=1, num_kf=None):
    """Keyframe headers (host, capi.FRAME_DTYPE) of keyframes kf0 .. kf0 + len(counts)."""
    from . import capi
    n = len(counts)
    ids = np.arange(kf0, kf0 + n, dtype=np.int64)
    per_mission = ((num_kf or (kf0 + n)) + num_missions - 1) // num_missions
    return capi.make_frames(ids * 1_000_000_000, ids, ids // per_mission, np.zeros(n, np.int32),
                            np.asarray(counts, np.int32))


def descriptor_bytes(lm, gidx, log2_inv_p=6):
    """Database descriptors of observations (landmark lm, global descriptor index gidx): [n][64] uint8."""
    return words_to_bytes(landmark_words(lm) ^ flip_words(gidx, 6, log2_inv_p))


def build_database(det, num_landmarks, rank, world, device, chunk_kf=8192, num_missions=1, on_chunk=None,
                   all_rows=False):
    """Shard-aware build of the synthetic map into `det` (created with shard_rank = rank, shard_count =
    world): every rank walks the keyframe headers and landmark numbers of the whole map (replicated
    metadata), but generates, projects (kernel 1) and inserts only the descriptors its shard owns.
    all_rows: hand every descriptor to the detector (shard_mode 1, sharding by cell).
    Returns dict(num_descriptors, num_keyframes)."""
    lm_per_kf, num_kf = layout(num_landmarks)
    stream = torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0
    base = 0
    for kf0 in range(0, num_kf, chunk_kf):
        kf1 = min(num_kf, kf0 + chunk_kf)
        counts, lm = observations(num_landmarks, kf0, kf1, device)
        n = int(lm.shape[0])
        gidx = base + torch.arange(n, device=device, dtype=torch.int64)
        own = torch.ones_like(gidx, dtype=torch.bool) if all_rows else (gidx % world) == rank
        bits = descriptor_bytes(lm[own], gidx[own])
        proj = torch.empty((bits.shape[0], det.dim), dtype=torch.float32, device=device)
        if bits.shape[0]:
            det.project_device(bits.data_ptr(), 64, bits.shape[0], proj.data_ptr(), stream)
        frames = frames_for(kf0, counts.cpu().numpy(), num_missions, num_kf)
        det.insert_batch_device(frames, proj.data_ptr(), bits.shape[0], lm.data_ptr(), stream)
        if on_chunk is not None:
            on_chunk(kf0, frames, lm, gidx, own, bits, proj)
        base += n
    return dict(num_descriptors=base, num_keyframes=num_kf)


def vocabulary_sample(num_landmarks, n=100_000, device="cpu"):
    """Bytes of the first ~n database descriptors (host numpy) — the training set of the synthetic
    vocabulary; identical on every rank."""
    device = torch.device(device)
    lm_per_kf, num_kf = layout(num_landmarks)
    kf1 = min(num_kf, max(n // 400, 8))
    counts, lm = observations(num_landmarks, 0, kf1, device)
    gidx = torch.arange(lm.shape[0], device=device, dtype=torch.int64)
    return descriptor_bytes(lm[:n], gidx[:n]).cpu().numpy()


def make_queries(num_landmarks, q0, q1, device, seed=11, desc_per_keyframe=500, outlier_frac=0.2,
                 pixel_noise=0.8, query_mission=1_000_000):
    """Query keyframes q0 .. q1 of the held-out mission (any rank can make any slice). Returns dict:
    frames (host), bits [n][64] uint8 (device), keypoints [n][2] float64 (device), T_G_I [Q][3][4]
    (host, ground truth), true_landmark [n] int64 (device; -1 = outlier)."""
    from . import capi
    lm_per_kf, num_kf = layout(num_landmarks, desc_per_keyframe)
    cam = synthetic.CAMERA
    Q = q1 - q0
    n_true = int(round(desc_per_keyframe * (1 - outlier_frac)))
    qid = torch.arange(q0, q1, device=device, dtype=torch.int64) + seed * 1_000_003
    lo, hi = KF_WINDOW, max(num_kf - KF_WINDOW, KF_WINDOW + 1)
    revisit = lo + (_lsr(_hash(qid, 10), 1) % (hi - lo))
    yaw = _uniform(qid, 11) * 0.1 - 0.05
    pitch = _uniform(qid, 12) * 0.06 - 0.03
    t = torch.stack([revisit.to(torch.float64) * KF_STEP + (_uniform(qid, 13) * 0.3 - 0.15),
                     _uniform(qid, 14) * 0.3 - 0.15, _uniform(qid, 15) * 0.3 - 0.15], 1)
    cy, sy, cp, sp = torch.cos(yaw), torch.sin(yaw), torch.cos(pitch), torch.sin(pitch)
    zero, one = torch.zeros_like(cy), torch.ones_like(cy)
    Ry = torch.stack([torch.stack([cy, zero, sy], 1), torch.stack([zero, one, zero], 1),
                      torch.stack([-sy, zero, cy], 1)], 1)
    Rx = torch.stack([torch.stack([one, zero, zero], 1), torch.stack([zero, cp, -sp], 1),
                      torch.stack([zero, sp, cp], 1)], 1)
    R = Ry @ Rx                                                        # R_G_I [Q][3][3]
    span = (2 * KF_WINDOW + 1) * lm_per_kf
    j = torch.arange(span, device=device, dtype=torch.int64)
    cand = (revisit[:, None] - KF_WINDOW) * lm_per_kf + j[None, :]     # [Q][span]
    inside = (cand >= 0) & (cand < num_landmarks)
    cand_c = torch.clamp(cand, 0, num_landmarks - 1)
    xyz = landmark_xyz(cand_c.reshape(-1), lm_per_kf, num_kf).reshape(Q, span, 3)
    pc = torch.einsum("qsi,qij->qsj", xyz - t[:, None, :], R)          # points in the camera frame
    u = cam["fu"] * pc[..., 0] / pc[..., 2] + cam["cu"]
    v = cam["fv"] * pc[..., 1] / pc[..., 2] + cam["cv"]
    vis = inside & (pc[..., 2] > 0.5) & (u >= 0) & (u < cam["width"]) & (v >= 0) & (v < cam["height"])
    # the first n_true visible landmarks in a per-query random order
    order_key = _lsr(_hash(qid[:, None] * 4096 + j[None, :], 16), 2)
    order_key = torch.where(vis, order_key, torch.full_like(order_key, 2 ** 62))
    pick = torch.argsort(order_key, dim=1)[:, :n_true]                 # [Q][n_true]
    picked_vis = torch.gather(vis, 1, pick)
    lm_true = torch.gather(cand_c, 1, pick)
    u_t, v_t = torch.gather(u, 1, pick), torch.gather(v, 1, pick)
    # slots of the keyframe: [true ..., outliers ...] shuffled by a per-query permutation
    slot = torch.arange(desc_per_keyframe, device=device, dtype=torch.int64)
    did = qid[:, None] * 1024 + slot[None, :]                          # query descriptor ids [Q][500]
    is_true = torch.zeros((Q, desc_per_keyframe), dtype=torch.bool, device=device)
    is_true[:, :n_true] = picked_vis                                   # unseen picks become outliers
    lm_slot = torch.full((Q, desc_per_keyframe), -1, dtype=torch.int64, device=device)
    lm_slot[:, :n_true] = torch.where(picked_vis, lm_true, torch.full_like(lm_true, -1))
    flat = did.reshape(-1)
    words_true = landmark_words(torch.clamp(lm_slot.reshape(-1), min=0)) ^ flip_words(flat, 17)
    w8 = torch.arange(DESC_WORDS, device=device, dtype=torch.int64)
    words_out = _hash(flat[:, None] * DESC_WORDS + w8[None, :], 18)
    words = torch.where(is_true.reshape(-1)[:, None], words_true, words_out)
    # Box-Muller pixel noise on the true keypoints, uniform keypoints for the outliers
    r = torch.sqrt(-2.0 * torch.log(1.0 - _uniform(flat, 19)))
    ang = 2.0 * np.pi * _uniform(flat, 20)
    kp_u = torch.zeros((Q, desc_per_keyframe), dtype=torch.float64, device=device)
    kp_v = torch.zeros_like(kp_u)
    kp_u[:, :n_true], kp_v[:, :n_true] = u_t, v_t
    kp_u = kp_u.reshape(-1) + pixel_noise * r * torch.cos(ang)
    kp_v = kp_v.reshape(-1) + pixel_noise * r * torch.sin(ang)
    out_u = _uniform(flat, 21) * cam["width"]
    out_v = _uniform(flat, 22) * cam["height"]
    tr = is_true.reshape(-1)
    kp = torch.stack([torch.where(tr, kp_u, out_u), torch.where(tr, kp_v, out_v)], 1)
    perm = torch.argsort(_lsr(_hash(did, 23), 1), dim=1)               # [Q][500]
    gather = (perm + slot.new_tensor(desc_per_keyframe) * torch.arange(Q, device=device)[:, None]).reshape(-1)
    bits = words_to_bytes(words[gather])
    kp = kp[gather].contiguous()
    lm_out = lm_slot.reshape(-1)[gather]
    ids = np.arange(q0, q1, dtype=np.int64)
    frames = capi.make_frames((10_000_000 + ids) * 1_000_000_000, 10_000_000 + ids,
                              np.full(Q, query_mission, np.int64), np.zeros(Q, np.int32),
                              np.full(Q, desc_per_keyframe, np.int32))
    T = torch.cat([R, t[:, :, None]], 2).cpu().numpy()
    return dict(frames=frames, bits=bits, keypoints=kp, T_G_I=T, true_landmark=lm_out,
                revisit=revisit.cpu().numpy())


def all_landmark_xyz(num_landmarks, device, chunk=1 << 22):
    """Landmark positions of the whole map on `device` ([L][3] float64, 24 B per landmark)."""
    lm_per_kf, num_kf = layout(num_landmarks)
    out = torch.empty((num_landmarks, 3), dtype=torch.float64, device=device)
    for s in range(0, num_landmarks, chunk):
        e = min(num_landmarks, s + chunk)
        out[s:e] = landmark_xyz(torch.arange(s, e, device=device, dtype=torch.int64), lm_per_kf, num_kf)
    return out
