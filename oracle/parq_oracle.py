"""CPU oracle: a restatement of the reference's recurrent pixel-aligned query
decoder (PARQ hot path).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module; the product path (parq_b200/) never
does and fails loudly when its CUDA library is missing.

Pinning: the reference has no tests or golden vectors for this path
(SURVEY.md 4, 8c), so the oracle is pinned against outputs of the unmodified
reference itself, run in the build container through oracle/ref_loader.py:
tests/golden/make_golden.py wrote tests/golden/*.npz and
tests/test_oracle_golden.py checks this file against them (bit-exact for the
projection, <=2e-5 for everything else, teacher-forced per iteration).

Arithmetic is plain fp32 torch on the CPU (the third-party arithmetic of the
reference *is* torch: F.grid_sample, softmax, layer/group norm, matmul), except
the pose/projection chain, which is restated with explicit per-operation fp32
rounding in numpy so that it is machine independent:
  * 3-term dot products inside Pose.inverse/compose round as
    ((a0*b0 + a1*b1) + a2*b2) with no FMA,
  * the point transform rounds as fma(p2,r2, fma(p1,r1, p0*r0)) then "+ t",
which is what torch 2.11 (MKL) does for these shapes on the build container
and reproduces the reference's center_im / center_valid bit for bit.

Reference citations are to /root/reference/<file>:<line>.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32
EPS_Z = F32(1e-3)          # Camera.eps, utils/wrappers.py:442

# data/average_scan2cad.txt through BoxProcessor.init_mean_size (utils/parq_utils.py:45-88):
# 8 Scan2CAD class means + "other" + "non-object" rows of ones; float64 in the reference.
MEAN_SIZE = np.array([
    [0.55067552, 0.84943989, 0.5786128],
    [1.24506049, 0.66165523, 0.72455878],
    [0.95658434, 0.99974904, 0.56246602],
    [0.36641966, 0.45580824, 0.27876528],
    [1.05132399, 1.3471979, 0.33744382],
    [0.60740744, 0.4752175, 0.16435075],
    [1.68820774, 0.76637348, 0.89351734],
    [0.85305378, 0.43925023, 0.51612006],
    [1.0, 1.0, 1.0],
    [1.0, 1.0, 1.0]], dtype=np.float64)


# --------------------------------------------------------------------------- #
# A.2  pose algebra with explicit fp32 rounding (utils/wrappers.py:247-267)
# --------------------------------------------------------------------------- #
def _dot3(a0, b0, a1, b1, a2, b2):
    """((a0*b0 + a1*b1) + a2*b2), every operation rounded to fp32, no FMA."""
    return (a0 * b0 + a1 * b1) + a2 * b2


def _fma(a, b, c):
    """fp32 fused multiply-add emulated through float64 (the product of two
    fp32 values is exact in float64; the final float64->fp32 rounding can differ
    from a true fma only in double-rounding corner cases of probability ~2^-29)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def pose_inverse(P):
    """Pose.inverse (wrappers.py:247-251): R' = R^T, t' = -(R^T t).  P: (...,12) fp32."""
    P = np.asarray(P, dtype=F32)
    R = P[..., :9].reshape(P.shape[:-1] + (3, 3))
    t = P[..., 9:]
    Rt = np.swapaxes(R, -1, -2)
    tt = np.stack([-_dot3(Rt[..., i, 0], t[..., 0], Rt[..., i, 1], t[..., 1], Rt[..., i, 2], t[..., 2])
                   for i in range(3)], -1)
    return np.concatenate([Rt.reshape(P.shape[:-1] + (9,)), tt], -1).astype(F32)


def pose_compose(A, B):
    """Pose.compose (wrappers.py:253-257): R = RA RB, t = tA + RA tB (broadcasting)."""
    A = np.asarray(A, dtype=F32)
    B = np.asarray(B, dtype=F32)
    shp = np.broadcast_shapes(A.shape, B.shape)
    A = np.broadcast_to(A, shp)
    B = np.broadcast_to(B, shp)
    RA = A[..., :9].reshape(shp[:-1] + (3, 3))
    RB = B[..., :9].reshape(shp[:-1] + (3, 3))
    tA, tB = A[..., 9:], B[..., 9:]
    R = np.empty(shp[:-1] + (3, 3), dtype=F32)
    for i in range(3):
        for j in range(3):
            R[..., i, j] = _dot3(RA[..., i, 0], RB[..., 0, j], RA[..., i, 1], RB[..., 1, j], RA[..., i, 2], RB[..., 2, j])
    t = np.stack([tA[..., i] + _dot3(RA[..., i, 0], tB[..., 0], RA[..., i, 1], tB[..., 1], RA[..., i, 2], tB[..., 2])
                  for i in range(3)], -1)
    return np.concatenate([R.reshape(shp[:-1] + (9,)), t], -1).astype(F32)


def camera_from_local(T_camera_pseudoCam, T_world_pseudoCam, T_world_local):
    """T_camera_local = T_cp @ (T_wp^-1 @ T_wl)  (transformer_parq.py:298-300).
    (B,T,12), (B,T,12), (B,1,12) -> (B,T,12) fp32 numpy."""
    return pose_compose(T_camera_pseudoCam, pose_compose(pose_inverse(T_world_pseudoCam), T_world_local))


def transform_points(T_camera_local, pts):
    """Pose.transform (wrappers.py:260-267): p @ R^T + t.  (B,T,12), (B,Nq,3) -> (B,T,Nq,3)."""
    Tcl = np.asarray(T_camera_local, dtype=F32)
    p = np.asarray(pts, dtype=F32)[:, None]                     # (B,1,Nq,3)
    R = Tcl[..., :9].reshape(Tcl.shape[:-1] + (3, 3))[:, :, None]   # (B,T,1,3,3)
    t = Tcl[..., 9:][:, :, None]                                # (B,T,1,3)
    out = []
    for i in range(3):
        acc = p[..., 0] * R[..., i, 0]
        acc = _fma(p[..., 1], R[..., i, 1], acc)
        acc = _fma(p[..., 2], R[..., i, 2], acc)
        out.append(acc + t[..., i])
    return np.stack(out, -1).astype(F32)


def pinhole_project(camera, pc):
    """Camera.project + in_image (wrappers.py:502-522).  camera (B,T,6), pc (B,T,Nq,3)
    -> center_im (B,T,Nq,2) fp32, valid (B,T,Nq) bool."""
    cam = np.asarray(camera, dtype=F32)[:, :, None]             # (B,T,1,6)
    z = pc[..., 2]
    in_front = z > EPS_Z
    zc = np.maximum(z, EPS_Z)
    u = (pc[..., 0] / zc) * cam[..., 2] + cam[..., 4]
    v = (pc[..., 1] / zc) * cam[..., 3] + cam[..., 5]
    wm1 = cam[..., 0] - F32(1)
    hm1 = cam[..., 1] - F32(1)
    valid = in_front & (u >= 0) & (u <= wm1) & (v >= 0) & (v <= hm1)
    return np.stack([u, v], -1).astype(F32), valid


def denormalize(ref, scale):
    """TransformerDecoder.denormalize (transformer_parq.py:198-209): p*(hi-lo)+lo per axis."""
    s = [float(x) for x in scale]
    return torch.stack([ref[..., 0] * (s[1] - s[0]) + s[0],
                        ref[..., 1] * (s[3] - s[2]) + s[2],
                        ref[..., 2] * (s[5] - s[4]) + s[4]], dim=-1)


def normalize(c, scale):
    """TransformerDecoder.normalize (transformer_parq.py:185-196): (c-lo)/(hi-lo) per axis."""
    s = [float(x) for x in scale]
    return torch.stack([(c[..., 0] - s[0]) / (s[1] - s[0]),
                        (c[..., 1] - s[2]) / (s[3] - s[2]),
                        (c[..., 2] - s[4]) / (s[5] - s[4])], dim=-1)


def inverse_sigmoid(x, eps=1e-3):
    """transformer_parq.py:38-42."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


# --------------------------------------------------------------------------- #
# A.3/A.4  projection + bilinear multi-view gather (transformer_parq.py:129-161)
# --------------------------------------------------------------------------- #
def project_sample(tokens, coord_pos, T_camera_local, camera, H, W):
    """tokens (B, T*H*W, C) fp32; coord_pos (B,Nq,3) metres in the local frame;
    T_camera_local (B,T,12); camera (B,T,6).
    Returns features (B,Nq,C), center_im (B,T,Nq,2), center_valid (B,T,Nq) bool."""
    B, Nq = coord_pos.shape[:2]
    Tv = T_camera_local.shape[1]
    C = tokens.shape[-1]
    pc = transform_points(np.asarray(T_camera_local), coord_pos.detach().numpy())
    center_im_np, valid_np = pinhole_project(np.asarray(camera), pc)
    center_im = torch.from_numpy(center_im_np)
    center_valid = torch.from_numpy(valid_np)
    w = np.float32(W)
    h = np.float32(H)
    # transformer_parq.py:148-150 (w,h are numpy float32 scalars there)
    gx = 2 * center_im[..., 0] / float(w - 1) - 1
    gy = 2 * center_im[..., 1] / float(h - 1) - 1
    grid = torch.stack([gx, gy], dim=-1).view(B * Tv, 1, Nq, 2)
    memory_hw = tokens.view(B * Tv, H, W, C).permute(0, 3, 1, 2)          # :302-303, channels-last strides
    feat = F.grid_sample(memory_hw, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    feat = feat.view(B, Tv, C, Nq).permute(0, 1, 3, 2).contiguous()
    feat = feat.sum(dim=1)                                               # ALL views, valid or not (:156)
    cnt = center_valid.sum(dim=1)
    cnt[cnt == 0] = 1
    feat = feat / cnt.unsqueeze(-1)
    return feat, center_im, center_valid


# --------------------------------------------------------------------------- #
# A.5  reference-point positional encoding (transformer_parq.py:45-64, 176-180)
# --------------------------------------------------------------------------- #
def pos_dim_t(num_pos_feats=128, temperature=10000):
    d = torch.arange(num_pos_feats, dtype=torch.float32)
    return temperature ** (2 * (d // 2) / num_pos_feats)


def pos2posemb3d(pos, num_pos_feats=128, temperature=10000):
    pos = pos * (2 * math.pi)
    dim_t = pos_dim_t(num_pos_feats, temperature)
    embs = []
    for axis in (1, 0, 2):                                      # concat order (y, x, z), :63
        a = pos[..., axis, None] / dim_t
        embs.append(torch.stack((a[..., 0::2].sin(), a[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(embs, dim=-1)


# --------------------------------------------------------------------------- #
# A.6  decoder layer, heads, box update
# --------------------------------------------------------------------------- #
def _mha(q_in, k_in, v_in, w_in, b_in, w_out, b_out, heads, kv=None):
    """nn.MultiheadAttention forward, need_weights path without the weights
    (torch/nn/functional.py multi_head_attention_forward): inputs (L,B,E)/(S,B,E).
    ``kv`` optionally carries cached (k, v) projections (iteration-invariant for
    the cross attention; the reference recomputes them every iteration)."""
    L, B, E = q_in.shape
    dh = E // heads
    q = F.linear(q_in, w_in[:E], b_in[:E])
    if kv is None:
        k = F.linear(k_in, w_in[E:2 * E], b_in[E:2 * E])
        v = F.linear(v_in, w_in[2 * E:], b_in[2 * E:])
    else:
        k, v = kv
    S = k.shape[0]
    q = q.view(L, B * heads, dh).transpose(0, 1)
    k = k.view(S, B * heads, dh).transpose(0, 1)
    v = v.view(S, B * heads, dh).transpose(0, 1)
    q = q * math.sqrt(1.0 / float(dh))
    attn = torch.softmax(torch.bmm(q, k.transpose(-2, -1)), dim=-1)
    out = torch.bmm(attn, v)
    out = out.transpose(0, 1).contiguous().view(L * B, E)
    return F.linear(out, w_out, b_out).view(L, B, E)


def _head3(x_cn, sd, name):
    """3-layer GenericMLP head (generic_mlp.py:94-110 with use_conv, norm 'ln' ->
    GroupNorm(1,C), parq_decoder.py:90-98): x_cn (B,C,Nq)."""
    p = "mlp_heads." + name + ".layers."
    y = F.conv1d(x_cn, sd[p + "0.weight"])
    y = F.relu(F.group_norm(y, 1, sd[p + "1.weight"], sd[p + "1.bias"], 1e-5))
    y = F.conv1d(y, sd[p + "4.weight"])
    y = F.relu(F.group_norm(y, 1, sd[p + "5.weight"], sd[p + "5.bias"], 1e-5))
    return F.conv1d(y, sd[p + "8.weight"], sd[p + "8.bias"])


def box_heads(x, ref, sd, scale):
    """bbox3d_prediction (transformer_parq.py:211-281) + BoxProcessor
    (utils/parq_utils.py:90-105).  x (B,Nq,C) decoder output, ref (B,Nq,3) normalised."""
    xc = x.permute(0, 2, 1).contiguous()
    logits = F.conv1d(xc, sd["mlp_heads.sem_cls_head.layers.0.weight"], sd["mlp_heads.sem_cls_head.layers.0.bias"]).transpose(1, 2)
    center_off = _head3(xc, sd, "center_head").transpose(1, 2)
    coord_pos = denormalize(ref, scale)
    center = denormalize((center_off + inverse_sigmoid(ref)).sigmoid(), scale)
    size_s = F.conv1d(xc, sd["mlp_heads.size_head.layers.0.weight"], sd["mlp_heads.size_head.layers.0.bias"]).transpose(1, 2)
    ortho6d = _head3(xc, sd, "rotation_head").transpose(1, 2)
    prob = torch.softmax(logits, dim=-1)
    mean_size = torch.from_numpy(MEAN_SIZE)[prob.argmax(-1)]
    size = torch.exp(size_s) * mean_size.float()
    return {"pred_logits": logits, "center_unnormalized": center * 1, "size_unnormalized": size,
            "ortho6d": ortho6d, "sem_cls_prob": prob, "coord_pos": coord_pos}


def decoder_iteration(tokens, memory, ref, T_camera_local, camera, H, W, sd, heads=4, scale=None, kv=None, layer=0):
    """One pass of the hot loop body (transformer_parq.py:311-332) for normalised
    reference points ``ref`` (B,Nq,3).  ``layer``: index of the decoder layer this iteration uses (0 when the
    weights are shared, the iteration number otherwise, :311-314).  Returns (out_dict, next_ref, aux)."""
    L = "parq_module.decoder.layers.%d." % layer
    P = "parq_module.decoder.position_encoder."
    pe = F.linear(F.relu(F.linear(pos2posemb3d(ref), sd[P + "0.weight"], sd[P + "0.bias"])), sd[P + "2.weight"], sd[P + "2.bias"])
    pe = pe.permute(1, 0, 2)
    feat, center_im, center_valid = project_sample(tokens, denormalize(ref, scale), T_camera_local, camera, H, W)
    x = feat.permute(1, 0, 2)
    qk = x + pe
    x2 = _mha(qk, qk, x, sd[L + "self_attn.in_proj_weight"], sd[L + "self_attn.in_proj_bias"],
              sd[L + "self_attn.out_proj.weight"], sd[L + "self_attn.out_proj.bias"], heads)
    sa_out = x2
    x = F.layer_norm(x + x2, (x.shape[-1],), sd[L + "norm1.weight"], sd[L + "norm1.bias"], 1e-5)
    x1 = x
    x2 = _mha(x + pe, memory, memory, sd[L + "multihead_attn.in_proj_weight"], sd[L + "multihead_attn.in_proj_bias"],
              sd[L + "multihead_attn.out_proj.weight"], sd[L + "multihead_attn.out_proj.bias"], heads, kv=kv)
    ca_out = x2
    x = F.layer_norm(x + x2, (x.shape[-1],), sd[L + "norm2.weight"], sd[L + "norm2.bias"], 1e-5)
    x2_ln = x
    x2 = F.linear(F.relu(F.linear(x, sd[L + "linear1.weight"], sd[L + "linear1.bias"])), sd[L + "linear2.weight"], sd[L + "linear2.bias"])
    x = F.layer_norm(x + x2, (x.shape[-1],), sd[L + "norm3.weight"], sd[L + "norm3.bias"], 1e-5)
    x = x.permute(1, 0, 2)
    out = box_heads(x, ref, sd, scale)
    nxt = normalize(out["center_unnormalized"], scale)
    aux = {"features": feat, "center_im": center_im, "center_valid": center_valid, "decoder_out": x,
           "pe": pe.permute(1, 0, 2), "self_attn_out": sa_out.permute(1, 0, 2), "x1": x1.permute(1, 0, 2),
           "cross_attn_out": ca_out.permute(1, 0, 2), "x2": x2_ln.permute(1, 0, 2)}
    return out, nxt, aux


def decoder_forward(tokens, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, sd,
                    iters=8, heads=4, scale=(-3, 3, -2, 0.5, 0.25, 5.25), forced_refs=None,
                    hoist_kv=True, return_aux=False):
    """PARQDecoder.forward (parq_decoder.py:134-163) -> TransformerDecoder.forward
    (transformer_parq.py:283-337).  All pose/camera arguments are raw tensors
    ((B,T,12)/(B,1,12)/(B,T,6)).  ``forced_refs`` (iters,B,Nq,3): teacher-forced
    normalised reference points per iteration.  ``hoist_kv=False`` recomputes the
    cross-attention K/V projection every iteration exactly as the reference does
    (used when this oracle is timed as the CPU baseline)."""
    with torch.no_grad():
        tokens = tokens.float()
        B = tokens.shape[0]
        cam = camera.float()
        W = int(cam[0, 0, 0].item())
        H = int(cam[0, 0, 1].item())
        Tcl = camera_from_local(T_camera_pseudoCam.numpy(), T_world_pseudoCam.numpy(), T_world_local.numpy())
        memory = tokens.permute(1, 0, 2)
        E = tokens.shape[-1]
        # SHARE_WEIGHTS False (transformer_parq.py:168-171): one layer per iteration, nothing to hoist
        shared = "parq_module.decoder.layers.1.norm1.weight" not in sd
        L = "parq_module.decoder.layers.0.multihead_attn."
        kv = None
        if hoist_kv and shared:
            kv = (F.linear(memory, sd[L + "in_proj_weight"][E:2 * E], sd[L + "in_proj_bias"][E:2 * E]),
                  F.linear(memory, sd[L + "in_proj_weight"][2 * E:], sd[L + "in_proj_bias"][2 * E:]))
        ref = sd["refpoint.weight"].unsqueeze(0).repeat(B, 1, 1).sigmoid()
        outs, auxs = [], []
        for it in range(iters):
            if forced_refs is not None:
                ref = forced_refs[it]
            out, ref, aux = decoder_iteration(tokens, memory, ref, Tcl, cam.numpy(), H, W, sd, heads, scale, kv, layer=0 if shared else it)
            outs.append(out)
            auxs.append(aux)
        return (outs, auxs) if return_aux else outs


def refs_from_outputs(outs, sd, scale=(-3, 3, -2, 0.5, 0.25, 5.25)):
    """Normalised reference points consumed by every iteration of a run whose
    per-iteration outputs are ``outs`` (teacher forcing):
    ref_0 = sigmoid(refpoint), ref_i = normalize(center_{i-1})."""
    B = outs[0]["center_unnormalized"].shape[0]
    refs = [sd["refpoint.weight"].unsqueeze(0).repeat(B, 1, 1).sigmoid()]
    for o in outs[:-1]:
        refs.append(normalize(o["center_unnormalized"], scale))
    return torch.stack(refs)


# --------------------------------------------------------------------------- #
# a13  ortho6d -> rotation (utils/ortho6d_transforms.py:23-66)
# --------------------------------------------------------------------------- #
def rotation_from_ortho6d(o):
    """(N,6) -> (N,3,3), columns [x y z]; norms clamped at 1e-8."""
    def nrm(v):
        return v / torch.sqrt(v.pow(2).sum(1)).clamp(min=1e-8).unsqueeze(1)
    x = nrm(o[:, 0:3])
    z = nrm(torch.cross(x, o[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


# --------------------------------------------------------------------------- #
# f-2  parse_pred + NMS (model/parq_decoder.py:372-424, utils/nms.py:20-70,141-179)
# --------------------------------------------------------------------------- #
def box_corners_local(center, size, ortho6d):
    """8 corners of every predicted box in the snippet-local frame, fp32 torch, in the reference's
    operation order: R = compute_rotation_matrix_from_ortho6d (parq_decoder.py:383-386), object-frame
    corners from +-size/2 in the order of Obb3D.bb3corners_object (utils/wrappers.py:355-392), then
    Pose.transform = p @ R^T + t (utils/nms.py:24, wrappers.py:260-267).  (B,K,3),(B,K,3),(B,K,6) -> (B,K,8,3)."""
    B, K = center.shape[:2]
    R = rotation_from_ortho6d(ortho6d.contiguous().view(-1, 6)).view(B, K, 3, 3)
    lo, hi = -size / 2, size / 2
    xs = torch.stack([lo[..., 0], hi[..., 0], hi[..., 0], lo[..., 0], lo[..., 0], hi[..., 0], hi[..., 0], lo[..., 0]], -1)
    ys = torch.stack([lo[..., 1], lo[..., 1], hi[..., 1], hi[..., 1], lo[..., 1], lo[..., 1], hi[..., 1], hi[..., 1]], -1)
    zs = torch.stack([lo[..., 2], lo[..., 2], lo[..., 2], lo[..., 2], hi[..., 2], hi[..., 2], hi[..., 2], hi[..., 2]], -1)
    c = torch.stack([xs, ys, zs], -1)                                   # (B,K,8,3)
    return c @ R.transpose(-1, -2) + center.unsqueeze(-2)


def nms_3d_faster(boxes, overlap_threshold, same_class=False):
    """utils/nms.py:141-179 (class-agnostic greedy NMS on AABBs, float64).  boxes (n, >=7): x1,y1,z1,x2,y2,z2,score.
    same_class: nms_3d_faster_samecls (utils/nms.py:182-224) -- column 7 is the class, the overlap of two boxes of
    different classes is multiplied by 0 before the threshold test."""
    x1, y1, z1, x2, y2, z2, score = (boxes[:, i] for i in range(7))
    area = (x2 - x1) * (y2 - y1) * (z2 - z1)
    order = np.argsort(score)
    pick = []
    while order.size != 0:
        last = order.size
        i = order[-1]
        pick.append(i)
        rest = order[: last - 1]
        l = np.maximum(0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        w = np.maximum(0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        h = np.maximum(0, np.minimum(z2[i], z2[rest]) - np.maximum(z1[i], z1[rest]))
        inter = l * w * h
        o = inter / (area[i] + area[rest] - inter)
        if same_class:
            o = o * (boxes[i, 7] == boxes[rest, 7])
        order = np.delete(order, np.concatenate(([last - 1], np.where(o > overlap_threshold)[0])))
    return pick


def parse_pred(last, track_scale=(-1.5, 1.5, -2, 1, 0, 2), num_semcls=9, overlap_threshold=None, for_vis=False):
    """PARQDecoder.parse_pred (parq_decoder.py:372-424) with ENABLE_NMS True (config/eval.yaml):
    scores, labels = max over ALL classes; AABB of the rotated corners; NMS over the non-background boxes;
    FOR_VIS False: class-agnostic NMS at IoU 0.1, pred_mask = nms & (x in (ts0,ts1)) & (z in (ts4,ts5));
    FOR_VIS True (:407-421): same-class NMS at IoU 0.2 and no track-scale filter.
    Returns dict(pred_mask (B,K) bool, scores (B,K), labels (B,K), aabb (B,K,6) float64)."""
    center, size, o6, prob = (last[k].detach().float().cpu() for k in
                              ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob"))
    scores, labels = torch.max(prob, -1)
    corners = box_corners_local(center, size, o6).numpy()
    B, K = scores.shape
    aabb = np.concatenate([corners.min(axis=2), corners.max(axis=2)], -1).astype(np.float64)     # (B,K,6)
    mask = np.zeros((B, K), dtype=bool)
    if overlap_threshold is None:
        overlap_threshold = 0.2 if for_vis else 0.1
    for b in range(B):
        fg = np.where(labels[b].numpy() != num_semcls)[0]
        boxes = np.concatenate([aabb[b, fg], scores[b, fg].numpy().astype(np.float64)[:, None],
                                labels[b, fg].numpy().astype(np.float64)[:, None]], 1)
        pick = nms_3d_faster(boxes, overlap_threshold, same_class=for_vis)
        mask[b, fg[pick]] = True
    ts = track_scale
    valid = (center[..., 0] > ts[0]) & (center[..., 0] < ts[1]) & (center[..., 2] > ts[4]) & (center[..., 2] < ts[5])
    if for_vis:
        valid = torch.ones_like(valid)
    return {"pred_mask": torch.from_numpy(mask) & valid, "nms_mask": torch.from_numpy(mask), "scores": scores, "labels": labels,
            "aabb": torch.from_numpy(aabb)}


# --------------------------------------------------------------------------- #
# f-1  AddRayPE + tokeniser (model/ray_positional_encoding.py:61-139,
#      utils/encoding_utils.py:15-100, model/parq_lightning.py:72-85)
# --------------------------------------------------------------------------- #
def _pose_parts(P):
    P = torch.as_tensor(P, dtype=torch.float32)
    return P[..., :9].reshape(P.shape[:-1] + (3, 3)), P[..., 9:]


def _pose_inv(P):
    R, t = _pose_parts(P)
    Rt = R.transpose(-1, -2)
    return torch.cat([Rt.flatten(-2), -(Rt @ t.unsqueeze(-1)).squeeze(-1)], -1)


def _pose_mul(A, B):
    RA, tA = _pose_parts(A)
    RB, tB = _pose_parts(B)
    return torch.cat([(RA @ RB).flatten(-2), tA + (RA @ tB.unsqueeze(-1)).squeeze(-1)], -1)


def _pose_apply(P, pts):
    R, t = _pose_parts(P)
    return pts @ R.transpose(-1, -2) + t.unsqueeze(-2)


def ray_depth_planes(num_samples=64, min_depth=0.25, max_depth=5.25):
    """utils/encoding_utils.py:82-88: log-spaced depth planes, fp32 torch arithmetic."""
    ramp = torch.linspace(0, 1, num_samples)
    mn, mx = torch.tensor([min_depth])[0], torch.tensor([max_depth])[0]
    return torch.exp(torch.log(mn) + torch.log(mx / mn) * ramp)


def ray_features(camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, H, W,
                 ray_points_scale=(-3, 3, -2, 0.5, 0.25, 5.25), num_samples=64, min_depth=0.25, max_depth=5.25):
    """The (B*T, H, W, 3*num_samples) input of AddRayPE's encoder (ray_positional_encoding.py:76-128):
    pixel grid (x = 0..W-1, y = 0..H-1, encoding_utils.py:15-20) -> unproject (wrappers.py:523-549) -> points at the
    depth planes -> pseudo-camera frame -> snippet-local frame -> normalise by ray_points_scale -> inverse sigmoid."""
    B, T = T_camera_pseudoCam.shape[:2]
    cam = camera.reshape(B * T, 6).float()
    xs = torch.linspace(0.0, W, W + 1)[:-1]
    ys = torch.linspace(0.0, H, H + 1)[:-1]
    xx, yy = torch.meshgrid(xs, ys, indexing="xy")
    uv = torch.stack([xx, yy], -1).reshape(1, H * W, 2)
    rays = (uv - cam[:, None, 4:6]) / cam[:, None, 2:4]
    rays = torch.cat([rays, torch.ones(B * T, H * W, 1)], -1)
    pts = rays.unsqueeze(-2) * ray_depth_planes(num_samples, min_depth, max_depth).view(1, 1, num_samples, 1)
    pts = pts.view(B * T, -1, 3)
    pts = _pose_apply(_pose_inv(T_camera_pseudoCam.reshape(B * T, 12)), pts)
    T_local_pseudoCam = _pose_mul(_pose_inv(T_world_local), T_world_pseudoCam).reshape(B * T, 12)
    pts = _pose_apply(T_local_pseudoCam, pts).view(B * T, H, W, num_samples, 3)
    s = [float(v) for v in ray_points_scale]
    pts = torch.stack([(pts[..., 0] - s[0]) / (s[1] - s[0]), (pts[..., 1] - s[2]) / (s[3] - s[2]),
                       (pts[..., 2] - s[4]) / (s[5] - s[4])], -1)
    return inverse_sigmoid(pts).reshape(B * T, H, W, 3 * num_samples)


def add_ray_pe(images_feat, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, sd, **kw):
    """AddRayPE.forward (ray_positional_encoding.py:61-139) + the tokeniser of PARQ.forward (parq_lightning.py:75-85).
    images_feat (B,T,C,H,W); sd holds encoder.{0,2}.{weight,bias}.  Returns (encoding (B,T,C,H,W), tokens (B,T*H*W,C))."""
    with torch.no_grad():
        B, T, C, H, W = images_feat.shape
        f = ray_features(camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, H, W, **kw)
        enc = F.linear(F.relu(F.linear(f, sd["encoder.0.weight"], sd["encoder.0.bias"])), sd["encoder.2.weight"], sd["encoder.2.bias"])
        enc = enc.view(B, T, H, W, C).permute(0, 1, 4, 2, 3)
        tokens = (images_feat + enc).permute(0, 1, 3, 4, 2).reshape(B, T * H * W, C)
        return enc.contiguous(), tokens.contiguous()


# --------------------------------------------------------------------------- #
# f-3  FPN upsample + concat (model/resnet_fpn.py:73-90)
# --------------------------------------------------------------------------- #
def fpn_concat(features, layer=0):
    """ResnetFPN.forward :73-80: every pyramid level resized to the size of level ``layer`` with
    F.interpolate(mode="bilinear") (align_corners=False), concatenated along channels."""
    size = features[str(layer)].shape[-2:]
    return torch.cat([F.interpolate(features[str(l)], size, mode="bilinear") for l in range(4)], dim=1)
