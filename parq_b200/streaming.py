"""Streaming window cache ("next" row f-4, BASELINE.json configs[4]: one clip, an 8-view window sliding by one view).

The reference slides the window by re-running everything: every window is a fresh snippet whose local frame is the
pseudo-camera of its middle view (/root/reference/datasets/transforms.py:191-208) and whose ray positional encoding --
and therefore every token -- is recomputed in that frame (model/ray_positional_encoding.py:88-117).  Seven of the eight
views of consecutive windows are the same images.

``StreamingWindow`` keeps, per view slot of a ring, the view's tokens and their K / V^T projections in the decoder's
workspace.  ``push`` replaces the oldest view: one view's K / V^T projection (1/T of the work) through
``parq_kv_project_views``; ``decode`` runs the recurrent iterations over the cached K / V^T (``PARQ_FLAG_SKIP_KV``),
replayed as one CUDA graph.  Two ways to use it:

* **exact** (``decode(T_world_local)``): whatever frame the caller's tokens were encoded in, the result is what
  ``PARQDecoder.forward`` returns for the ring's views and that ``T_world_local`` (the views sit in ring order, not
  temporal order; the decoder is invariant to the view order up to fp32 summation order).  This is the cache alone, no
  change of semantics: the caller still re-encodes tokens when its local frame changes.
* **anchor frame** (``decode_anchored(T_world_local_now)``): tokens are encoded ONCE per view in a fixed anchor frame A
  (``T_world_anchor``), so no view is ever re-encoded; the learned reference points, which live in the window's local
  frame L, are re-expressed in A (p_A = T_A<-L p_L), the decoder runs in A, and the boxes are mapped back to L.  The
  projection into the images is unchanged (T_cam<-A T_A<-L = T_cam<-L), hence so is the sampled query content; what
  changes is the frame in which the reference-point MLP and the box update see coordinates.  That is a change of
  semantics against the reference (it is what makes the window incremental); tests/test_gpu_streaming.py reports the
  deviation from the per-window reference next to the exact mode's parity.
"""
import ctypes as C

import torch

from . import _lib
from .decoder import OUTPUT_KEYS, _ptr, _stream
from .wrappers import raw


def _pose_inv(P):
    R = P[..., :9].reshape(P.shape[:-1] + (3, 3))
    t = P[..., 9:]
    Rt = R.transpose(-1, -2)
    return torch.cat([Rt.reshape(P.shape[:-1] + (9,)), -(Rt @ t.unsqueeze(-1)).squeeze(-1)], -1)


def _pose_mul(A, B):
    RA, RB = A[..., :9].reshape(A.shape[:-1] + (3, 3)), B[..., :9].reshape(B.shape[:-1] + (3, 3))
    return torch.cat([(RA @ RB).reshape(A.shape[:-1] + (9,)), A[..., 9:] + (RA @ B[..., 9:].unsqueeze(-1)).squeeze(-1)], -1)


class StreamingWindow:
    def __init__(self, engine, T, H, W, B=1):
        if (H * W) % 32 != 0:
            raise ValueError("per-view K / V^T updates need H*W % 32 == 0")
        if engine.n_layers != 1:
            raise NotImplementedError("the K / V^T cache needs shared decoder layers (SHARE_WEIGHTS=True)")
        self.eng, self.B, self.T, self.H, self.W = engine, B, T, H, W
        dev, Cc = engine.device, engine.C
        self.tokens = torch.zeros(B, T * H * W, Cc, dtype=torch.bfloat16, device=dev)      # ring of view slots
        self.camera = torch.zeros(B, T, 6, dtype=torch.float32, device=dev)
        self.T_cp = torch.zeros(B, T, 12, dtype=torch.float32, device=dev)
        self.T_wp = torch.zeros(B, T, 12, dtype=torch.float32, device=dev)
        self.count = 0
        self._shape = engine._shape(B, T, H, W)
        self._ws = engine._workspace(self._shape, (B, T, H, W))

    @property
    def full(self):
        return self.count >= self.T

    def slot_order(self):
        """View slots from the oldest to the newest view."""
        n = min(self.count, self.T)
        first = self.count % self.T if self.count >= self.T else 0
        return [(first + i) % self.T for i in range(n)]

    def push(self, view_tokens, camera, T_camera_pseudoCam, T_world_pseudoCam):
        """view_tokens (B, H*W, C) bf16/fp32 of the NEW view; camera (B, 6); poses (B, 12).  Replaces the oldest view."""
        eng, hw = self.eng, self.H * self.W
        slot = self.count % self.T
        if self._ws is not eng._workspace(self._shape, (self.B, self.T, self.H, self.W)):
            raise RuntimeError("the engine's workspace was re-created for another shape: the K / V^T cache is gone")
        self.tokens[:, slot * hw:(slot + 1) * hw].copy_(view_tokens.reshape(self.B, hw, eng.C))
        self.camera[:, slot].copy_(raw(camera).reshape(self.B, 6))
        self.T_cp[:, slot].copy_(raw(T_camera_pseudoCam).reshape(self.B, 12))
        self.T_wp[:, slot].copy_(raw(T_world_pseudoCam).reshape(self.B, 12))
        view = self.tokens[:, slot * hw:(slot + 1) * hw]
        view = view if self.B == 1 else view.contiguous()
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.parq_kv_project_views(C.byref(self._shape), _ptr(view), slot, 1, _ptr(eng.packed), _ptr(self._ws),
                                                     self._ws.numel(), eng.flags, _stream()), "parq_kv_project_views")
        self.count += 1
        return slot

    def decode(self, T_world_local, ref0=None, forced_refs=None, graph=True, debug=False):
        """The decoder over the cached window; returns the dict of stacked per-iteration tensors of DecoderEngine.forward."""
        if not self.full:
            raise RuntimeError("the window holds %d of %d views" % (self.count, self.T))
        return self.eng.forward(self.tokens, self.camera, self.T_cp, self.T_wp, raw(T_world_local).reshape(self.B, 1, 12), self.H, self.W,
                                ref0=ref0, forced_refs=forced_refs, skip_kv=True, graph=graph, debug=debug)

    def decode_anchored(self, T_world_anchor, T_world_local, graph=True):
        """Decode in the fixed anchor frame A with the learned reference points re-expressed from the window's local frame
        L, then map centres and rotations back to L.  Returns (outs_in_L, outs_in_A)."""
        eng = self.eng
        Twa, Twl = raw(T_world_anchor).reshape(self.B, 12).float(), raw(T_world_local).reshape(self.B, 12).float()
        T_al = _pose_mul(_pose_inv(Twa), Twl)                       # A <- L
        lo = torch.tensor(eng.scale[0::2], device=eng.device)
        span = torch.tensor(eng.scale[1::2], device=eng.device) - lo
        p_l = eng._ref0(1)[0] * span + lo                           # (Nq, 3) metres in L
        R_al = T_al[:, :9].reshape(self.B, 3, 3)
        p_a = p_l.unsqueeze(0) @ R_al.transpose(1, 2) + T_al[:, None, 9:]
        ref0 = ((p_a - lo) / span).contiguous()
        outs_a = self.decode(Twa.reshape(self.B, 1, 12), ref0=ref0, graph=graph)
        T_la = _pose_inv(T_al)
        R_la = T_la[:, :9].reshape(self.B, 1, 1, 3, 3)
        outs_l = dict(outs_a)
        for k in ("center_unnormalized", "coord_pos"):
            outs_l[k] = (outs_a[k].unsqueeze(-2) @ R_la.transpose(-1, -2)).squeeze(-2) + T_la[:, None, 9:]
        o6 = outs_a["ortho6d"]                                      # two column vectors of the rotation: rotate both
        a, b = o6[..., :3], o6[..., 3:]
        rot = lambda v: (v.unsqueeze(-2) @ R_la.transpose(-1, -2)).squeeze(-2)
        outs_l["ortho6d"] = torch.cat([rot(a), rot(b)], -1)
        return outs_l, outs_a


def per_iteration(outs, iters):
    """Stacked tensors -> the reference's list of per-iteration dicts."""
    return [{k: outs[k][i] for k, _ in OUTPUT_KEYS} for i in range(iters)]
