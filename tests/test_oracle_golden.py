"""The oracle (oracle/parq_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import OUT_KEYS, bit_equal, load_golden, regenerate_case, relerr
from oracle import parq_oracle as O

FULL = ["small", "ragged_wild", "white_noise"]
PROJ = ["proj_c1", "proj_c4_views", "proj_wild"]


@pytest.mark.parametrize("name", FULL + PROJ)
def test_pose_chain_bit_exact(name):
    gold = load_golden(name)
    c = regenerate_case(gold)
    Tcl = O.camera_from_local(c["T_cp"].numpy(), c["T_wp"].numpy(), c["T_wl"].numpy())
    assert bit_equal(Tcl, gold["T_camera_local"])


@pytest.mark.parametrize("name", PROJ)
def test_projection_bit_exact(name):
    gold = load_golden(name)
    c = regenerate_case(gold)
    feat, cim, val = O.project_sample(c["tokens"], c["points"], gold["T_camera_local"], c["camera"].numpy(), c["H"], c["W"])
    assert bit_equal(cim, gold["center_im"])
    assert np.array_equal(val.numpy(), gold["center_valid"])
    assert relerr(feat[..., ::16], gold["features"]) <= 1e-6


@pytest.mark.parametrize("name", FULL)
def test_decoder_teacher_forced(name):
    gold = load_golden(name)
    c = regenerate_case(gold)
    sd = c["sd"]
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(gold["coord_pos"].shape[0])]
    refs = O.refs_from_outputs(gold_outs, sd)
    outs, auxs = O.decoder_forward(c["tokens"], c["camera"], c["T_cp"], c["T_wp"], c["T_wl"], sd, forced_refs=refs, return_aux=True)
    for i, (o, a) in enumerate(zip(outs, auxs)):
        assert bit_equal(o["coord_pos"], gold["coord_pos"][i]), "teacher-forced reference points differ at iteration %d" % i
        assert bit_equal(a["center_im"], gold["center_im"][i])
        assert np.array_equal(a["center_valid"].numpy(), gold["center_valid"][i])
        assert relerr(a["features"][..., ::16], gold["features"][i]) <= 1e-6
        for k in OUT_KEYS:
            assert relerr(o[k], gold[k][i]) <= 2e-5, (k, i)


def test_decoder_free_running_first_iterations():
    # the recurrence amplifies rounding noise: gate the first two iterations only
    gold = load_golden("small")
    c = regenerate_case(gold)
    outs = O.decoder_forward(c["tokens"], c["camera"], c["T_cp"], c["T_wp"], c["T_wl"], c["sd"], iters=2)
    for i in range(2):
        for k in OUT_KEYS:
            assert relerr(outs[i][k], gold[k][i]) <= 1e-4, (k, i)


def test_unhoisted_kv_matches_hoisted():
    gold = load_golden("ragged_wild")
    c = regenerate_case(gold)
    a = O.decoder_forward(c["tokens"], c["camera"], c["T_cp"], c["T_wp"], c["T_wl"], c["sd"], iters=1, hoist_kv=True)
    b = O.decoder_forward(c["tokens"], c["camera"], c["T_cp"], c["T_wp"], c["T_wl"], c["sd"], iters=1, hoist_kv=False)
    for k in OUT_KEYS:
        assert relerr(a[0][k], b[0][k]) <= 1e-5


def test_mean_size_table_and_rotation():
    assert O.MEAN_SIZE.shape == (10, 3) and np.all(O.MEAN_SIZE[8:] == 1.0)
    g = torch.Generator().manual_seed(0)
    R = O.rotation_from_ortho6d(torch.randn(64, 6, generator=g))
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)


@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_oracle_parse_pred_matches_reference_detection_sets(name):
    # post-NMS detection set and the NMS decision alone, from the reference's own parse_pred / nms (make_golden.py)
    gold = load_golden(name)
    last = {k: torch.from_numpy(gold[k][-1]) for k in ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")}
    r = O.parse_pred(last)
    assert np.array_equal(r["pred_mask"].numpy(), gold["pred_mask"])
    assert np.array_equal(r["nms_mask"].numpy(), gold["nms_mask"])
    assert gold["nms_mask"].sum() >= 5                       # the fixture exercises real suppression
    obb = gold["obbs_pred"]
    assert np.array_equal(obb[..., 18], r["labels"].numpy().astype(np.float32))
    aabb = O.box_corners_local(last["center_unnormalized"], last["size_unnormalized"], last["ortho6d"])
    assert np.allclose(np.concatenate([aabb.min(2)[0].numpy(), aabb.max(2)[0].numpy()], -1), r["aabb"].numpy(), atol=0)


@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_oracle_parse_pred_for_vis_branch(name):
    # FOR_VIS=True (parq_decoder.py:407-421): same-class NMS (utils/nms.py:182-224) at IoU 0.2, no track-scale filter
    gold = load_golden(name)
    last = {k: torch.from_numpy(gold[k][-1]) for k in ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")}
    r = O.parse_pred(last, for_vis=True)
    assert np.array_equal(r["pred_mask"].numpy(), gold["pred_mask_vis"])
    assert gold["pred_mask_vis"].sum() > gold["nms_mask"].sum()      # fewer suppressions than the class-agnostic NMS


@pytest.mark.parametrize("name", ["raype_small", "raype_c1_view"])
def test_oracle_add_ray_pe_matches_reference(name):
    # f-1: AddRayPE encoding of the unmodified reference module (make_golden.py) against the oracle restatement
    from parq_b200 import inputs as I
    gold = load_golden(name)
    B, T, H, W, seed = [int(x) for x in gold["shape"]]
    sd = I.make_raype_weights(seed)
    feat = I.make_features(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    assert I.tensor_checksum(feat, cam._data, Tcp._data, Twp._data, Twl._data, *[sd[k] for k in sorted(sd)]) == str(gold["inputs_sum"])
    enc, tokens = O.add_ray_pe(feat, cam._data, Tcp._data, Twp._data, Twl._data, sd)
    assert relerr(enc[:, :, ::16], gold["encoding"]) <= 2e-6
    assert torch.equal(tokens.view(B, T, H, W, 1024), (feat + enc).permute(0, 1, 3, 4, 2))


@pytest.mark.parametrize("name", ["fpn_small", "fpn_odd"])
def test_oracle_fpn_concat_matches_reference(name):
    # f-3: `all_features` / `camera_feature` of the reference's own ResnetFPN.forward on a stubbed backbone (make_golden.py)
    from parq_b200 import inputs as I
    from parq_b200.fpn import camera_feature
    gold = load_golden(name)
    B, T, H, W, seed = [int(x) for x in gold["shape"]]
    pyr = I.make_pyramid(B * T, H, W, seed=seed)
    assert I.tensor_checksum(*[pyr[str(l)] for l in range(4)]) == str(gold["inputs_sum"])
    af = O.fpn_concat(pyr).view(B, T, 1024, H, W)
    assert bit_equal(af[:, :, ::16], gold["all_features"])
    cam = I.make_geometry(B, T, H * 4, W * 4, seed=seed)[0]
    assert bit_equal(camera_feature(cam)._data, gold["camera_feature"])


def test_oracle_unshared_decoder_layers():
    # SHARE_WEIGHTS False (transformer_parq.py:168-171, 311-314): fixture from the unmodified reference with 3 distinct layers
    from parq_b200 import inputs as I
    gold = load_golden("unshared")
    B, T, H, W, Nq, seed, layers = [int(x) for x in gold["shape"]]
    sd = I.make_weights(seed, Nq, n_layers=layers)
    assert I.tensor_checksum(*[sd[k] for k in sorted(sd)]) == str(gold["weights_sum"])
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(layers)]
    outs = O.decoder_forward(tokens, cam._data, Tcp._data, Twp._data, Twl._data, sd, iters=layers, forced_refs=O.refs_from_outputs(gold_outs, sd))
    for i in range(layers):
        for k in OUT_KEYS:
            assert relerr(outs[i][k], gold[k][i]) <= 2e-5, (k, i)
