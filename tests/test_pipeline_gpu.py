"""GPU parity of the whole inference step, FPM_Pipeline.test_step (pipeline.py:306-355): network -> masks / IoU ->
matching -> lifting -> PointDSC -> pose rows and CSV lines, on a synthetic batch with the CollateWrapper schema.

The network itself is checked against the oracle in test_backbone_gpu.py (1e-3).  Here the oracle's post-network
loop (oracle.post_network_step, pair by pair in the reference's order, CPU generator seeded as on_test_start does)
is run on the SAME feature maps / mask logits the GPU produced, so every discrete quantity must agree exactly:
masks, IoU, validity, the [500,4] correspondences (both multinomial draws included).

Poses are compared (1e-4) on the planted case, where the correspondences describe a real rigid motion.  With seeded
random network weights the matches are geometrically meaningless, registration has no consensus to find, and its
answer hinges on which zero-score points ATen's unstable argsort happens to put in the seed list (see
test_pointdsc_gpu.py): there only the rigid-transform property of the result is checked.
"""
import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from gpu_util import need_gpu
from oryon_b200 import synth, synth_backbone as sb
from oryon_b200.net import Oryon
from oryon_b200.pipeline import FPM_Pipeline, mask_postproc
from oryon_b200.utils.pointdsc.init import PointDSCSolver

pytestmark = pytest.mark.gpu

CFG = synth.POINTDSC_DEFAULT_CFG
ARGS = dict(device="cuda:0", corrs_device="cpu", use_seed=False, seed=1,
            dataset=dict(img_size=[224, 224], max_corrs=500),
            model=dict(image_encoder=dict(img_size=[192, 192])),
            test=dict(mask="predicted", src_sampling=5000, solver="pointdsc", n_corrs=500, dist_th=0.25, mask_threshold=0.5))


@pytest.fixture(scope="module")
def system():
    need_gpu()
    from oryon_b200 import _lib
    _lib.destroy_all()
    model = Oryon(None, "cuda:0", state_dict=sb.oryon_state_dict(11))
    psd = synth.pointdsc_state_dict(300)
    solver = PointDSCSolver(psd, in_dim=CFG["in_dim"], num_layers=CFG["num_layers"], num_channels=CFG["num_channels"],
                            num_iterations=CFG["num_iterations"], ratio=CFG["ratio"], sigma_d=CFG["sigma_d"], k=CFG["k"],
                            nms_radius=CFG["inlier_threshold"], device="cuda:0")
    return model, solver, psd


def _cuda_batch(batch):
    out = dict(batch)
    for key in ("anchor", "query"):
        v = dict(batch[key])
        v["rgb"], v["mask"] = v["rgb"].cuda(), v["mask"].cuda()
        v["orig_depth"] = [d.cuda() for d in v["orig_depth"]]
        out[key] = v
    out["prompt_tokens"] = batch["prompt_tokens"].cuda()
    return out


def _stacked_batch(batch):
    """The batch as the B200 collate hands it over: depth frames stacked in one pinned host tensor per view."""
    out = dict(batch)
    for key in ("anchor", "query"):
        v = dict(batch[key])
        v["orig_depth"] = torch.stack(v["orig_depth"]).pin_memory()
        out[key] = v
    return out


@pytest.mark.parametrize("batched_tail", [True, False])
@pytest.mark.parametrize("mask_mode", ["predicted", "oracle"])
def test_test_step_equals_reference_loop(system, mask_mode, batched_tail, tmp_path):
    model, solver, psd = system
    args = dict(ARGS, test=dict(ARGS["test"], mask=mask_mode, batched_tail=batched_tail))
    pipe = FPM_Pipeline(args, test_model=True, model=model, pointdsc_solver=solver)
    batch = synth.synthetic_batch(7, 3, empty_mask_pairs=(1,) if mask_mode == "oracle" else ())
    gb = _cuda_batch(batch)
    csv = tmp_path / "pred.csv"
    pipe.on_test_start(str(csv))                  # seeds torch CPU generator with 1 (pipeline.py:296-299)
    outputs = pipe.forward(gb)
    torch.manual_seed(1)
    rows = pipe.test_step(gb, 0)
    pipe.on_test_end()
    cpu_out = {k: v.cpu() for k, v in outputs.items()}
    torch.manual_seed(1)
    torch.set_num_threads(8)
    ref = oracle.post_network_step(cpu_out, batch, psd, CFG, mask_mode=mask_mode)
    assert [r["status"] for r in rows] == [r["status"] for r in ref]
    for i, (r, o) in enumerate(zip(rows, ref)):
        assert (np.isnan(r["iou_a"]) and np.isnan(o["iou_a"])) or r["iou_a"] == o["iou_a"], i
        assert (np.isnan(r["iou_q"]) and np.isnan(o["iou_q"])) or r["iou_q"] == o["iou_q"], i
        if o["corrs"] is None:
            assert r["corrs"] is None and torch.equal(r["pred_pose_rel"], torch.eye(4))
            continue
        assert torch.equal(r["corrs"].cpu(), o["corrs"]), f"pair {i}: correspondences differ"
        R = r["pred_pose_rel"][:3, :3].double()
        assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-5) and abs(torch.det(R).item() - 1) < 1e-5
        np.testing.assert_allclose(r["pred_pose"].numpy(), (r["pred_pose_rel"] @ batch["anchor"]["pose"][i].float()).numpy(), atol=1e-6)
    if mask_mode == "oracle":
        assert rows[1]["status"] == "invalid_mask"
    lines = csv.read_text().strip().split("\n")
    assert len(lines) == 3
    f = lines[0].split(",")
    assert f[0] == "a_7_0" and f[1] == "q_7_0" and len(f[2].split(" ")) == 12 and len(f) == 5   # id_a,id_q,<12 floats>,iou_a,iou_q


class _PlantedModel:
    """Stands in for the network: returns fixed (planted) outputs, so the post-network path has a well-posed answer."""

    def __init__(self, outputs):
        self.outputs = outputs

    def forward(self, xs):
        return self.outputs


@pytest.mark.parametrize("tail", ["batched_cuda_list", "batched_pinned_stack", "per_pair"])
def test_post_network_path_recovers_and_matches_planted_pose(system, tail):
    _, solver, psd = system
    outputs, batch = synth.planted_network_outputs(21, 4)
    args = dict(ARGS, test=dict(ARGS["test"], batched_tail=tail != "per_pair"))
    pipe = FPM_Pipeline(args, test_model=True, model=_PlantedModel({k: v.cuda() for k, v in outputs.items()}), pointdsc_solver=solver)
    gb = _stacked_batch(batch) if tail == "batched_pinned_stack" else _cuda_batch(batch)
    pipe.on_test_start()
    rows = pipe.test_step(gb, 0)
    torch.manual_seed(1)
    torch.set_num_threads(8)
    ref = oracle.post_network_step(outputs, batch, psd, CFG)
    K = synth.NOCS_INTRINSICS
    for i, (r, o) in enumerate(zip(rows, ref)):
        assert r["status"] == o["status"] == "ok"
        assert r["iou_a"] == o["iou_a"] and r["iou_q"] == o["iou_q"]
        assert torch.equal(r["corrs"].cpu(), o["corrs"]), f"pair {i}: correspondences differ"
        np.testing.assert_allclose(r["pred_pose_rel"].numpy(), o["pred_pose_rel"].numpy(), atol=1e-4)
        np.testing.assert_allclose(r["pred_pose"].numpy(), o["pred_pose"].numpy(), atol=2e-4)
        # the planted motion: a translation of (10 px * z / fx, 5 px * z / fy, 0) at z ~ 1 m
        t = r["pred_pose_rel"][:3, 3].numpy()
        np.testing.assert_allclose(t, [10.0 / K[0], 5.0 / K[4], 0.0], atol=5e-3)
        np.testing.assert_allclose(r["pred_pose_rel"][:3, :3].numpy(), np.eye(3), atol=2e-2)


def test_per_pair_interface_matches_batched(system):
    """is_detection_valid / get_featmap_corrs / get_pose (reference signatures) give what test_step gives."""
    model, solver, _ = system
    pipe = FPM_Pipeline(ARGS, test_model=True, model=model, pointdsc_solver=solver)
    gb = _cuda_batch(synth.synthetic_batch(8, 2))
    torch.manual_seed(1)
    rows = pipe.test_step(gb, 0)
    outputs = pipe.forward(gb)
    results = pipe.mask_results(gb, outputs)
    torch.manual_seed(1)
    for i in range(2):
        assert pipe.is_detection_valid(results, gb, i) == (rows[i]["status"] != "invalid_mask")
        corrs, pos_a, pos_q = pipe.get_featmap_corrs(gb, outputs, results, i)
        if corrs is None:
            assert rows[i]["corrs"] is None
            continue
        assert torch.equal(corrs, rows[i]["corrs"]) and pos_a.shape == (500, 32)
        pose = pipe.get_pose(gb, corrs, i)
        assert pose.dtype == torch.float32 and pose.device.type == "cpu"
        np.testing.assert_allclose(pose.numpy(), rows[i]["pred_pose_rel"].numpy(), atol=1e-6)


def test_mask_postproc_bit_exact():
    need_gpu()
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(4, 1, 192, 192, generator=g)
    logits[0, 0, :5, :5] = 0.0
    logits[3] = -5.0                                   # empty prediction -> IoU of an empty union is NaN / 0
    gt = (torch.rand(4, 224, 224, generator=g) > 0.6).to(torch.uint8)
    gt[3] = 0
    mp = mask_postproc(logits.cuda(), gt.cuda(), (192, 192), 0.5)
    ref_mask = oracle.predicted_mask(logits, 0.5)
    gt_c = torch.nn.functional.interpolate(gt.unsqueeze(1).float(), (192, 192), mode="nearest").squeeze(1)
    assert torch.equal(mp["pred"].cpu().long(), ref_mask)
    assert torch.equal(mp["gt_resized"].cpu().float(), gt_c)
    iou = oracle.mask_iou(gt_c, ref_mask)
    got = mp["iou"].cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(iou)) and torch.equal(got[~torch.isnan(got)], iou[~torch.isnan(iou)])
    assert mp["n_pred"].cpu().tolist() == [int((ref_mask[i] == 1).sum()) for i in range(4)]
    for i in range(4):
        assert torch.equal(oracle.resize_mask_nearest(gt[i], (192, 192)), mp["gt_resized"][i].cpu())


def test_test_step_feeds_evaluator(system):
    """test_step with an Evaluator attached: the reference's register_test / register_test_failure bookkeeping
    (pipeline.py:321-350) in pair order, pose errors on the GPU; compared with the oracle's arithmetic on the same poses."""
    import eval_oracle
    from oryon_b200.utils.evaluator import Evaluator, format_sym_set
    _, solver, psd = system
    outputs, batch = synth.planted_network_outputs(21, 4)
    obj = synth.eval_objects(0)
    batch["cls_id"] = [1, 2, 3, 2]
    batch["anchor"]["mask"][2] = 0                      # pair 2 fails the detection test in oracle-mask mode
    ev = Evaluator("planted", compute_vsd=False, compute_iou=True, device="cuda:0")
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    args = dict(ARGS, test=dict(ARGS["test"], mask="oracle"))
    pipe = FPM_Pipeline(args, test_model=True, model=_PlantedModel({k: v.cuda() for k, v in outputs.items()}), pointdsc_solver=solver,
                        evaluator=ev)
    pipe.on_test_start()
    rows = pipe.test_step(_cuda_batch(batch), 0)
    assert [r["status"] for r in rows] == ["ok", "ok", "invalid_mask", "ok"]
    assert ev.metrics["instance_id"] == batch["instance_id"] and ev.metrics["cls_id"] == batch["cls_id"]
    assert ev.counts["Missing segm"] == [0, 0, 1, 0]
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    for i in (0, 1, 3):
        ref = eval_oracle.pose_errors(obj["models"], syms, [batch["cls_id"][i]], rows[i]["pred_pose"].numpy()[None],
                                      batch["query"]["pose"][i].numpy()[None], batch["query"]["camera"][i].numpy()[None])[0]
        assert abs(ev.metrics["T error"][i] - ref[1]) < 1e-9
        assert ev.metrics["ADD(S)-0.1d"][i] == float(ref[2] <= ev.add_diams[batch["cls_id"][i]] * 0.1)
        assert ev.metrics["MSSD"][i] == (ref[4] < ev.mssd_rec * obj["diams"][batch["cls_id"][i]]).mean()
    assert ev.metrics["R error"][2] == 0.0 and ev.metrics["MSSD"][2] == 0.0


def test_decoded_frames_to_metrics_end_to_end(system, tmp_path):
    """The whole widened path in one piece: decoded uint8 frames -> GpuCollate (N1: staging on the GPU) -> test_step
    (network, masks, matching, lifting, PointDSC) -> prediction CSV (N3) -> Evaluator on the CUDA backend (N2), and the CSV
    read back by dict_from_preds carries the poses the evaluator saw."""
    from oryon_b200.datasets import GpuCollate
    from oryon_b200.utils.evaluator import Evaluator, dict_from_preds
    import stage_oracle
    model, solver, _ = system
    B = 3
    frames = synth.raw_frames(11, 2 * B)
    pairs = [synth.synthetic_rgbd_pair(900 + i) for i in range(B)]
    K = np.asarray(synth.NOCS_INTRINSICS).reshape(3, 3)
    tokens = sb.synthetic_tokens(7, 1)

    def item(f, depth, tag, i):
        return dict(rgb=f["rgb"], mask=f["mask"], depth=depth.numpy(), camera=K, instance_id=f"scene{i} {tag}{i} mug",
                    metadata=dict(mask_ids=[f["mask_id"]], poses=[np.eye(4)]))

    data = [(item(frames[2 * i], pairs[i]["depth_a"], 1, i), item(frames[2 * i + 1], pairs[i]["depth_q"], 2, i), ["mug"] * 81,
             np.eye(4), 1 + i % 3, f"pair{i}", True) for i in range(B)]
    batch = GpuCollate((224, 224), "cuda:0")(data)
    assert torch.equal(batch["anchor"]["rgb"][1].cpu(), stage_oracle.stage_rgb(frames[2]["rgb"], (224, 224)))
    batch["prompt_tokens"] = tokens.expand(B, -1, -1).contiguous()
    obj = synth.eval_objects(0)
    ev = Evaluator("e2e", compute_vsd=False, compute_iou=True, device="cuda:0")
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    args = dict(ARGS, test=dict(ARGS["test"], mask="oracle"))
    pipe = FPM_Pipeline(args, test_model=True, model=model, pointdsc_solver=solver, evaluator=ev)
    csv = tmp_path / "pred.csv"
    pipe.on_test_start(str(csv))
    rows = pipe.test_step(batch, 0)
    pipe.on_test_end()
    assert len(rows) == B and all(r["status"] in ("ok", "no_corrs", "invalid_mask") for r in rows)
    assert len(ev.metrics["R error"]) == B and ev.metrics["instance_id"] == [f"pair{i}" for i in range(B)]
    preds, ia, iq, present = dict_from_preds(str(csv))
    assert present and len(preds) == B
    for i, r in enumerate(rows):
        key = f"scene{i}_1{i}_scene{i}_2{i}_mug"
        np.testing.assert_array_equal(preds[key].astype(np.float32), r["pred_pose_rel"].numpy()[:3])
        assert np.float32(ia[key]) == np.float32(r["iou_a"])


@pytest.mark.parametrize("mask_mode", ["oracle", "predicted"])
def test_pipelined_steps_equal_unpipelined(system, mask_mode, tmp_path):
    """``test.pipelined`` (the tail of batch k on a second stream under the network pass of batch k+1) must change nothing but
    WHEN rows are returned: same CSV bytes, same statuses, correspondences and poses, rows delivered one call later and the
    last batch by ``flush`` / ``on_test_end``."""
    model, solver, _ = system
    batches = [_stacked_batch(synth.synthetic_batch(40 + k, 3, empty_mask_pairs=(1,) if k == 1 else ())) for k in range(3)]
    out = {}
    for mode in (False, True):
        args = dict(ARGS, test=dict(ARGS["test"], mask=mask_mode, pipelined=mode))
        pipe = FPM_Pipeline(args, test_model=True, model=model, pointdsc_solver=solver)
        path = str(tmp_path / f"pred_{int(mode)}.csv")
        pipe.on_test_start(path)
        returned = [pipe.test_step(dict(b), k) for k, b in enumerate(batches)]
        returned.append(pipe.on_test_end())
        torch.cuda.synchronize()
        out[mode] = (returned, list(pipe.rows), open(path).read())
    plain, piped = out[False], out[True]
    assert [len(r) for r in plain[0]] == [3, 3, 3, 0] and [len(r) for r in piped[0]] == [0, 3, 3, 3]
    assert piped[2] == plain[2] and len(plain[2].splitlines()) == 9
    for a, b in zip(plain[1], piped[1]):
        assert a["status"] == b["status"] and a["instance_id_a"] == b["instance_id_a"]
        assert torch.equal(a["pred_pose_rel"], b["pred_pose_rel"]) and a["iou_a"] == b["iou_a"] or (a["iou_a"] != a["iou_a"])
        assert (a["corrs"] is None) == (b["corrs"] is None) and (a["corrs"] is None or torch.equal(a["corrs"], b["corrs"]))
    if mask_mode == "oracle":
        assert [r["status"] for r in plain[1]][4] == "invalid_mask"
