"""Mirror of the reference ``utils/pointdsc/init.py``: ``get_pointdsc_solver`` (:32-57) and
``get_pointdsc_pose`` (:10-29), same names, argument meaning and return types.  The registration itself
(reference models/pointdsc/PointDSC.py test-mode forward) runs in liboryon_b200.so
(``oryon_pointdsc_load`` / ``oryon_pointdsc_pose``); there is no PyTorch implementation behind it.
"""
from __future__ import annotations

import ctypes
import json
from ctypes import c_double, c_int32, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from ... import _lib
from ..._torch_glue import as_device, ptr, require_cuda, stream_ptr


class PointDSCSolver:
    """What ``get_pointdsc_solver`` returns in place of the reference's ``nn.Module``: the constructor
    arguments of reference PointDSC.py:79-91 plus the ``state_dict``; the weights live in the library
    handle of ``device`` (BatchNorm folded, transposed) after the first call."""

    def __init__(self, state_dict: Dict[str, Tensor], in_dim: int = 6, num_layers: int = 6, num_channels: int = 128,
                 num_iterations: int = 10, ratio: float = 0.1, inlier_threshold: float = 0.10, sigma_d: float = 0.10,
                 k: int = 40, nms_radius: float = 0.10, device: Optional[torch.device] = None):
        self.in_dim, self.num_layers, self.num_channels = int(in_dim), int(num_layers), int(num_channels)
        self.num_iterations, self.ratio, self.k = int(num_iterations), float(ratio), int(k)
        self.inlier_threshold, self.nms_radius = float(inlier_threshold), float(nms_radius)
        # nn.Parameter sigma (1.0 unless the checkpoint says otherwise) and the sigma_spat buffer
        self.sigma = float(state_dict["sigma"].reshape(-1)[0]) if "sigma" in state_dict else 1.0
        self.sigma_d = float(state_dict["sigma_spat"].reshape(-1)[0]) if "sigma_spat" in state_dict else float(sigma_d)
        self._packed = self._pack(state_dict)
        self._loaded_on: Optional[int] = None
        self.device = torch.device(device) if device is not None else None

    # the reference calls .eval() / .to(device) on the module it gets back
    def eval(self):
        return self

    def to(self, device):
        self.device = torch.device(device)
        return self

    def parameters(self):
        return iter(())

    def _names(self) -> List[str]:
        bn = ("weight", "bias", "running_mean", "running_var")
        names = ["encoder.layer0.weight", "encoder.layer0.bias"]
        for i in range(self.num_layers):
            p = f"encoder.blocks.PointCN_layer_{i}"
            names += [f"{p}.0.weight", f"{p}.0.bias"] + [f"{p}.1.{s}" for s in bn]
            p = f"encoder.blocks.NonLocal_layer_{i}"
            names += [f"{p}.fc_message.0.weight", f"{p}.fc_message.0.bias"] + [f"{p}.fc_message.1.{s}" for s in bn]
            names += [f"{p}.fc_message.3.weight", f"{p}.fc_message.3.bias"] + [f"{p}.fc_message.4.{s}" for s in bn]
            names += [f"{p}.fc_message.6.weight", f"{p}.fc_message.6.bias"]
            for proj in ("projection_q", "projection_k", "projection_v"):
                names += [f"{p}.{proj}.weight", f"{p}.{proj}.bias"]
        for j in (0, 2, 4):
            names += [f"classification.{j}.weight", f"classification.{j}.bias"]
        return names

    def _pack(self, sd: Dict[str, Tensor]) -> Tensor:
        missing = [n for n in self._names() if n not in sd]
        if missing:
            raise KeyError(f"PointDSC state_dict lacks {missing[:4]}{'...' if len(missing) > 4 else ''}")
        return torch.cat([sd[n].detach().to("cpu", torch.float32).reshape(-1) for n in self._names()]).contiguous()

    def ensure_loaded(self, dev: torch.device) -> None:
        if self._loaded_on == dev.index:
            return
        cfg = _lib.PointDSCConfig(self.in_dim, self.num_layers, self.num_channels, self.num_iterations, self.k, 0,
                                  self.ratio, self.sigma_d, self.sigma, self.nms_radius, self.inlier_threshold)
        _lib.check(_lib.load().oryon_pointdsc_load(_lib.handle(dev.index), ctypes.byref(cfg), c_void_p(self._packed.data_ptr()),
                                                   self._packed.numel(), stream_ptr(dev)))
        self._loaded_on = dev.index


def get_pointdsc_solver(ckpt_path: str, device) -> PointDSCSolver:
    """Initialises the pretrained PointDSC solver from ``<ckpt_path>/snapshot/PointDSC_3DMatch_release``
    (reference utils/pointdsc/init.py:32-57): hyper-parameters from ``config.json``, ``nms_radius`` fed from
    ``config.inlier_threshold`` (:49), weights from ``models/model_best.pkl`` (non-strict, :51)."""
    root = f"{ckpt_path}/snapshot/PointDSC_3DMatch_release"
    config = json.load(open(f"{root}/config.json", "r"))
    sd = torch.load(f"{root}/models/model_best.pkl", map_location="cpu")
    return PointDSCSolver(sd, in_dim=config["in_dim"], num_layers=config["num_layers"], num_channels=config["num_channels"],
                          num_iterations=config["num_iterations"], ratio=config["ratio"], sigma_d=config["sigma_d"],
                          k=config["k"], nms_radius=config["inlier_threshold"], device=device)


def pointdsc_poses(model: PointDSCSolver, pcd1, pcd2, *, counts: Optional[Sequence[int]] = None, return_debug: bool = False):
    """Batched form: ``P`` correspondence sets ``[n_p,3]`` -> ``[P,4,4]`` float32 on the GPU (one library call).
    ``pcd1/pcd2``: sequences of ``[n_p,3]`` tensors, or padded device tensors ``[P,cap,3]`` with ``counts[p]`` valid rows."""
    dev = model.device if (model.device is not None and model.device.type == "cuda") else require_cuda()
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    model.ensure_loaded(dev)
    P = len(pcd1)
    if counts is not None:
        if not (isinstance(pcd1, Tensor) and isinstance(pcd2, Tensor)) or pcd1.shape != pcd2.shape or pcd1.dim() != 3 or pcd1.shape[2] != 3:
            raise ValueError("pointdsc_poses: with `counts`, pcd1 and pcd2 must be padded [P,cap,3] tensors")
        ns = [int(v) for v in counts]
        cap = pcd1.shape[1]
        if len(ns) != P or max(ns) > cap:
            raise ValueError("pointdsc_poses: `counts` does not fit the padded clouds")
        src, tgt = as_device(pcd1, dev, torch.float32), as_device(pcd2, dev, torch.float32)
    else:
        ns = [int(a.shape[0]) for a in pcd1]
        if P == 0 or any(a.shape != b.shape or a.dim() != 2 or a.shape[1] != 3 for a, b in zip(pcd1, pcd2)):
            raise ValueError("pointdsc_poses: pcd1[p] and pcd2[p] must both be [n_p,3]")
        cap = max(ns)
        src = torch.zeros(P, cap, 3, dtype=torch.float32, device=dev)
        tgt = torch.zeros(P, cap, 3, dtype=torch.float32, device=dev)
        for p in range(P):
            src[p, :ns[p]] = as_device(pcd1[p], dev, torch.float32)
            tgt[p, :ns[p]] = as_device(pcd2[p], dev, torch.float32)
    out = torch.empty(P, 4, 4, dtype=torch.float32, device=dev)
    n_arr = (c_int32 * P)(*ns)
    dbg_struct, dbg = None, None
    if return_debug:
        smax = max(1, max(int(n * model.ratio) for n in ns))
        dbg = dict(conf=torch.zeros(P, cap, dtype=torch.float32, device=dev),
                   features=torch.zeros(P, cap, model.num_channels, dtype=torch.float32, device=dev),
                   seeds=torch.full((P, smax), -1, dtype=torch.int32, device=dev),
                   fitness=torch.zeros(P, smax, dtype=torch.int32, device=dev),
                   initial_trans=torch.zeros(P, 4, 4, dtype=torch.float32, device=dev),
                   best_seed=torch.zeros(P, dtype=torch.int32, device=dev))
        dbg_struct = _lib.PointDSCDebug(ptr(dbg["conf"]), ptr(dbg["features"]), ptr(dbg["seeds"]), ptr(dbg["fitness"]), smax, 0,
                                        ptr(dbg["initial_trans"]), ptr(dbg["best_seed"]))
    _lib.check(_lib.load().oryon_pointdsc_pose(_lib.handle(dev.index), ptr(src), ptr(tgt), n_arr, P, cap, ptr(out),
                                               ctypes.byref(dbg_struct) if dbg_struct is not None else None, stream_ptr(dev)))
    return (out, dbg) if return_debug else out


def get_pointdsc_pose(pointdsc_model: PointDSCSolver, pcd1: Tensor, pcd2: Tensor, device) -> Tensor:
    """``pcd1``, ``pcd2``: ``[N,3]`` points of the correspondences -> ``[4,4]`` float32 CPU tensor
    (reference utils/pointdsc/init.py:10-29)."""
    if device is not None and torch.device(device).type == "cuda":
        pointdsc_model.to(device)
    return pointdsc_poses(pointdsc_model, [pcd1], [pcd2])[0].cpu().to(torch.float32)
