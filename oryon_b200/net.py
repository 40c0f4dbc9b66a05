"""Mirror of the reference ``net.py``: ``Oryon`` with ``forward(xs) -> {'featmap_a','featmap_q','mask_a','mask_q'}``
(net.py:142-167).  The whole network runs in liboryon_b200.so (``oryon_backbone_forward`` / ``oryon_text_forward``);
this class only owns the weight hand-over, the prompt-embedding cache and tensor plumbing.  No PyTorch arithmetic,
no fallback: without the library or an sm_100 device every call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_void_p
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._torch_glue import as_device, ptr, require_cuda, stream_ptr

N_PROMPTS, CTX, EMBED = 80, 77, 768
FEAT_C, FEAT_HW = 32, 192
DEFAULT_BPE = "pretrained_models/bpe_simple_vocab_16e6.txt.gz"


class Oryon:
    """``Oryon(args, device)`` keeps the reference's constructor shape; ``args`` may be ``None`` (the network
    dimensions are those of the released configuration, configs/config.yaml:30-40).  Weights come from a
    ``state_dict`` with the reference's names -- ``Oryon.state_dict()`` keys, i.e. the Lightning checkpoint keys
    with ``model.`` stripped."""

    def __init__(self, args=None, device="cuda", *, state_dict: Optional[Dict[str, Tensor]] = None, vis_layers: int = 24,
                 txt_layers: int = 12, precision: int = 2, max_pairs_per_pass: int = 32, tokenizer=None):
        self.args = getattr(args, "model", args)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.OryonError("oryon_b200.net.Oryon runs on a CUDA (sm_100) device only; there is no CPU path")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device()) \
            if torch.cuda.is_available() else require_cuda()
        self.vis_layers, self.txt_layers, self.precision = int(vis_layers), int(txt_layers), int(precision)
        self.max_pairs_per_pass = int(max_pairs_per_pass)
        if tokenizer is None and os.path.exists(DEFAULT_BPE):   # the path the reference hard-codes (models/vlm.py:23)
            from .models.tokenizer import SimpleTokenizer
            tokenizer = SimpleTokenizer(DEFAULT_BPE)
        self.tokenizer = tokenizer
        self.training = False
        self._loaded = False
        self._prompt_cache: Dict[tuple, Tensor] = {}
        self._load_error: Optional[str] = None
        if state_dict is None and args is not None:
            # the reference's constructor reads its pretrained files itself (net.py:27-34, :99-139); so does this one when they
            # lie where the reference's libraries put them.  Missing files are reported by the first forward, not silently skipped.
            from . import checkpoint
            try:
                state_dict = checkpoint.reference_layout_state_dict(args)
            except FileNotFoundError as e:
                self._load_error = str(e)
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # the reference calls these on the module
    def train(self, mode=True):
        return self

    def eval(self):
        return None  # reference net.py:87-89 returns None

    def to(self, device):
        return self

    def load_state_dict(self, state_dict: Dict[str, Tensor], strict: bool = False):
        """Hands every floating-point tensor of the network over to the library and packs them
        (``oryon_backbone_set_weight`` / ``oryon_backbone_finalize``).  Accepts keys with or without the
        Lightning ``model.`` prefix."""
        lib, h = _lib.load(), _lib.handle(self.device.index)
        n = 0
        for k, v in state_dict.items():
            if k.startswith("model."):
                k = k[len("model."):]
            if not k.startswith(("vlm.clip_model.", "guidance_backbone.", "fusion.", "decoder.")) or not torch.is_floating_point(v):
                continue
            if k.endswith("attn_mask") or ".visual.proj" in k or k.endswith("logit_scale"):
                continue  # derived buffers / unused by encode_image (vlm.py:56-59)
            t = v.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.oryon_backbone_set_weight(h, k.encode(), c_void_p(t.data_ptr()), t.numel()))
            n += 1
        cfg = _lib.BackboneConfig(self.vis_layers, self.txt_layers, self.precision, self.max_pairs_per_pass)
        _lib.check(lib.oryon_backbone_finalize(h, ctypes.byref(cfg), stream_ptr(self.device)))
        self._loaded = True
        return n

    # ---- prompts (vlm.py:63-86) -----------------------------------------------------------------------
    def encode_tokens(self, tokens: Tensor) -> Tensor:
        """``[n,77]`` token ids -> ``[n,768]`` prompt embeddings on the GPU."""
        tok = as_device(tokens, self.device, torch.int32)
        if tok.dim() != 2 or tok.shape[1] != CTX:
            raise ValueError(f"encode_tokens: expected [n,{CTX}] token ids, got {tuple(tok.shape)}")
        out = torch.empty(tok.shape[0], EMBED, dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().oryon_text_forward(_lib.handle(self.device.index), ptr(tok), tok.shape[0], ptr(out), stream_ptr(self.device)))
        return out

    def encode_prompt(self, prompts: Sequence[Sequence[str]]) -> Tensor:
        """``CLIPEncoder.encode_prompt``: drops the un-templated first prompt of every list (vlm.py:67), tokenises
        the remaining 80 and runs the text tower; one evaluation per distinct prompt list (the embeddings are a
        pure function of the strings)."""
        if self.tokenizer is None:
            raise _lib.OryonError("encode_prompt needs a tokenizer (oryon_b200.models.tokenizer.SimpleTokenizer with the CLIP BPE vocabulary)")
        out = []
        for plist in prompts:
            key = tuple(plist[1:])
            if key not in self._prompt_cache:
                self._prompt_cache[key] = self.encode_tokens(self.tokenizer(list(key)))
            out.append(self._prompt_cache[key])
        return torch.stack(out)

    # ---- network ----------------------------------------------------------------------------------------
    def forward_tensors(self, rgb_a: Tensor, rgb_q: Tensor, prompt_emb: Tensor, return_debug: bool = False):
        if not self._loaded:
            raise _lib.OryonError("Oryon: weights not loaded (load_state_dict)" + (f"; {self._load_error}" if self._load_error else ""))
        dev = self.device
        ra, rq = as_device(rgb_a, dev, torch.float32), as_device(rgb_q, dev, torch.float32)
        te = as_device(prompt_emb, dev, torch.float32)
        B = ra.shape[0]
        if ra.shape != (B, 3, 224, 224) or rq.shape != ra.shape or te.shape != (B, N_PROMPTS, EMBED):
            raise ValueError(f"forward: expected rgb [B,3,224,224] x2 and prompt embeddings [B,80,768]; got {tuple(ra.shape)}, "
                             f"{tuple(rq.shape)}, {tuple(te.shape)}")
        out = dict(featmap_a=torch.empty(B, FEAT_C, FEAT_HW, FEAT_HW, dtype=torch.float32, device=dev),
                   featmap_q=torch.empty(B, FEAT_C, FEAT_HW, FEAT_HW, dtype=torch.float32, device=dev),
                   mask_a=torch.empty(B, 1, FEAT_HW, FEAT_HW, dtype=torch.float32, device=dev),
                   mask_q=torch.empty(B, 1, FEAT_HW, FEAT_HW, dtype=torch.float32, device=dev))
        dbg, dbg_struct = None, None
        if return_debug:
            dbg = dict(clip_tokens=torch.empty(2 * B, 1024, 24, 24, device=dev), guid1=torch.empty(2 * B, 512, 24, 24, device=dev),
                       guid2=torch.empty(2 * B, 256, 48, 48, device=dev), guid3=torch.empty(2 * B, 128, 96, 96, device=dev),
                       fusion=torch.empty(2 * B, 128, 24, 24, device=dev))
            dbg_struct = _lib.BackboneDebug(B, 0, ptr(dbg["clip_tokens"]), ptr(dbg["guid1"]), ptr(dbg["guid2"]), ptr(dbg["guid3"]),
                                            ptr(dbg["fusion"]))
        _lib.check(_lib.load().oryon_backbone_forward(
            _lib.handle(dev.index), ptr(ra), ptr(rq), B, ptr(te), ptr(out["featmap_a"]), ptr(out["featmap_q"]), ptr(out["mask_a"]),
            ptr(out["mask_q"]), ctypes.byref(dbg_struct) if dbg_struct is not None else None, stream_ptr(dev)))
        return (out, dbg) if return_debug else out

    def forward(self, xs: dict) -> Dict[str, Tensor]:
        """``xs['anchor'|'query']['rgb'] [B,3,224,224]`` in [0,1] and ``xs['prompt']`` (81 strings per sample) -- or
        ``xs['prompt_tokens'] [B,80,77]`` / ``xs['prompt_emb'] [B,80,768]`` when the caller tokenised / cached --
        -> the reference's output dict (net.py:162-167)."""
        if "prompt_emb" in xs:
            emb = xs["prompt_emb"]
        elif "prompt_tokens" in xs:
            t = xs["prompt_tokens"]
            emb = self.encode_tokens(t.reshape(-1, CTX)).view(t.shape[0], N_PROMPTS, EMBED)
        else:
            emb = self.encode_prompt(xs["prompt"])
        return self.forward_tensors(xs["anchor"]["rgb"], xs["query"]["rgb"], emb)

    __call__ = forward


def gemm_counters(device_index: int = 0, with_tensor_flops: bool = False):
    """(launches, algorithmic FLOPs[, tensor-pipe FLOPs issued, fp16-equivalent]) of the tensor-core GEMM since the last call."""
    n, f, t = ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.load().oryon_gemm_counters(_lib.handle(device_index), ctypes.byref(n), ctypes.byref(f), ctypes.byref(t)))
    return (n.value, f.value, t.value) if with_tensor_flops else (n.value, f.value)
