"""Import shims that let the *reference's own modules* load read-only from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oryon_b200/`` may import this file.  It exists in
the build container only (``/root/reference`` does not travel to the GPU box); its single user
is ``oracle/make_golden.py``, which runs the unmodified reference code on seeded inputs and
commits the outputs as fixtures under ``tests/golden/``.

The shims cover packages the reference imports for type annotations or thin helpers and which
are not installed in this image:

* ``omegaconf``          -- ``DictConfig`` / ``OmegaConf`` are used as annotations only
                            (reference utils/misc.py:8, net.py:4).
* ``timm.models.layers`` -- ``Mlp``, ``DropPath``, ``to_2tuple``, ``to_ntuple`` used by
                            reference models/fusion.py:7.  timm 0.9.12 semantics restated:
                            Mlp = fc1 -> act -> drop -> fc2 -> drop; DropPath is identity in eval.
* ``easydict``           -- attribute dict used by reference utils/pointdsc/init.py:5.
"""
import collections.abc
import os
import sys
import types
from itertools import repeat

REFERENCE_ROOT = os.environ.get("ORYON_REFERENCE_ROOT", "/root/reference")


def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(repeat(x, n))
    return parse


def install():
    """Register the stub modules and put the reference on sys.path (idempotent)."""
    import torch.nn as nn

    sys.dont_write_bytecode = True  # never write __pycache__ into the read-only reference
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")

        class DictConfig(dict):
            __getattr__ = dict.get

        class OmegaConf:  # annotation-only in the hot path
            pass

        m.DictConfig, m.OmegaConf = DictConfig, OmegaConf
        sys.modules["omegaconf"] = m

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class Mlp(nn.Module):
            def __init__(self, in_features, hidden_features=None, out_features=None,
                         act_layer=nn.GELU, drop=0.0):
                super().__init__()
                out_features = out_features or in_features
                hidden_features = hidden_features or in_features
                self.fc1 = nn.Linear(in_features, hidden_features)
                self.act = act_layer()
                self.drop1 = nn.Dropout(drop)
                self.fc2 = nn.Linear(hidden_features, out_features)
                self.drop2 = nn.Dropout(drop)

            def forward(self, x):
                return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0, scale_by_keep=True):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                assert not self.training or self.drop_prob == 0.0
                return x

        layers.Mlp, layers.DropPath = Mlp, DropPath
        layers.to_2tuple, layers.to_ntuple = _ntuple(2), _ntuple
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__

        m.EasyDict = EasyDict
        sys.modules["easydict"] = m

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))
