"""TEST INFRASTRUCTURE ONLY — imports the *real* reference (rhasspy/larynx) read-only.

Only usable in the build container, where /root/reference exists.  It is used
to (a) pin oracle/port.py against the reference's own PyTorch modules and
(b) generate the golden fixtures under tests/golden/ (see make_golden.py).
Nothing in the product package, bench.py's GPU arm or the -m gpu tests imports it.

The reference needs four third-party modules that are not installed here
(dataclasses_json, gruut, onnxruntime, phonemes2ids).  None of them carries
arithmetic of the hot path; they are replaced by inert stubs.
"""
import dataclasses
import json
import sys
import types
import typing

REFERENCE_ROOT = "/root/reference"


def _install_stubs():
    if "dataclasses_json" not in sys.modules:
        m = types.ModuleType("dataclasses_json")

        class DataClassJsonMixin:
            def to_dict(self):
                return dataclasses.asdict(self)

            @classmethod
            def from_dict(cls, d):
                hints = typing.get_type_hints(cls)
                kwargs = {}
                for f in dataclasses.fields(cls):
                    if f.name not in d:
                        continue
                    v = d[f.name]
                    t = hints.get(f.name)
                    if dataclasses.is_dataclass(t) and isinstance(v, dict):
                        v = t.from_dict(v)
                    kwargs[f.name] = v
                return cls(**kwargs)

            @classmethod
            def from_json(cls, s):
                return cls.from_dict(json.loads(s))

        m.DataClassJsonMixin = DataClassJsonMixin
        sys.modules["dataclasses_json"] = m
    for name in ("gruut", "phonemes2ids"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "onnxruntime" not in sys.modules:
        ort = types.ModuleType("onnxruntime")

        class SessionOptions:
            graph_optimization_level = None

        class InferenceSession:
            def __init__(self, *a, **k):
                raise RuntimeError("onnxruntime is a stub in this container")

        class GraphOptimizationLevel:
            ORT_DISABLE_ALL = 0

        ort.SessionOptions = SessionOptions
        ort.InferenceSession = InferenceSession
        ort.GraphOptimizationLevel = GraphOptimizationLevel
        sys.modules["onnxruntime"] = ort


def available() -> bool:
    import os

    return os.path.isdir(REFERENCE_ROOT + "/glow_tts")


def import_reference():
    """Returns (glow_tts.models, glow_tts.config, hifi_gan.models, hifi_gan.config)."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import glow_tts.config as gcfg
    import glow_tts.models as gmodels
    import hifi_gan.config as hcfg
    import hifi_gan.models as hmodels

    return gmodels, gcfg, hmodels, hcfg


# -----------------------------------------------------------------------------
# Running the reference's own classes on synthetic reference-format voices
# -----------------------------------------------------------------------------


def load_reference_models(glow_dir, hifigan_dir):
    """Load through the reference's *own* loaders (glow_tts/checkpoint.py:26-68,
    hifi_gan/checkpoint.py:36-70) and apply the same post-load steps as
    larynx/glow_tts.py:94-95 and larynx/hifi_gan.py:99-100."""
    import contextlib
    import io
    import warnings
    from pathlib import Path

    gmodels, gcfg, hmodels, hcfg = import_reference()
    import glow_tts.checkpoint as gck
    import hifi_gan.checkpoint as hck

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(Path(glow_dir) / "config.json", "r", encoding="utf-8") as f:
            gconf = gcfg.TrainingConfig.load(f)
        glow = gck.load_checkpoint(Path(glow_dir) / "generator.pth", gconf, use_cuda=False).model
        glow.decoder.store_inverse()
        glow.eval()
        with open(Path(hifigan_dir) / "config.json", "r", encoding="utf-8") as f:
            hconf = hcfg.TrainingConfig.load(f)
        gen = hck.load_checkpoint(Path(hifigan_dir) / "generator.pth", hconf, use_cuda=False).training_model.generator
        gen.eval()
        with contextlib.redirect_stdout(io.StringIO()):
            gen.remove_weight_norm()
    return glow, gen


def run_reference_glow(glow, ids, noise_scale, length_scale, eps=None):
    """FlowGenerator.forward (glow_tts/models.py:308-354) with torch.randn_like replaced by `eps`
    (the only RNG draw on the path, models.py:348)."""
    import numpy as np
    import torch

    x = torch.from_numpy(np.asarray(ids, dtype=np.int64)).unsqueeze(0)
    xl = torch.tensor([x.shape[1]], dtype=torch.int64)
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):
        if eps is None:
            return torch.zeros_like(t)
        e = torch.from_numpy(eps)[:, : t.shape[-1]].unsqueeze(0)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.to(t.dtype)

    torch.randn_like = fake_randn_like
    try:
        with torch.no_grad():
            (mel, z_m, z_logs, _, z_mask), (x_m, x_logs, x_mask), (attn, logw, logw_) = glow(
                x, xl, noise_scale=noise_scale, length_scale=length_scale
            )
    finally:
        torch.randn_like = orig
    return {"mel": mel, "attn": attn, "logw": logw, "x_m": x_m, "z_mask": z_mask}
