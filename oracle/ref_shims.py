"""Import shims that let the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in `founddiff_b200/` may import this file.  It is used by
`oracle/gen_golden.py` (fixture generation) and by the `-m "not gpu"` tests that cross-check the oracle
against the live reference when `/root/reference` is present (it is absent on the GPU box).

What is shimmed and why:
  * missing third-party modules the reference imports at module load but never uses on the sampling
    path (ipdb, Augmentor, accelerate, ema_pytorch, open_clip, lpips, timm, clip, kornia, pywt, skimage,
    matplotlib, lmdb, wandb...) -> fabricated stub modules;
  * `datasets` -> namespace clash with HuggingFace `datasets`; pre-registered to the reference directory;
  * `src.DACLIP.load` / `src.model_clipiqa.load` / `src.DADiff.load` download RN50 weights -> replaced by a
    random-init `CLIP(1024, 224, (3,4,6,3), 64, None, 77, 49408, 512, 8, 12)` (the RN50 hyper-parameters of
    `build_model`, src/DACLIP.py:608-640);
  * `load_file_from_url` (CLIP-IQA+ learned prompts) -> local (2,16,512) tensor;
  * `Dose-CLIP.pth` read from cwd (src/DADiff.py:595) -> written to a temp dir, cwd switched during build;
  * `selective_scan_cuda.fwd` (third-party CUDA-only extension, src/emamba2.py:23-34,152) -> backed by the
    C restatement in oracle/selective_scan_ref.c.
No reference source is copied; the reference is imported from where it lies.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("FOUNDDIFF_REFERENCE", "/root/reference")

_STUB_ROOTS = (
    "ipdb", "Augmentor", "accelerate", "ema_pytorch", "open_clip", "lpips", "timm", "clip", "kornia",
    "pywt", "skimage", "matplotlib", "lmdb", "wandb", "selective_scan_cuda", "cv2", "scipy", "PIL",
    "torchvision",
)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


class _Stub(types.ModuleType):
    """A module whose every attribute is another stub / a do-nothing callable class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _really_missing(name: str) -> bool:
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


_installed = False


def install():
    """Install stubs + patches; returns the imported reference modules as a namespace."""
    global _installed
    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must be imported before the stub finder shadows anything)

    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    if not _installed:
        missing = [r for r in _STUB_ROOTS if _really_missing(r)]
        sys.meta_path.append(_StubFinder(missing))

        # real bodies for the few stubbed names that are actually executed
        if "timm" in missing:
            import timm.models.layers as tl
            import timm.models.registry as tr

            class DropPath(nn.Identity):
                def __init__(self, *a, **k):
                    super().__init__()
            tl.DropPath = DropPath
            tl.trunc_normal_ = lambda t, *a, **k: t
            tr.register_model = lambda f: f
        if "lpips" in missing:
            import lpips

            class LPIPS(nn.Module):
                def __init__(self, *a, **k):
                    super().__init__()
            lpips.LPIPS = LPIPS
        if "clip" in missing:
            import clip

            def tokenize(texts, *a, **k):
                # text tower output is discarded by Unet.forward (src/DADiff.py:692); any (n,77) long tensor
                # with a unique arg-max per row works.
                t = torch.zeros(len(texts), 77, dtype=torch.long)
                t[:, 0] = 49406
                for i in range(len(texts)):
                    t[i, 1:21] = torch.arange(1000 + i, 1020 + i)
                    t[i, 21] = 49407
                return t
            clip.tokenize = tokenize
        if "ema_pytorch" in missing:
            import ema_pytorch

            class EMA(nn.Module):
                def __init__(self, model, *a, **k):
                    super().__init__()
                    self.ema_model = model
            ema_pytorch.EMA = EMA

        # `datasets` namespace clash
        ds = types.ModuleType("datasets")
        ds.__path__ = [os.path.join(REFERENCE_ROOT, "datasets")]
        sys.modules["datasets"] = ds
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)

        # selective scan: bind the C restatement
        from oracle import scan_cpu
        ssc = sys.modules.get("selective_scan_cuda")
        if ssc is None or isinstance(ssc, _Stub) or "selective_scan_cuda" in missing:
            ssc = types.ModuleType("selective_scan_cuda")
            sys.modules["selective_scan_cuda"] = ssc

        def fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
            assert z is None
            out = scan_cpu.selective_scan_fwd(u, delta, A, B, C, D, delta_bias, delta_softplus)
            return out, None
        ssc.fwd = fwd
        _installed = True

    import src.DACLIP as DACLIP
    import src.model_clipiqa as model_clipiqa

    def _rn50(*a, **k):
        return DACLIP.CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12).float().eval()

    def _rn50_b(*a, **k):
        return model_clipiqa.CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12).float().eval()

    tmp = tempfile.mkdtemp(prefix="fd_ref_")
    prompts = os.path.join(tmp, "prompts.pth")
    torch.save(torch.zeros(2, 16, 512), prompts)
    DACLIP.load = _rn50
    DACLIP.load_file_from_url = lambda *a, **k: prompts
    model_clipiqa.load = _rn50_b
    model_clipiqa.load_file_from_url = lambda *a, **k: prompts

    import src.DADiff as DADiff
    DADiff.load = _rn50_b
    import src.emamba2 as emamba2
    import src.denoising_diffusion_pytorch as ddp

    ns = types.SimpleNamespace(DADiff=DADiff, DACLIP=DACLIP, emamba2=emamba2, ddp=ddp,
                               model_clipiqa=model_clipiqa, tmpdir=tmp)
    return ns


def build_reference(sampling_timesteps=2, dim=64, dim_mults=(1, 2, 4, 8), image_size=512, num_unet=1,
                    objective="pred_res", test_res_or_noise="res", loss_type="l2"):
    """Construct UnetRes + ResidualDiffusion exactly as train.py:97-119 does (random init); the defaults are the shipped
    configuration (train.py:78-82), `num_unet=2, objective='pred_res_noise', test_res_or_noise='res_noise'` the
    commented-out one (train.py:75-77)."""
    import torch
    ns = install()
    cwd = os.getcwd()
    os.chdir(ns.tmpdir)
    try:
        if not os.path.exists("Dose-CLIP.pth"):
            torch.save(ns.DACLIP.CLIPIQA(model_type="clipiqa+").state_dict(), "Dose-CLIP.pth")
        model = ns.DADiff.UnetRes(dim=dim, dim_mults=dim_mults, num_unet=num_unet, condition=True,
                                  input_condition=False, objective=objective, test_res_or_noise=test_res_or_noise)
        diffusion = ns.DADiff.ResidualDiffusion(
            model, image_size=image_size, timesteps=1000, sampling_timesteps=sampling_timesteps,
            objective=objective, loss_type=loss_type, condition=True, sum_scale=0.01,
            input_condition=False, input_condition_mask=False, test_res_or_noise=test_res_or_noise)
    finally:
        os.chdir(cwd)
    return ns, model.eval(), diffusion.eval()
