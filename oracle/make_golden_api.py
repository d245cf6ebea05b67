#!/usr/bin/env python
"""The drop-in boundary as data: names and call signatures of the reference's public API for the path, recorded from
the REAL reference (build container only):

    python oracle/make_golden_api.py        # -> tests/golden/api_signatures.json

Modules: k_diffusion/sampling.py (everything public), layers.Denoiser, external.DiscreteSchedule, config.load_config / make_model /
make_denoiser_wrapper, models.ImageTransformerDenoiserModelV2 (+ its spec dataclasses).  A signature is the ordered list of
(parameter name, kind, repr(default) or null)."""
import inspect
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as G


def sig_of(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        out.append([name, p.kind.name, None if p.default is inspect._empty else repr(p.default)])
    return out


def main():
    G._stub_missing()
    sys.path.insert(0, str(G.REF))
    import k_diffusion as K
    from k_diffusion.models import image_transformer_v2 as itv2
    api = {}
    S = K.sampling
    for name in sorted(dir(S)):
        obj = getattr(S, name)
        if name.startswith("_") or getattr(obj, "__module__", "") != "k_diffusion.sampling":
            continue
        if inspect.isfunction(obj):
            api[f"sampling.{name}"] = sig_of(obj)
        elif inspect.isclass(obj):
            api[f"sampling.{name}.__init__"] = sig_of(obj.__init__)
            for m, fn in inspect.getmembers(obj, inspect.isfunction):
                if not m.startswith("_") and m in ("__call__", "propose_step", "limiter", "dpm_solver_fast", "dpm_solver_adaptive", "t", "sigma"):
                    api[f"sampling.{name}.{m}"] = sig_of(fn)
            if "__call__" in obj.__dict__:
                api[f"sampling.{name}.__call__"] = sig_of(obj.__call__)
    for label, fn in (("layers.Denoiser.__init__", K.layers.Denoiser.__init__), ("layers.Denoiser.get_scalings", K.layers.Denoiser.get_scalings),
                      ("layers.Denoiser.forward", K.layers.Denoiser.forward),
                      ("external.DiscreteSchedule.__init__", K.external.DiscreteSchedule.__init__),
                      ("external.DiscreteSchedule.get_sigmas", K.external.DiscreteSchedule.get_sigmas),
                      ("external.DiscreteSchedule.sigma_to_t", K.external.DiscreteSchedule.sigma_to_t),
                      ("external.DiscreteSchedule.t_to_sigma", K.external.DiscreteSchedule.t_to_sigma),
                      ("config.load_config", K.config.load_config), ("config.make_model", K.config.make_model),
                      ("config.make_denoiser_wrapper", K.config.make_denoiser_wrapper),
                      ("models.ImageTransformerDenoiserModelV2.__init__", itv2.ImageTransformerDenoiserModelV2.__init__),
                      ("models.ImageTransformerDenoiserModelV2.forward", itv2.ImageTransformerDenoiserModelV2.forward),
                      ("utils.append_dims", K.utils.append_dims), ("utils.to_pil_image", K.utils.to_pil_image)):
        api[label] = sig_of(fn)
    (G.OUT / "api_signatures.json").write_text(json.dumps(api, indent=1))
    print("wrote", G.OUT / "api_signatures.json", len(api), "entries")


if __name__ == "__main__":
    main()
