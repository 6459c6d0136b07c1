"""Load the UNMODIFIED reference modules from `oracle/_ref/` (see oracle/make_ref.py) -- TEST / BASELINE INFRASTRUCTURE.

`load_latte()` returns the reference's `models/latte.py` as a module (its one third-party import, timm's `Mlp` /
`PatchEmbed`, comes from `oracle/ref_shim`), `load_diffusion()` the reference's `diffusion` package; both return None when
`oracle/_ref` was not materialised (then callers use the oracle port and label their numbers `kind: "port"`)."""
import importlib
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "models", "latte.py"))


def load_latte():
    if not available():
        return None
    shim = os.path.join(HERE, "ref_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    spec = importlib.util.spec_from_file_location("ref_latte", os.path.join(REF, "models", "latte.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_diffusion():
    if not os.path.exists(os.path.join(REF, "diffusion", "__init__.py")):
        return None
    spec = importlib.util.spec_from_file_location("ref_diffusion", os.path.join(REF, "diffusion", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(REF, "diffusion")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_diffusion"] = mod
    spec.loader.exec_module(mod)
    return mod


def build_latte(model_name: str, cfg, sd):
    """The reference model for `model_name` (e.g. "Latte-XL/2") with the seeded weights `sd` loaded strictly."""
    ref = load_latte()
    if ref is None:
        return None
    m = ref.Latte_models[model_name](input_size=cfg.input_size, num_classes=cfg.num_classes, num_frames=cfg.num_frames,
                                     learn_sigma=cfg.learn_sigma, extras=cfg.extras)
    m.load_state_dict(sd, strict=True)
    return m.eval()
