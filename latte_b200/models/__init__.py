"""Drop-in for the reference's `models` package factory (Vchitect/Latte models/__init__.py:31-51).

`sample/sample.py:56`, `sample/sample_ddp.py:88` and `train.py:90` call `get_models(args)`; putting this
package ahead of the reference's on sys.path swaps the denoiser and nothing else."""
from latte_b200.latte import Latte, Latte_models  # noqa: F401  (absolute: also importable as top-level `models` via a symlink)


def get_models(args):
    name = args.model
    if "LatteIMG" in name:
        raise NotImplementedError("LatteIMG (video+image joint training variant, models/latte_img.py) is not built")
    if "LatteT2V" in name:
        # models/__init__.py:40-41
        from latte_b200.latte_t2v import LatteT2V
        return LatteT2V.from_pretrained(args.pretrained_model_path, subfolder="transformer", video_length=args.video_length)
    if "Latte" in name:
        # same keyword set as the reference factory (models/__init__.py:42-49)
        return Latte_models[name](input_size=args.latent_size, num_classes=args.num_classes,
                                  num_frames=args.num_frames, learn_sigma=args.learn_sigma, extras=args.extras)
    raise ValueError(f"{name} Model Not Supported!")
