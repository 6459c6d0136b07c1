"""Materialise `oracle/_ref/`: a VERBATIM copy of the reference's pure-Python hot-path sources, made at build time.

TEST / BASELINE INFRASTRUCTURE.  The reference (Vchitect/Latte) is pure Python, so there is nothing to compile; what the
GPU box lacks is the source tree itself (`/root/reference` exists only in the build container).  This recipe copies the few
files of the path -- `models/latte.py` and the `diffusion/` package -- into `oracle/_ref/`, which is git-ignored (the
sources never enter this repository's history) but travels with the gpurun snapshot like a built `.so`.  `bench.py --impl
reference`, `cpu_baseline` and `gpu_eager_baseline` then time the UNMODIFIED reference module (through `oracle/ref_loader`
and the 40-line timm shim), and fall back to the oracle port -- saying so -- when `_ref` is absent.
Run by `__graft_entry__.build()` whenever `/root/reference` is present.  Never imported by `latte_b200/`."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["models/latte.py", "diffusion/__init__.py", "diffusion/gaussian_diffusion.py", "diffusion/respace.py",
         "diffusion/diffusion_utils.py", "diffusion/timestep_sampler.py"]


def materialise() -> bool:
    if not os.path.isdir(REF_ROOT):
        return os.path.isdir(OUT)
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("verbatim copies of /root/reference files (oracle/make_ref.py); git-ignored, shipped to the GPU box only\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref present:", materialise())
