"""Stand-in for the `diffusers==0.24.0` names that /root/reference/models/latte_t2v.py:9-20 imports.

TEST INFRASTRUCTURE ONLY (never imported by latte_b200/).  diffusers is pinned by the reference
(environment.yml:13) but is neither vendored there nor installed in this image, and there is no network.
With this package on sys.path the UNMODIFIED reference module imports and runs, so that everything the
reference itself defines -- `LatteT2V.forward` control flow, `BasicTransformerBlock_`, `AdaLayerNormSingle`,
`FeedForward`, the temporal sin-cos table -- is executed as written when goldens are generated
(oracle/make_golden_t2v.py).

What IS restated here, from the published 0.24.0 behaviour, are the library leaves:
`Attention` + `AttnProcessor2_0`, `BasicTransformerBlock` (the spatial block), `PatchEmbed`, `CaptionProjection`,
`CombinedTimestepSizeEmbeddings`, `GELU`, the LoRA-compatible Linear/Conv and the config/model mixins.
Parity claims that rest on those leaves are "shim-restated", not reference-pinned; DESIGN.md section 3 says so.
Names the Latte-1 configuration never instantiates are placeholders that raise when constructed.
"""
__version__ = "0.24.0+latte_b200.shim"
