"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement of the reference's ViT encoder hot path (khuongnd6/ViT_torch): the un-vendored
DINO ViT (facebookresearch/dino@main, unpinned: call site models/vision_all.py:156), the timm (~0.4.9-0.4.12,
unpinned in requirements.txt:12) Mlp / PatchEmbed / VisionTransformer pieces the reference imports
(models/cait.py:8-10, models/deit.py:7-9), the reference's own CaiT (models/cait.py:21-253) and DeiT wrappers
(models/deit.py:20-91), the classifier head and init regimes of models/vision_all.py:154-221,299-329, and the
optimisation step of utils_network.py:406-452.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and `--impl reference`) may import this package,
and only as the checker / the timed CPU baseline. vit_torch_b200/ never imports it.

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path, and the
DINO / timm arithmetic lives in third-party code absent from the reference tree. The CaiT restatement IS pinned:
tests/test_oracle_vs_reference.py imports the reference's own models/cait.py (with a stub of the five timm symbols it
uses) in the build container and checks bit-level agreement; tests/golden/ holds vectors generated from that reference
file by tests/golden/make_golden.py. The DINO / DeiT restatement has no reference-side vectors: "parity unpinned"
for those two (stated in DESIGN.md); they share Block / Attention / Mlp code with the pinned CaiT path where the math
is the same (Mlp, LayerNorm, PatchEmbed) and are additionally checked against torch.nn.MultiheadAttention-free
first-principles formulas in tests/test_oracle.py.
"""
