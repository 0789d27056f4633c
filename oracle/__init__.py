"""CPU oracle for the Far3D per-frame forward path.  TEST INFRASTRUCTURE ONLY.

This package restates the reference's algorithm (megvii-research/Far3D @ 5efb9d7,
`projects/mmdet3d_plugin/...`) in plain PyTorch-CPU fp32 (plus one scalar C file
for the bilinear-sampling index/mask arithmetic).  Every function cites the
reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it - and there only as the checker or as the
reported CPU baseline, never as the product path.  Nothing under `far3d_b200/`
imports `oracle`.

PARITY PINNED AGAINST THE REFERENCE RUN HERE: the reference ships no tests or
golden vectors, and mmcv / mmdet / mmdet3d cannot be installed offline - but the
reference's OWN modules (VoVNet, Far3D, FarHead, YOLOXHeadCustom, DepthPredictor,
Detr3DTransformer/Decoder/TemporalDecoderLayer, DeformableFeatureAggregationCuda,
MLN, positional encoders, NMSFreeCoder) import and execute on CPU, unmodified, once
the third-party symbols they use are supplied (tests/golden/ref_shims.py restates
those: ConvModule, mmcv MultiheadAttention/FFN, mmdet FPN, MlvlPointGenerator, and
mmcv's own pure-PyTorch statement of multi-scale deformable attention).
tests/golden/make_ref_golden.py runs them on seeded inputs with this oracle's
state_dict loaded STRICTLY (1065 identical names/shapes at the full config) and
commits the outputs (ref_modules.npz, ref_tiny_model.npz, ref_state_dict_full.json);
tests/test_ref_golden.py holds the oracle to them at 2e-5 ... 2e-4 relative, and
tests/test_gpu_ref_golden.py holds the CUDA path to them at the 1e-3 bar.
What stays restated rather than executed: the third-party pieces listed above
(their published algorithms at mmcv-full 1.6.2 / mmdet 2.28.2), cross-checked by
(i) the reference's own in-repo `grid_sample` form of the sampling math
(`models/utils/sparse_blocks.py:234-255`), (ii) an independent scalar statement of
mmcv's `ms_deformable_im2col` bilinear rule (oracle/msda.py, oracle/deform_agg_ref.c)
and (iii) torch's own `nn.MultiheadAttention` / `F.grid_sample`.
"""
