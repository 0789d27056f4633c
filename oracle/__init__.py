"""CPU oracle for the Far3D per-frame forward path.  TEST INFRASTRUCTURE ONLY.

This package restates the reference's algorithm (megvii-research/Far3D @ 5efb9d7,
`projects/mmdet3d_plugin/...`) in plain PyTorch-CPU fp32 (plus one scalar C file
for the bilinear-sampling index/mask arithmetic).  Every function cites the
reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it - and there only as the checker or as the
reported CPU baseline, never as the product path.  Nothing under `far3d_b200/`
imports `oracle`.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
cannot be imported in the build container (mmcv / mmdet / mmdet3d absent, see
SURVEY.md section 8c), so this oracle is pinned only by (i) line-by-line
restatement of the reference Python, (ii) the reference's own in-repo
`grid_sample` restatement of the sampling math
(`models/utils/sparse_blocks.py:234-255`), which `oracle.msda` reproduces and
cross-checks against an independent scalar implementation of mmcv's
`ms_deformable_im2col` bilinear rule, and (iii) torch's own
`nn.MultiheadAttention` / `F.grid_sample` for the third-party pieces.
"""
