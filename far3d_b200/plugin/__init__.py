"""Mirror of the reference's `projects.mmdet3d_plugin` registration-by-import (projects/mmdet3d_plugin/__init__.py:1-9):
importing this package registers Far3D / VoVNet / FPN / FarHead / Detr3DTransformer* / DeformableFeatureAggregationCuda /
MultiheadAttention / NMSFreeCoder under the reference's names."""
from .backbone import FPN, VoVNet                                                        # noqa: F401
from .transformer import (DeformableFeatureAggregationCuda, Detr3DTemporalDecoderLayer,  # noqa: F401
                          Detr3DTransformer, Detr3DTransformerDecoder, FFN, MultiheadAttention)
from .head import MLN, FarHead, NMSFreeCoder                                             # noqa: F401
from .roi_head import YOLOXHeadCustom                                                    # noqa: F401
from .detector import Far3D                                                              # noqa: F401
from ..compat import mirror_into_mmcv

mirror_into_mmcv()
