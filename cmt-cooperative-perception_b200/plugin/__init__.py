"""Host-side mirror of the reference's mmdet3d_plugin classes on the hot path: same class names,
constructor arguments, forward signatures and state-dict keys, registered by name so reference
config dicts (`pts_bbox_head=dict(type='CmtHead', ...)`) build unchanged."""
from .registry import (ALL, ATTENTION, BBOX_CODERS, HEADS, TRANSFORMER, TRANSFORMER_LAYER,
                       TRANSFORMER_LAYER_SEQUENCE, ConfigDict, register_into_openmmlab)
from .attention import FlashAttention, FlashMHA, KVCache
from .petr_transformer import (FFN, PETRMultiheadAttention, PETRMultiheadFlashAttention, PETRTransformerDecoder,
                               PETRTransformerDecoderLayer)
from .cmt_transformer import CmtImageTransformer, CmtLidarTransformer, CmtTransformer
from .bbox_coder import MultiTaskBBoxCoder, denormalize_bbox
from .cmt_head import (CmtHead, CmtImageHead, CmtLidarHead, GroupLayerNorm1d, SeparateTaskHead, inverse_sigmoid,
                       multi_apply, pos2embed)
from .cmt_head_coop import (CmtHeadCoop, CmtImageHeadCoop, CmtLidarHeadCoop, filter_img_metas,
                            get_infrastructure_image_metas, get_vehicle_image_metas)
from .detector_glue import (attach_calibration, bbox3d2result, coop_simple_test, coop_simple_test_pts, device_calibration,
                            simple_test, simple_test_pts)


def build_head(cfg):
    """mmdet3d.models.builder.build_head equivalent: `cfg['type']` selects the class."""
    head = HEADS.build(cfg)
    head.eval()
    return head


register_into_openmmlab()
