"""ModelConfig: kwargs-bag configuration with the reference's defaults.

Reference: src/model_config.py:1-108.  Attribute names, defaults and ``to_dict()`` are part of the
drop-in contract (checkpoints store ``model.config.to_dict()``: src/train.py:304).
"""

_CNN_DEFAULTS = dict(
    image_size=(500, 500), in_channels=3 + 1 + 17, num_joints=17,
    heatmap_size=500, heatmap_sigma=10.0,
    initial_channels=64, initial_kernel_size=5, initial_stride=2,
    stage_channels=[128, 256, 512], stage_depths=[3, 4, 5], stage_strides=[2, 2, 2],
    stage_expand_ratios=[1, 3, 6],
    use_se_blocks=True, se_reduction=16, use_dual_path_blocks=True,
    global_pool_size=8, global_feature_dim=1024,
    regression_dims=[1024, 512], regression_dropout=0.2,
    activation="silu", normalization="batch",
    residual_scale=1.0, depthwise_kernel_size=3,
)

_TRANSFORMER_DEFAULTS = dict(
    num_joints=17, heatmap_sigma=2.0,
    vit_model_name="vit_base_patch16_384", vit_pretrained=True, vit_freeze_backbone=False,
    image_size=(512, 512), image_in_channels=4,
    heatmap_size=64, heatmap_patch_size=16, heatmap_in_channels=None,  # None -> num_joints
    transformer_heads=16, transformer_mlp_ratio=4.0,
    transformer_dropout_rate=0.1, transformer_attention_dropout_rate=0.1,
    num_cross_modal_layers=2, final_encoder_depth=4,
    activation="gelu",
    regression_hidden_dims=(1024, 512, 256), regression_dropout=0.25,
    transformer_embed_dim=768,
)


class ModelConfig:
    def __init__(self, model_type, **kwargs):
        if model_type == "cnn":
            table = _CNN_DEFAULTS
        elif model_type == "transformer":
            table = _TRANSFORMER_DEFAULTS
        else:
            raise ValueError(f"Unsupported model type: {model_type}")
        for key, default in table.items():
            value = kwargs.get(key, default)
            if isinstance(default, list) and value is default:
                value = list(default)
            setattr(self, key, value)
        if model_type == "transformer" and self.heatmap_in_channels is None:
            self.heatmap_in_channels = self.num_joints

    def to_dict(self):
        return {k: v for k, v in self.__dict__.items() if not callable(v) and not k.startswith("__")}
