from ciaosr_b200.cross_scale_attention import CrossScaleAttention  # noqa: F401
