from ciaosr_b200.refiners import MLPRefiner  # noqa: F401
