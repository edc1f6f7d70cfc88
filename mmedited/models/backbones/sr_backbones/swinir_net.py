from ciaosr_b200.swinir import SwinIR  # noqa: F401
