from ciaosr_b200.restorers import RealCiaoSR  # noqa: F401
