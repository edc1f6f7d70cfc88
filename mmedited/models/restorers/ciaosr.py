from ciaosr_b200.restorers import CiaoSR  # noqa: F401
