from ciaosr_b200.restorers import BasicRestorer  # noqa: F401
