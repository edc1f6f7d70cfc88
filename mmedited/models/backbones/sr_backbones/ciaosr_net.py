from ciaosr_b200.generators import (LocalImplicitSREDSR, LocalImplicitSRNet,  # noqa: F401
                                    LocalImplicitSRRDN, LocalImplicitSRSWINIR)
