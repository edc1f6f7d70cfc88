from ciaosr_b200.pipelines import (GenerateCoordinateAndCell, GenerateCoordinateAndCell1,  # noqa: F401
                                   GenerateCoordinateAndCell2)
