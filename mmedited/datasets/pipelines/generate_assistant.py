from ._placeholder import placeholder

_W = "mmedited/datasets/pipelines/generate_assistant.py"
GenerateCoordinateAndCell = placeholder("GenerateCoordinateAndCell", _W)
GenerateCoordinateAndCell1 = placeholder("GenerateCoordinateAndCell1", _W)
