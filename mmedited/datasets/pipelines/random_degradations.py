from ._placeholder import placeholder

_W = "mmedited/datasets/pipelines/random_degradations.py"
RandomScaleResize1 = placeholder("RandomScaleResize1", _W)
DegradationsWithShuffle1 = placeholder("DegradationsWithShuffle1", _W)
RandomBlur = placeholder("RandomBlur", _W)
RandomJPEGCompression = placeholder("RandomJPEGCompression", _W)
RandomResize = placeholder("RandomResize", _W)
RandomNoise = placeholder("RandomNoise", _W)
