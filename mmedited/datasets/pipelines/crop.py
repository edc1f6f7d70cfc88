from ._placeholder import placeholder

PairedRandomCropwScale = placeholder("PairedRandomCropwScale", "mmedited/datasets/pipelines/crop.py")
