"""Model families: the tutorial's MNIST ConvNet and ResNet-18 (BASELINE.json config #3)."""
from .convnet import Net, PARAM_SHAPES, PARAM_NUMEL  # noqa: F401
from .resnet import ResNet18  # noqa: F401
