from .predictor import Predictor, SAM2Classifier, get_predictor  # noqa: F401
