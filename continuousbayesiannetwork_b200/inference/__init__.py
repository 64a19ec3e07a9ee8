from .exact import ExactInference

INFERENCE_OBJS = {"exact": ExactInference}
