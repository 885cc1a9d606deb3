"""Drop-in evaluators: same class names, constructor signatures, methods, attributes and log
lines as the reference's ``tao_amodal.evaluation`` package, computed by the CUDA library."""
