class UniformRandomSampler:
    """Accepted and ignored (scripts/train_model.py:72-78 builds one when --use_placement_initializer is set)."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
