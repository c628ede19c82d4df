class _Spec:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


Bounded = Composite = Unbounded = UnboundedContinuous = UnboundedDiscrete = _Spec
BoundedTensorSpec = CompositeSpec = UnboundedContinuousTensorSpec = UnboundedDiscreteTensorSpec = _Spec
