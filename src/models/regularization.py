"""Drop-in for src/models/regularization.py.  Same call signature (an iterable of parameters -> scalar).
In the fused path the penalty's *gradient* (lambda*sign(p) / 2*lambda*p) is applied inside the Adam kernel;
these callables give the penalty value and keep the unfused autograd path working."""
import torch


class Regularization_L1:
    def __init__(self, reg_strength: float = 0.001) -> None:
        self.reg_strength = reg_strength
        self.kind = "L1"

    def __call__(self, model_parameters):
        return self.reg_strength * sum(p.abs().sum() for p in model_parameters)


class Regularization_L2:
    def __init__(self, reg_strength: float = 0.001) -> None:
        self.reg_strength = reg_strength
        self.kind = "L2"

    def __call__(self, model_parameters):
        return self.reg_strength * abs(sum(p.pow(2).sum() for p in model_parameters))
